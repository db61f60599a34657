#!/bin/bash
mkdir -p gpurun_out
{
for v in 1 2 1 2; do
  echo "### PDK_SPLIT=$v"; PDK_SPLIT=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step[^,]*' | head -2
done
PDK_SPLIT=2 timeout 600 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_dit.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
} 2>&1 | tee gpurun_out/ab.log
