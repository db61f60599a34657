#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_dit.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step[^,]*' | head -2; done
} 2>&1 | tee gpurun_out/t.log
