#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -x -k "fused" 2>&1 | tail -2
timeout 200 python tools/time_gemm.py 2>&1 | grep "fused"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step[^,]*' | head -1; done
} 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/ab.log
