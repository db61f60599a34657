#!/bin/bash
mkdir -p gpurun_out
{
echo "== fused transition test"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -x -k "gemm or attention or transition or qkv" 2>&1 | tail -25
timeout 200 python tools/time_gemm.py 2>&1 | grep "fused\|adaln"
echo "== all gpu tests"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -5
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
} 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/trans.log
