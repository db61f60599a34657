#!/bin/bash
mkdir -p gpurun_out
{
echo "== gemm tests"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -x -k "gemm or qkv or transition or block" 2>&1 | tail -12
echo "### wide"; timeout 200 python tools/time_gemm.py 2>&1 | grep "us$"
echo "### PDK_NO_WIDE"; PDK_NO_WIDE=1 timeout 200 python tools/time_gemm.py 2>&1 | grep "us$"
echo "== all gpu tests"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -5
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo "== bench PDK_NO_WIDE"
PDK_NO_WIDE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
} 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/trans.log
