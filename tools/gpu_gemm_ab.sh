#!/bin/bash
mkdir -p gpurun_out
{
echo "== gemm tests"
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dit.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -15
echo "### product"; timeout 120 python tools/time_gemm.py
echo "### product force pair"; PDK_FORCE_PAIR=1 timeout 120 python tools/time_gemm.py
echo "### product nopair"; PDK_NO_PAIR=1 timeout 120 python tools/time_gemm.py
echo "### NO_EPI"; PHYSDOCK_B200_LIB=/root/repo/build/dbg/libpdk_NO_EPI.so timeout 120 python tools/time_gemm.py
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
} 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/gemm_ab.log
