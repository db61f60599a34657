#!/bin/bash
# GPU visit: validate the cta_group::2 GEMM tiling (probe 10), GPU parity tests with and without it, bench A/B, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== probe 10"
timeout 30 ./tests/cuda/umma_probe 10 0 2>&1 | tail -2 | tee gpurun_out/probe10.log
echo "  (exit $?)"
echo "== gemm test (pair)"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "gemm or linear" 2>&1 | tail -30 | tee gpurun_out/pytest_gemm_pair.log
echo "== pytest (pair)"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
echo "== bench nopair"
PDK_NO_PAIR=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_nopair.log
echo "== bench pair"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
wc -l gpurun_out/launches.csv
