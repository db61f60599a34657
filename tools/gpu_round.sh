#!/bin/bash
# One GPU-box visit: tcgen05 probes, GPU parity tests, bench, ncu launch list.  Run under gpurun from the repo root.
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
if [ -z "$SKIP_PROBE" ]; then
echo "== probes" ; 
for tv in "0 0" "0 1" "1 0" "3 0" "3 1" "2 0" "2 1" "2 2" "2 3" "2 4" "2 5"; do
  timeout 30 ./tests/cuda/umma_probe $tv 2>&1 | tail -2
  echo "  (exit $?)"
done | tee gpurun_out/probe.log
fi
echo "== pytest"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
if [ -z "$SKIP_NCU" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
wc -l gpurun_out/launches.csv
fi
