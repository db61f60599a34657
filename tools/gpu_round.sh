#!/bin/bash
# One full GPU-box visit: GPU parity tests, smoke, bench (ours, physics variant, reference arm).  Run under gpurun.
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest"
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps ${BENCH_STEPS:-40} --warmup 3 2> gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json | cut -c1-400
tail -5 gpurun_out/bench.err
if [ -n "$WITH_PHYSICS" ]; then
echo "== bench --physics"
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --physics 2>&1 | tail -1 | tee gpurun_out/bench_physics.json | cut -c1-200
fi
if [ -n "$WITH_REFERENCE" ]; then
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-300
fi
