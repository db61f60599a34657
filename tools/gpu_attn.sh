#!/bin/bash
mkdir -p gpurun_out
{
echo "== attention tests"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -x -k "attention" 2>&1 | tail -8
echo "== all gpu tests"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -5
for i in 1 2; do
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_attn.json | grep -o '"ms_per_step[^,]*\|launch_ms[^,]*' | head -3
done
} 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/attn.log
