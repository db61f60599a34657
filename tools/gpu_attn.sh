#!/bin/bash
mkdir -p gpurun_out
{
echo "== attention tests"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -x -k "attention" 2>&1 | tail -15
echo "== all gpu tests"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -15
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_attn.json | cut -c1-330
grep -o '"roofline.*' gpurun_out/bench_attn.json
echo "== bench (PDK_NO_ATTN_BALANCE)"
PDK_NO_ATTN_BALANCE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step[^,]*\|launch_ms[^,]*'
echo "== trace"
timeout 120 python tools/trace_attention.py 2>&1 | tail -12
} 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/attn.log
