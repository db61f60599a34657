#!/bin/bash
mkdir -p gpurun_out
{
echo "== tests"
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -15
echo "== bench (balanced)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bal.json | cut -c1-300
echo "== bench (PDK_NO_ATTN_BALANCE)"
PDK_NO_ATTN_BALANCE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
} 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/attn.log
grep -o '"roofline.*' gpurun_out/bench_bal.json
