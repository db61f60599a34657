#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step[^,]*' | head -1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:precond|segment_mean|gather_add|time_embed|denoise_out|split_kernel" -c 14 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "gpu__time" | awk -F'","' '{print $5, $NF}' | tail -8
} 2>&1 | tee gpurun_out/t.log
