#!/bin/bash
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
for v in PARTIAL "" PARTIAL ""; do
  if [ -z "$v" ]; then echo "### product"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step[^,]*\|"clocks[^}]*' | head -2
  else echo "### $v"; PHYSDOCK_B200_LIB=/root/repo/build/dbg/libpdk_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step[^,]*\|"clocks[^}]*' | head -2; fi
done
} 2>&1 | tee gpurun_out/ab.log
