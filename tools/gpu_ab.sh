#!/bin/bash
mkdir -p gpurun_out
{
for i in 1 2; do
timeout 200 python tools/time_attention.py
PHYSDOCK_B200_LIB=/root/repo/build/dbg/libpdk_PREV.so timeout 200 python tools/time_attention.py
done
} 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/ab.log
