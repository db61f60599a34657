#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/time_gemm.py 2>&1 | grep -v "^\[W\|Warning" | tee gpurun_out/time_gemm.log
cat > /tmp/prof_trans.py <<'PY'
import sys, os, torch
sys.path.insert(0, "/root/repo")
from physdock_b200 import ops
dev = torch.device("cuda"); g = torch.Generator(device=dev).manual_seed(0)
def planes(r, c): return ops.split_planes(torch.randn(r, c, generator=g, device=dev))
x = torch.randn(32768, 128, generator=g, device=dev); mod = torch.randn(16, 384, device=dev) * 0.1
w13, w2 = planes(768, 128), planes(128, 384)
for i in range(3): ops.transition_fused(x, mod, 0, *w13, *w2, 2048, 1e-8)
torch.cuda.synchronize(); print("done")
PY
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:transition_umma -s 1 -c 1 -o gpurun_out/transition python /tmp/prof_trans.py > gpurun_out/ncu_trans.log 2>&1
tail -2 gpurun_out/ncu_trans.log
