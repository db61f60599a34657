"""Runs the fused atom transition kernel at the benchmark shape (B=16, Na=2048) a few times (for ncu captures)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import pdk_ops as ops
dev = torch.device("cuda"); g = torch.Generator(device=dev).manual_seed(0)
def planes(r, c): return ops.split_planes(torch.randn(r, c, generator=g, device=dev))
x = torch.randn(32768, 128, generator=g, device=dev); mod = torch.randn(16, 384, device=dev) * 0.1
w13, w2 = planes(768, 128), planes(128, 384)
for i in range(3): ops.transition_fused(x, mod, 0, *w13, *w2, 2048, 1e-8)
torch.cuda.synchronize(); print("done")
