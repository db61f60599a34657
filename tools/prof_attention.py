"""Runs the atom-shaped attention kernel a few times (for ncu captures): B=16, H=4, S=2048."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import pdk_ops as ops
B, H, S = int(os.environ.get("B", 16)), int(os.environ.get("H", 4)), int(os.environ.get("S", 2048))
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = [torch.randn(B, H, S, 32, generator=g, device=dev) for _ in range(3)]
planes = [ops.interleave_planes(q * 0.25), ops.interleave_planes(k), ops.interleave_planes(v)]
bias = torch.randn(3, H, S, S, generator=g, device=dev)
for i in range(int(os.environ.get("N", 4))):
    ops.attention(*planes, bias[i % 3])
torch.cuda.synchronize()
print("done")
