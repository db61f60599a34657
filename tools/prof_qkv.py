"""Runs the atom- and token-shaped QKV GEMMs a few times (for ncu captures)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import pdk_ops as ops
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def planes(r, c): return ops.split_planes(torch.randn(r, c, generator=g, device=dev))
xa, wa = planes(32768, 128), planes(384, 128)
xt, wt = planes(4096, 512), planes(1536, 512)
nq = nk = torch.ones(32, device=dev)
for i in range(3):
    ops.gemm_qkv(*xa, *wa, nq, nk, 1e-8, 16, 2048)
    ops.gemm_qkv(*xt, *wt, nq, nk, 1e-8, 16, 256)
torch.cuda.synchronize(); print("done")
