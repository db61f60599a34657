"""Runs the token SwiGLU-shaped and atom QKV-shaped GEMMs a few times (for ncu captures)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import pdk_ops as ops
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def planes(r, c): 
    x = torch.randn(r, c, generator=g, device=dev)
    return ops.split_planes(x)
ah, al = planes(4096, 512); wh, wl = planes(2816, 512)
bh, bl = planes(32768, 128); vh, vl = planes(768, 128)
for i in range(3):
    ops.gemm_swiglu(ah, al, wh, wl)
    ops.gemm_swiglu(bh, bl, vh, vl)
torch.cuda.synchronize(); print("done")
