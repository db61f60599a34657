"""Graph-timed atom- and token-shaped attention launches (us per launch, best of 5 replays of 12 launches)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import pdk_ops as ops
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def run(B, H, S, n=12):
    q, k, v = [torch.randn(B, H, S, 32, generator=g, device=dev) for _ in range(3)]
    planes = [ops.interleave_planes(q * 0.25), ops.interleave_planes(k), ops.interleave_planes(v)]
    bias = torch.randn(3, H, S, S, generator=g, device=dev)
    for i in range(3): ops.attention(*planes, bias[i % 3])
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(n): ops.attention(*planes, bias[i % 3])
    gr.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best
print(os.environ.get("PHYSDOCK_B200_LIB", "product build"))
print(f"  atom  attention B=16 H=4  S=2048: {run(16, 4, 2048):7.1f} us")
print(f"  token attention B=16 H=16 S=256 : {run(16, 16, 256):7.1f} us")
