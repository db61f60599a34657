"""Graph-timed glue kernels at the benchmark shape (B=16, Na=2048): us per launch."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import pdk_ops as ops
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def timeit(fn, n=20, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best
B, Na = 16, 2048
ba = torch.randn(B, Na, 128, generator=g, device=dev)
x_hat = torch.randn(B, Na, 3, generator=g, device=dev)
coef = torch.rand(B, 8, generator=g, device=dev) + 0.5
ln_w, ln_b = torch.randn(128, generator=g, device=dev), torch.randn(128, generator=g, device=dev)
wr = torch.randn(3, 128, generator=g, device=dev)
print(os.environ.get("PHYSDOCK_B200_LIB", "product build"))
print(f"  denoise_out B=16 Na=2048: {timeit(lambda: ops.denoise_out(ba, x_hat, coef, ln_w, ln_b, wr, 1e-8)):6.2f} us")
