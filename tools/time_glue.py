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
Sa, St, Nt = 2048, 256, 256
a = torch.randn(Na, 128, generator=g, device=dev)
wx, bx = torch.randn(128, 3, generator=g, device=dev), torch.randn(128, generator=g, device=dev)
mod = torch.randn(B, 384, generator=g, device=dev) * 0.1
print(f"  precond + first AdaLN    : {timeit(lambda: ops.precond_adaln(x_hat, coef, a, wx, bx, Sa, mod, 0, 1e-8)):6.2f} us   (writes 50 MB)")
up = torch.randn(B, St, 128, generator=g, device=dev)
a2t = torch.arange(Na, device=dev, dtype=torch.int32) // 8
bab = ba.clone()
print(f"  upscale gather-add + AdaLN: {timeit(lambda: ops.upscale_adaln(bab, up, a2t, Na, mod, 0, 1e-8)):6.2f} us   (reads 19 MB, writes 50 MB)")
h = torch.randn(B, Sa, 512, generator=g, device=dev)
tok_start = torch.cat([torch.arange(0, 2016, 9), torch.arange(2016, 2049)]).int().to(dev)
s = torch.randn(Nt, 512, generator=g, device=dev)
print(f"  segment_mean             : {timeit(lambda: ops.segment_mean(h, tok_start, s, St)):6.2f} us   (reads 67 MB)")
print(f"  split (atom stream)      : {timeit(lambda: ops.split_planes(ba.view(-1, 128))):6.2f} us   (reads 16.8 MB, writes 16.8 MB)")
bs = torch.randn(B * St, 512, generator=g, device=dev)
print(f"  split (token stream)     : {timeit(lambda: ops.split_planes(bs)):6.2f} us")
