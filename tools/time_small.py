"""Launch-latency floor at small batches: graph-timed chains of the step's kernels at the C1 shape (4 samples, 64 tokens /
512 atoms: Mt = 512 token rows, Ma = 2048 atom rows), us per launch back to back with PDL, as the step runs them."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import pdk_ops as ops
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def planes(r, c):
    return ops.split_planes(torch.randn(r, c, generator=g, device=dev))
def timeit(fn, n=40, reps=5):
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
Sa, St = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 128)
Ma, Mt = B * Sa, B * St
xa, xt = planes(Ma, 128), planes(Mt, 512)
ha, ht = planes(Ma, 384), planes(Mt, 1408)
W = {"a13": planes(768, 128), "a2": planes(128, 384), "adown": planes(512, 128), "t13": planes(2816, 512),
     "t2": planes(512, 1408), "to": planes(512, 512), "aqkv": planes(384, 128), "ao": planes(128, 128), "tqkv": planes(1536, 512)}
nq, nk = torch.ones(32, device=dev), torch.ones(32, device=dev)
gate_a, gate_t = torch.randn(B, 128, device=dev), torch.randn(B, 512, device=dev)
out_a, out_t = torch.zeros(Ma, 128, device=dev), torch.zeros(Mt, 512, device=dev)
mod_a, mod_t = torch.randn(B, 384, device=dev) * 0.1, torch.randn(B, 1536, device=dev) * 0.1
def attn(H, S):
    q, k, v = [torch.randn(B, H, S, 32, generator=g, device=dev) for _ in range(3)]
    pl = [ops.interleave_planes(q * 0.25), ops.interleave_planes(k), ops.interleave_planes(v)]
    bias = torch.randn(H, S, S, generator=g, device=dev)
    return lambda: ops.attention(*pl, bias)
shapes = {
  f"atom adaln        {Ma}x128": lambda: ops.adaln(out_a.view(B, Sa, 128), mod_a, 0, 1e-8),
  f"token adaln       {Mt}x512": lambda: ops.adaln(out_t.view(B, St, 512), mod_t, 0, 1e-8),
  f"atom qkv          {Ma}x384x128": lambda: ops.gemm_qkv(*xa, *W["aqkv"], nq, nk, 1e-8, B, Sa),
  f"atom out          {Ma}x128x128": lambda: ops.gemm_gate_resid(*xa, *W["ao"], None, gate_a, 128, Sa, out_a),
  f"atom fused transition": lambda: ops.transition_fused(out_a, mod_a, 0, *W["a13"], *W["a2"], Sa, 1e-8),
  f"atom attention    H=4  S={Sa}": attn(4, Sa),
  f"tok  qkv          {Mt}x1536x512": lambda: ops.gemm_qkv(*xt, *W["tqkv"], nq, nk, 1e-8, B, St),
  f"tok  attention    H=16 S={St}": attn(16, St),
  f"tok  out          {Mt}x512x512": lambda: ops.gemm_gate_resid(*xt, *W["to"], None, gate_t, 512, St, out_t),
  f"tok  swiglu       {Mt}x2816x512": lambda: ops.gemm_swiglu(*xt, *W["t13"]),
  f"tok  w2           {Mt}x512x1408": lambda: ops.gemm_gate_resid(*ht, *W["t2"], None, gate_t, 512, St, out_t),
}
print(f"B={B} Sa={Sa} St={St}", os.environ.get("PHYSDOCK_B200_LIB", "product build"))
tot = 0.0
for k, fn in shapes.items():
    t = timeit(fn); print(f"  {k:40s}: {t:7.2f} us")
