"""CUDA-event timing of the GEMM shapes of one atom block and one token block (B=16): us per launch."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import pdk_ops as ops
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def planes(r, c):
    return ops.split_planes(torch.randn(r, c, generator=g, device=dev))
def timeit(fn, n=20, reps=5):
    """n launches captured in one CUDA graph (no CPU launch cost in the timed region), best of `reps` replays."""
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
Ma, Mt = 32768, 4096
xa, xt = planes(Ma, 128), planes(Mt, 512)
ha, ht = planes(Ma, 384), planes(Mt, 1408)
shapes = {
  "atom swiglu 32768x768x128": lambda: ops.gemm_swiglu(*xa, *W["a13"]),
  "atom w2     32768x128x384": lambda: ops.gemm_gate_resid(*ha, *W["a2"], None, gate_a, 128, 2048, out_a),
  "atom store  32768x512x128": lambda: ops.gemm_store(*xa, *W["adown"]),
  "atom qkv    32768x384x128": lambda: ops.gemm_qkv(*xa, *W["aqkv"], nq, nk, 1e-8, 16, 2048),
  "atom out    32768x128x128": lambda: ops.gemm_gate_resid(*xa, *W["ao"], None, gate_a, 128, 2048, out_a),
  "atom fused transition    ": lambda: ops.transition_fused(out_a, mod_a, 0, *W["a13"], *W["a2"], 2048, 1e-8),
  "atom adaln               ": lambda: ops.adaln(out_a.view(16, 2048, 128), mod_a, 0, 1e-8),
  "tok  qkv    4096x1536x512": lambda: ops.gemm_qkv(*xt, *W["tqkv"], nq, nk, 1e-8, 16, 256),
  "tok  swiglu 4096x2816x512": lambda: ops.gemm_swiglu(*xt, *W["t13"]),
  "tok  w2     4096x512x1408": lambda: ops.gemm_gate_resid(*ht, *W["t2"], None, gate_t, 512, 256, out_t),
  "tok  out    4096x512x512 ": lambda: ops.gemm_gate_resid(*xt, *W["to"], None, gate_t, 512, 256, out_t),
}
W = {"a13": planes(768, 128), "a2": planes(128, 384), "adown": planes(512, 128), "t13": planes(2816, 512),
     "t2": planes(512, 1408), "to": planes(512, 512), "aqkv": planes(384, 128), "ao": planes(128, 128), "tqkv": planes(1536, 512)}
nq, nk = torch.ones(32, device=dev), torch.ones(32, device=dev)
gate_a, gate_t = torch.randn(16, 128, device=dev), torch.randn(16, 512, device=dev)
out_a, out_t = torch.zeros(Ma, 128, device=dev), torch.zeros(Mt, 512, device=dev)
mod_a = torch.randn(16, 384, device=dev) * 0.1
print(os.environ.get("PHYSDOCK_B200_LIB", "product build"))
for k, fn in shapes.items():
    print(f"  {k}: {timeit(fn):7.1f} us")
