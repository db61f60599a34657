"""Per-unit timeline of one CTA of the attention kernel (debug).  Prints cycle stamps relative to the first."""
import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from physdock_b200 import _lib
from tests import pdk_ops as ops
B, H, S = 16, 4, 2048
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
q, k, v = [torch.randn(B, H, S, 32, generator=g, device=dev) for _ in range(3)]
planes = [ops.interleave_planes(q * 0.25), ops.interleave_planes(k), ops.interleave_planes(v)]
bias = torch.randn(H, S, S, generator=g, device=dev)
ops.attention(*planes, bias); torch.cuda.synchronize()
lib = _lib.load()
U = (S // 64) * 4
trace = torch.zeros(U * 8, dtype=torch.int64, device=dev)
lib.pdk_debug_attention_trace.argtypes = [ctypes.c_void_p]
lib.pdk_debug_attention_trace(trace.data_ptr())
ops.attention(*planes, bias); torch.cuda.synchronize()
lib.pdk_debug_attention_trace(None)
t = trace.view(U, 8).cpu()
t0 = int(t[:, 3][t[:, 3] > 0].min())
names = ["mma:p_ready", "mma:pv_issued", "mma:qk_issued", "sm:wait_S", "sm:got_S", "sm:max_done", "sm:P_pub"]
print("unit j g | " + " ".join(f"{n:>13s}" for n in names))
for u in list(range(0, 24)) + list(range(64, 76)):
    r = t[u]
    print(f"{u:4d} {u//4:2d} {u%4} | " + " ".join(f"{int(x)-t0:13d}" if int(x) > 0 else f"{'-':>13s}" for x in r[:7]))
# steady-state statistics over the middle units
import statistics as st
mid = range(16, U - 8)
def diff(a, b): return [int(t[u][b]) - int(t[u][a]) for u in mid]
print("softmax: wait for S      ", st.mean(diff(3, 4)))
print("softmax: ld+bias+max     ", st.mean(diff(4, 5)))
print("softmax: exp+split+st    ", st.mean(diff(5, 6)))
print("pv warp: p_ready->pv issued  ", st.mean(diff(0, 1)))
print("P_pub -> mma saw p_ready ", st.mean([int(t[u][0]) - int(t[u][6]) for u in mid]))
print("qk issued(u) -> softmax got S(u)", st.mean([int(t[u][4]) - int(t[u][2]) for u in mid]))
print("per-unit period (qk warp) ", (int(t[U - 9][2]) - int(t[16][2])) / (U - 9 - 16))
print("per-unit period (mma)    ", (int(t[U - 9][0]) - int(t[16][0])) / (U - 9 - 16))
