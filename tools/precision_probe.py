"""Emulates candidate tensor-core number formats inside the oracle to size the parity budget (CPU)."""
import sys, types, torch
sys.path.insert(0, '.')
from oracle import physdock_oracle as O
from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex
import torch.nn.functional as F
torch.set_num_threads(8)

def tf32_rna(x):
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)
def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
def bf16(x): return x.bfloat16().float()
def fp16(x): return x.half().float()

def split(x, r, n):
    parts, rem = [], x
    for _ in range(n):
        p = r(rem); parts.append(p); rem = rem - p
    return parts

def make_linear(r, n, terms):
    def lin(x, w, b=None):
        xs, ws = split(x, r, n), split(w, r, n)
        y = None
        for (i, j) in reversed(terms):
            p = F.linear(xs[i], ws[j])
            y = p if y is None else y + p
        return y if b is None else y + b
    return lin

LIN = {
 "fp32": F.linear,
 "tf32x1": make_linear(tf32_rna, 1, [(0,0)]),
 "bf16x2(3)": make_linear(bf16, 2, [(0,0),(0,1),(1,0)]),
 "fp16x2(3)": make_linear(fp16, 2, [(0,0),(0,1),(1,0)]),
 "tf32x2(3)": make_linear(tf32_rna, 2, [(0,0),(0,1),(1,0)]),
 "tf32trunc x2(3)": make_linear(tf32_trunc, 2, [(0,0),(0,1),(1,0)]),
 "bf16x3(6)": make_linear(bf16, 3, [(0,0),(0,1),(1,0),(1,1),(0,2),(2,0)]),
}

def make_sdpa(rq, nq, qterms, rp, np_, pterms):
    def sdpa(q, k, v, bias, dropout_p=0, scale=None):
        D = q.shape[-1]
        qs, ks = split(q, rq, nq), split(k, rq, nq)
        s = None
        for (i, j) in reversed(qterms):
            p = qs[i] @ ks[j].transpose(-1, -2)
            s = p if s is None else s + p
        s = s * (D ** -0.5) + bias
        m = s.max(-1, keepdim=True).values
        p = torch.exp(s - m)
        l = p.sum(-1, keepdim=True)
        ps, vs = split(p, rp, np_), split(v, rp, np_)
        o = None
        for (i, j) in reversed(pterms):
            t = ps[i] @ vs[j]
            o = t if o is None else o + t
        return o / l
    return sdpa

one = [(0,0)]; three = [(0,0),(0,1),(1,0)]
SDPA = {
 "fp32": F.scaled_dot_product_attention,
 "tf32 QK,PV": make_sdpa(tf32_rna,1,one, tf32_rna,1,one),
 "tf32trunc QK,PV": make_sdpa(tf32_trunc,1,one, tf32_trunc,1,one),
 "bf16 QK,PV": make_sdpa(bf16,1,one, bf16,1,one),
 "bf16x2 QK, bf16x2 PV": make_sdpa(bf16,2,three, bf16,2,three),
 "bf16x2 QK, bf16 PV": make_sdpa(bf16,2,three, bf16,1,one),
 "tf32x2(3) QK, tf32 PV": make_sdpa(tf32_rna,2,three, tf32_rna,1,one),
 "fp16x2 QK, fp16 PV": make_sdpa(fp16,2,three, fp16,1,one),
 "fp16 QK,PV": make_sdpa(fp16,1,one, fp16,1,one),
}

class Shim:
    def __init__(self, lin, sdpa):
        self.linear, self.scaled_dot_product_attention = lin, sdpa
    def __getattr__(self, k): return getattr(F, k)

dims = DiTDims.named("medium")
sd = make_dit_state(dims, seed=0)
Nt, Na, B = (int(sys.argv[1]), int(sys.argv[2]), 2) if len(sys.argv) > 2 else (64, 512, 4)
cx = make_complex(Nt, Na, dims, seed=1)
sd64 = {k: v.double() for k, v in sd.items()}
cx64 = {k: (v.double() if v.is_floating_point() else v) for k, v in cx.items()}
g = torch.Generator().manual_seed(3)
cases = []
for t in [4608.0, 100.0, 10.0, 1.0, 0.2]:
    x_hat = torch.randn(B, Na, 3, generator=g) * (t**2 + 100)**0.5
    t_hat = torch.full([B], t)
    with torch.no_grad():
        y64 = O.af3dit_forward(sd64, cx64, x_hat.double(), t_hat.double(), cx64["a"], cx64["ap"], cx64["s"], cx64["z"])
    cases.append((t, x_hat, t_hat, y64))

def run(ln, an):
    O.F = Shim(LIN[ln], SDPA[an])
    out = []
    for t, x_hat, t_hat, y64 in cases:
        with torch.no_grad():
            y = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        out.append(float(O.rmsd(y, y64).max()))
    O.F = F
    print(f"{ln:18s} | {an:24s} | " + "  ".join(f"{e:.2e}" for e in out), flush=True)

SDPA["fp16x2 QK, fp16x2 PV"] = make_sdpa(fp16,2,three, fp16,2,three)
SDPA["tf32x2 QK, tf32x2 PV"] = make_sdpa(tf32_rna,2,three, tf32_rna,2,three)
print("linear             | attention                | rmsd(A) vs fp64 at t=4608,100,10,1,0.2")
if len(sys.argv) > 3 and sys.argv[3] == "full":
    for ln in LIN: run(ln, "fp32")
    for an in SDPA: run("fp32", an)
run("fp32", "fp32")
run("fp32", "fp16x2 QK, fp16x2 PV")
run("fp16x2(3)", "fp16x2 QK, fp16x2 PV")
run("bf16x2(3)", "bf16x2 QK, bf16x2 PV")
run("fp32", "tf32x2 QK, tf32x2 PV")
