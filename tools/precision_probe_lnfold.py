"""Would folding the token stack's AdaLayerNormZero into its consumer GEMM hold the parity budget?  (CPU, emulation in the oracle)
   y = LN(x) * (1 + sc) + sh;  y W^T  ==  rstd * ( (x * (1 + sc)) W^T  -  mu * c1 ) + c2,   c1 = (1 + sc) W^T,  c2 = sh W^T
with (x * (1 + sc)) split into fp16 hi/lo planes by the PRODUCER GEMM's epilogue, row moments (sum x, sum x^2) from its partial
sums, and the correction applied in the consumer GEMM's epilogue.  Prints the RMSD of x_denoised vs an fp64 evaluation."""
import sys, torch
sys.path.insert(0, '.')
from oracle import physdock_oracle as O
from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex
import torch.nn.functional as F
torch.set_num_threads(8)

def split16(x):
    hi = x.half().float()
    return hi, (x - hi).half().float()
def lin3(x, w):            # split-fp16, three products, fp32 accumulation
    xh, xl = split16(x); wh, wl = split16(w)
    return F.linear(xl, wh) + F.linear(xh, wl) + F.linear(xh, wh)

MODE = {"fold": False}
def folded(sd, pfx, x, t, eps, weights):
    """AdaLN-Zero (modulation from sd[pfx + 'linear.*']) followed by the bias-free linears `weights`, folded."""
    shift, scale, gate = F.linear(F.silu(t[..., None, :]), sd[pfx + "linear.weight"], sd[pfx + "linear.bias"]).chunk(3, dim=-1)
    c = x.shape[-1]
    if not MODE["fold"] or c != 512:
        xn = F.layer_norm(x, (c,), None, None, eps) * (1 + scale) + shift
        return [lin3(xn, w) for w in weights], gate
    xs = x * (1 + scale)
    mu = x.sum(-1, keepdim=True) / c
    var = (x * x).sum(-1, keepdim=True) / c - mu * mu
    rstd = torch.rsqrt(var + eps)
    outs = []
    for w in weights:
        acc = lin3(xs, w)
        c1 = lin3((1 + scale).expand(x.shape[0], 1, c), w)
        c2 = lin3(shift.expand(x.shape[0], 1, c), w)
        outs.append(rstd * (acc - mu * c1) + c2)
    return outs, gate

def dit_attention(sd, p, bs, z, t, z_mask, inf, eps, bias=None):
    B, S, c = bs.shape
    D, H = 32, c // 32
    (q, k, v), gate = folded(sd, p + "norm_s.", bs, t, eps, [sd[p + "linear_q.weight"], sd[p + "linear_k.weight"], sd[p + "linear_v.weight"]])
    q, k, v = [u.reshape([B, S, H, D]).transpose(-2, -3) for u in (q, k, v)]
    q = O.rms_norm(q, sd[p + "norm_q.weight"], eps)
    k = O.rms_norm(k, sd[p + "norm_k.weight"], eps)
    if bias is None:
        bias = O.pair_bias(sd, p, z, z_mask, inf)
    o = F.scaled_dot_product_attention(q, k, v, bias.to(q.dtype), dropout_p=0, scale=None).transpose(-2, -3).reshape([B, S, -1])
    return (lin3(o, sd[p + "linear_o.weight"]) + sd[p + "linear_o.bias"]) * gate

def dit_transition(sd, p, x, t, eps):
    f = p + "feed_forward."
    (h1, h3), gate = folded(sd, p + "ffn_norm.", x, t, eps, [sd[f + "w1.weight"], sd[f + "w3.weight"]])
    return lin3(F.silu(h1) * h3, sd[f + "w2.weight"]) * gate

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
orig = (O.dit_attention, O.dit_transition)
O.dit_attention, O.dit_transition = dit_attention, dit_transition
print("variant                         | rmsd(A) vs fp64 at t=4608,100,10,1,0.2")
for fold in (False, True):
    MODE["fold"] = fold
    out = []
    for t, x_hat, t_hat, y64 in cases:
        with torch.no_grad():
            y = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        out.append(float(O.rmsd(y, y64).max()))
    print(f"{'token AdaLN folded into GEMM' if fold else 'split-fp16 linears (as shipped)':31s} | " + "  ".join(f"{e:.2e}" for e in out), flush=True)
O.dit_attention, O.dit_transition = orig
