"""GPU parity, op by op, through the C ABI, against the oracle and the reference-generated KATs
(tests/golden/kat_modules.npz, shapes 75 atoms / 37 tokens: nothing is a multiple of a tile)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import physdock_oracle as O
from tests.helpers import T, load_npz, medium_state, rel_close

pytestmark = pytest.mark.gpu

DEV = "cuda"
ATOM_BI, TOK_BI = 1, 3 + 5          # atom_dit_encoder.blocks.1, token_dit.blocks.5 (see make_golden.py)


@pytest.fixture(scope="module")
def env():
    from physdock_b200.dit import B200DiT
    dims, sd, _ = medium_state()
    dit = B200DiT.from_state_dict(sd, dims, device=DEV)
    dit._pack()
    k = {n: T(v).to(DEV) for n, v in load_npz("kat_modules.npz").items() if v.ndim > 0}
    return dims, {n: v.to(DEV) for n, v in sd.items()}, dit, k


def pad_rows(x, S_pad):
    B, S, c = x.shape
    out = torch.zeros(B, S_pad, c, device=x.device, dtype=x.dtype)
    out[:, :S] = x
    return out


# ------------------------------------------------------------------------------------- GEMM / attention core
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 128), (384, 128, 1408), (128, 512, 512), (2048, 256, 192)])
def test_gemm_store_vs_fp64(M, N, K):
    from tests import pdk_ops as ops
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 3).to(DEV)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    ah, al = ops.split_planes(A)
    wh, wl = ops.split_planes(W)
    assert float((ops.planes_to_float(ah, al) - A).abs().max()) <= 3e-7 * float(A.abs().max())
    got = ops.gemm_store(ah, al, wh, wl, bias)
    want = (A.double() @ W.double().t() + bias.double())
    # split-fp16 operand error (2^-22) + fp32 accumulation over K terms
    atol = (2e-6 + 2e-7 * math.sqrt(K)) * float(want.abs().max())
    rel_close("gemm_store", got, want, rtol=0, atol=atol)
    got = ops.gemm_store(ah, al, wh, wl, None, silu=True)
    rel_close("gemm_store+silu", got, F.silu(A.double() @ W.double().t()), rtol=0, atol=atol)


@pytest.mark.parametrize("B,H,S", [(1, 4, 128), (2, 4, 384), (2, 16, 256), (5, 4, 256), (7, 16, 128), (9, 4, 640)])
def test_attention_vs_fp64(B, H, S):
    from tests import pdk_ops as ops
    g = torch.Generator().manual_seed(B * 1000 + S)
    q, k, v = [torch.randn(B, H, S, 32, generator=g).to(DEV) * s for s in (2.0, 2.0, 1.0)]
    bias = (torch.randn(H, S, S, generator=g) * 2).to(DEV)
    bias[:, :, S - 5:] = -1e9       # masked keys
    scale = ops.LOG2E / math.sqrt(32.0)
    planes = [ops.interleave_planes(q * scale), ops.interleave_planes(k), ops.interleave_planes(v)]
    oh, ol = ops.attention(*planes, (bias * ops.LOG2E).contiguous())
    got = ops.planes_to_float(oh, ol).view(B, S, H, 32).transpose(1, 2)
    want = torch.softmax(q.double() @ k.double().transpose(-1, -2) / math.sqrt(32.0) + bias.double(), -1) @ v.double()
    rel_close("attention", got, want, rtol=0, atol=3e-6 * float(want.abs().max()))


def test_attention_many_samples_chunked_work_list():
    """BASELINE.json configs[2]-sized call (Na = 3072, dozens of samples): the work list no longer fits the kernel
    parameters and launch_attention splits the samples into several launches (attention_umma.cu: kMaxWork)."""
    from tests import pdk_ops as ops
    B, H, S = 40, 4, 3072
    g = torch.Generator(device=DEV).manual_seed(77)
    q, k, v = [torch.randn(B, H, S, 32, generator=g, device=DEV) * s for s in (2.0, 2.0, 1.0)]
    bias = torch.randn(H, S, S, generator=g, device=DEV) * 2
    bias[:, :, S - 7:] = -1e9
    scale = ops.LOG2E / math.sqrt(32.0)
    planes = [ops.interleave_planes(q * scale), ops.interleave_planes(k), ops.interleave_planes(v)]
    oh, ol = ops.attention(*planes, (bias * ops.LOG2E).contiguous())
    got = ops.planes_to_float(oh, ol).view(B, S, H, 32).transpose(1, 2)
    for b in (0, 19, 20, 39):         # B = 40 is launched as 20 + 20: samples on both sides of the boundary, fp64 reference per sample
        want = torch.softmax(q[b].double() @ k[b].double().transpose(-1, -2) / math.sqrt(32.0) + bias.double(), -1) @ v[b].double()
        # 3072 keys = 192 sequential fp32 accumulations in TMEM per output (tensor-core accumulation truncates): the
        # error grows with the key count, 5.6e-6 of the scale here against 3e-6 allowed at S <= 640
        rel_close(f"attention sample {b}", got[b], want, rtol=0, atol=1e-5 * float(want.abs().max()))


# ------------------------------------------------------------------------------------- conditioning
def test_time_embed_and_coef(env):
    from tests import pdk_ops as ops
    dims, sd, dit, k = env
    P = dit._packed
    tsilu, coef = ops.time_embed(k["k_t_hat"], P["freq"], P["te_w1"], P["te_b1"], P["te_w2"], P["te_b2"], 16.0)
    rel_close("tsilu", tsilu, F.silu(k["precond_t"]), rtol=0, atol=5e-5)
    t = k["k_t_hat"].double()
    rel_close("c_in", coef[:, 0], 1 / torch.sqrt(t ** 2 + 256), rtol=2e-7, atol=0)
    rel_close("c_skip", coef[:, 1], 256 / (256 + t ** 2), rtol=2e-7, atol=0)
    rel_close("c_out", coef[:, 2], 16 * t / torch.sqrt(256 + t ** 2), rtol=2e-7, atol=0)


def test_mod_gemv_and_adaln(env):
    from tests import pdk_ops as ops
    dims, sd, dit, k = env
    P = dit._packed
    tsilu = F.silu(k["t_emb"])
    mod = ops.mod_gemv(tsilu, P["wmod"], P["bmod"])
    want_mod = tsilu.double() @ P["wmod"].double().t() + P["bmod"].double()
    rel_close("mod", mod, want_mod, rtol=0, atol=2e-6 * float(want_mod.abs().max()))
    off = int(dit._block_array[ATOM_BI].mod_attn_off)
    x = pad_rows(k["ba"], 128)
    hi, lo = ops.adaln(x, mod, off, dims.eps)
    got = ops.planes_to_float(hi, lo)[:, :75]
    rel_close("adaln", got, k["atom_adaln_x"], rtol=0, atol=3e-6 * float(k["atom_adaln_x"].abs().max()))
    rel_close("gate", mod[:, off + 256: off + 384], k["atom_adaln_gate"][:, 0], rtol=0, atol=5e-6)


def test_pair_bias_atom_and_token(env):
    from tests import pdk_ops as ops
    dims, sd, dit, k = env
    P = dit._packed
    for name, pair, mask, wT, bz, stack, bi_local, H in (
            ("atom", k["ap"], k["ap_mask"], P["wz_atom_T"], P["bz_atom"], "atom_dit_encoder", 1, 4),
            ("tok", k["z"], k["z_mask_holes"], P["wz_tok_T"], P["bz_tok"], "token_dit", 5, 16)):
        S = pair.shape[0]
        got = ops.pair_bias(pair, mask, wT, bz, 128)
        want = O.pair_bias(sd, f"{stack}.blocks.{bi_local}.attention.", pair, mask, dims.inf)[0] * ops.LOG2E
        blk = got[bi_local * H:(bi_local + 1) * H]
        rel_close(f"bias[{name}]", blk[:, :S, :S], want, rtol=2e-6, atol=2e-5)
        assert bool((blk[:, :, S:] == -1.0e30).all()), "pad key columns"
        assert bool((blk[:, S:, :S] == 0).all()), "pad query rows"


# ------------------------------------------------------------------------------------- whole sub-blocks
def run_attention_block(dit, dims, k, bi, x, pair, mask, wT, bz, bi_local, S):
    from tests import pdk_ops as ops
    P = dit._packed
    B, _, c = x.shape
    H = c // 32
    mod = ops.mod_gemv(F.silu(k["t_emb"]), P["wmod"], P["bmod"])
    off = int(dit._block_array[bi].mod_attn_off)
    xp = pad_rows(x, 128)
    hi, lo = ops.adaln(xp, mod, off, dims.eps)
    planes = ops.gemm_qkv(hi.view(-1, c), lo.view(-1, c), P[f"b{bi}.wqkv_h"], P[f"b{bi}.wqkv_l"], P[f"b{bi}.norm_q"],
                          P[f"b{bi}.norm_k"], dims.eps, B, 128)
    bias = ops.pair_bias(pair, mask, wT, bz, 128)[bi_local * H:(bi_local + 1) * H].contiguous()
    oh, ol = ops.attention(*planes, bias)
    out = torch.zeros(B * 128, c, device=DEV)
    ops.gemm_gate_resid(oh, ol, P[f"b{bi}.wo_h"], P[f"b{bi}.wo_l"], P[f"b{bi}.bo"], mod[:, off + 2 * c:], mod.shape[1],
                        128, out)
    return out.view(B, 128, c)[:, :S], planes


def test_qkv_epilogue(env):
    """q/k/v planes against the oracle's projections + per-head RMSNorm (attentions.py:248-252)."""
    dims, sd, dit, k = env
    P = dit._packed
    _, planes = run_attention_block(dit, dims, k, ATOM_BI, k["ba"], k["ap"], k["ap_mask"], P["wz_atom_T"], P["bz_atom"], 1, 75)
    from tests import pdk_ops as ops
    p = "atom_dit_encoder.blocks.1.attention."
    xn, _ = O.ada_layer_norm_zero(sd, p + "norm_s.", k["ba"], k["t_emb"], dims.eps)
    want = {}
    for n in "qkv":
        y = F.linear(xn, sd[p + f"linear_{n}.weight"]).reshape(2, 75, 4, 32).transpose(1, 2)
        if n != "v":
            y = O.rms_norm(y, sd[p + f"norm_{n}.weight"], dims.eps)
        want[n] = y * (ops.LOG2E / math.sqrt(32.0) if n == "q" else 1.0)
    for i, n in enumerate("qkv"):
        got = ops.deinterleave_to_float(planes[i])[:, :, :75]
        rel_close(n, got, want[n], rtol=0, atol=3e-6 * float(want[n].abs().max()))


def test_atom_attention_block(env):
    dims, sd, dit, k = env
    P = dit._packed
    got, _ = run_attention_block(dit, dims, k, ATOM_BI, k["ba"], k["ap"], k["ap_mask"], P["wz_atom_T"], P["bz_atom"], 1, 75)
    rel_close("atom_attn", got, k["atom_attn_out"], rtol=0, atol=1e-5 * float(k["atom_attn_out"].abs().max()))


@pytest.mark.parametrize("mask_key,out_key", [("z_mask", "tok_attn_out"), ("z_mask_holes", "tok_attn_holes_out")])
def test_token_attention_block(env, mask_key, out_key):
    dims, sd, dit, k = env
    P = dit._packed
    got, _ = run_attention_block(dit, dims, k, TOK_BI, k["bs"], k["z"], k[mask_key], P["wz_tok_T"], P["bz_tok"], 5, 37)
    want = k[out_key]
    if mask_key == "z_mask_holes":
        # rows of fully masked queries see uniform attention in the reference; compare the live rows
        live = torch.ones(37, dtype=torch.bool, device=DEV)
        live[[3, 17]] = False
        got, want = got[:, live], want[:, live]
    rel_close(out_key, got, want, rtol=0, atol=1e-5 * float(want.abs().max()))


@pytest.mark.parametrize("bi,xk,outk,S", [(ATOM_BI, "ba", "atom_trans_out", 75), (TOK_BI, "bs", "tok_trans_out", 37)])
def test_transition_block(env, bi, xk, outk, S):
    from tests import pdk_ops as ops
    dims, sd, dit, k = env
    P = dit._packed
    x = k[xk]
    B, _, c = x.shape
    mod = ops.mod_gemv(F.silu(k["t_emb"]), P["wmod"], P["bmod"])
    off = int(dit._block_array[bi].mod_ffn_off)
    hi, lo = ops.adaln(pad_rows(x, 128), mod, off, dims.eps)
    hh, hl = ops.gemm_swiglu(hi.view(-1, c), lo.view(-1, c), P[f"b{bi}.w13_h"], P[f"b{bi}.w13_l"])
    out = torch.zeros(B * 128, c, device=DEV)
    ops.gemm_gate_resid(hh, hl, P[f"b{bi}.w2_h"], P[f"b{bi}.w2_l"], None, mod[:, off + 2 * c:], mod.shape[1], 128, out)
    got = out.view(B, 128, c)[:, :S]
    rel_close(outk, got, k[outk], rtol=0, atol=1e-5 * float(k[outk].abs().max()))


def test_fused_atom_transition_vs_oracle_and_unfused(env):
    """transition_umma.cu (one kernel) against the oracle KAT and against the three-kernel path it replaces."""
    from tests import pdk_ops as ops
    dims, sd, dit, k = env
    P = dit._packed
    bi, x = ATOM_BI, k["ba"]
    B, _, c = x.shape
    mod = ops.mod_gemv(F.silu(k["t_emb"]), P["wmod"], P["bmod"])
    off = int(dit._block_array[bi].mod_ffn_off)
    xp = pad_rows(x, 128).contiguous()                       # [B,128,c]; the fused kernel works in place on the residual
    fused = xp.clone().view(-1, c)
    ops.transition_fused(fused, mod, off, P[f"b{bi}.w13_h"], P[f"b{bi}.w13_l"], P[f"b{bi}.w2_h"], P[f"b{bi}.w2_l"], 128, dims.eps)
    hi, lo = ops.adaln(xp, mod, off, dims.eps)
    hh, hl = ops.gemm_swiglu(hi.view(-1, c), lo.view(-1, c), P[f"b{bi}.w13_h"], P[f"b{bi}.w13_l"])
    ref = xp.clone().view(-1, c)
    ops.gemm_gate_resid(hh, hl, P[f"b{bi}.w2_h"], P[f"b{bi}.w2_l"], None, mod[:, off + 2 * c:], mod.shape[1], 128, ref)
    want = x + k["atom_trans_out"]                            # the KAT holds the transition's delta
    S = want.shape[1]
    rel_close("fused transition vs oracle", fused.view(B, 128, c)[:, :S], want, rtol=0, atol=1e-5 * float(want.abs().max()))
    rel_close("fused transition vs unfused kernels", fused, ref, rtol=0, atol=2e-6 * float(ref.abs().max()))
    # several row tiles per CTA (persistent loop, both TMEM/H buffers, A operand rebuilt per tile): 300 tiles on 148 SMs
    g = torch.Generator(device=DEV).manual_seed(5)
    big = torch.randn(300 * 128, c, generator=g, device=DEV)
    modb = torch.randn(300, mod.shape[1], generator=g, device=DEV) * 0.3
    a = big.clone()
    ops.transition_fused(a, modb, off, P[f"b{bi}.w13_h"], P[f"b{bi}.w13_l"], P[f"b{bi}.w2_h"], P[f"b{bi}.w2_l"], 128, dims.eps)
    hi, lo = ops.adaln(big.view(300, 128, c), modb, off, dims.eps)
    hh, hl = ops.gemm_swiglu(hi.view(-1, c), lo.view(-1, c), P[f"b{bi}.w13_h"], P[f"b{bi}.w13_l"])
    b = big.clone()
    ops.gemm_gate_resid(hh, hl, P[f"b{bi}.w2_h"], P[f"b{bi}.w2_l"], None, modb[:, off + 2 * c:], modb.shape[1], 128, b)
    rel_close("fused transition, 300 tiles", a, b, rtol=0, atol=2e-6 * float(b.abs().max()))


# ------------------------------------------------------------------------------------- glue
def test_precond_downscale_upscale_denoise(env):
    from tests import pdk_ops as ops
    dims, sd, dit, k = env
    P = dit._packed
    _, coef = ops.time_embed(k["k_t_hat"], P["freq"], P["te_w1"], P["te_b1"], P["te_w2"], P["te_b2"], 16.0)
    ba = ops.precond(k["k_x_hat"], coef, k["k_a"], P["wx"], P["bx"], 128)
    rel_close("precond", ba[:, :75], k["precond_ba"], rtol=0, atol=2e-6 * float(k["precond_ba"].abs().max()))
    assert bool((ba[:, 75:] == 0).all())
    # downscale
    bap = pad_rows(k["ba"], 128)
    ah, al = ops.split_planes(bap.view(-1, 128))
    h = ops.gemm_store(ah, al, P["wdown_h"], P["wdown_l"], P["bdown"], silu=True).view(2, 128, 512)
    tok_start = torch.cat([torch.zeros(1, dtype=torch.long, device=DEV), torch.cumsum(k["chunk"], 0)]).int()
    bs = ops.segment_mean(h, tok_start, k["k_s"], 128)
    rel_close("downscale", bs[:, :37], k["downscale_out"], rtol=0, atol=5e-6 * float(k["downscale_out"].abs().max()))
    # upscale
    bsp = pad_rows(k["bs"], 128)
    sh, sl = ops.split_planes(bsp.view(-1, 512))
    up = ops.gemm_store(sh, sl, P["wup_h"], P["wup_l"], P["bup"]).view(2, 128, 128)
    ba2 = ops.gather_add(bap.clone(), up, k["a2t"].int().contiguous(), 75)
    rel_close("upscale", ba2[:, :75], k["upscale_out"], rtol=0, atol=3e-6 * float(k["upscale_out"].abs().max()))
    # denoise
    xd = ops.denoise_out(bap, k["k_x_hat"], coef, P["norm_r_w"], P["norm_r_b"], P["wr"], dims.eps)
    rel_close("denoise", xd, k["denoise_out"], rtol=0, atol=3e-6 * float(k["denoise_out"].abs().max()))


def test_fused_adaln_sources_are_bit_identical(env):
    """precond / upscale gather-add fused into the first AdaLN of their atom stack == the stand-alone kernels, bit for bit."""
    from tests import pdk_ops as ops
    dims, sd, dit, k = env
    P = dit._packed
    _, coef = ops.time_embed(k["k_t_hat"], P["freq"], P["te_w1"], P["te_b1"], P["te_w2"], P["te_b2"], 16.0)
    g = torch.Generator(device=DEV).manual_seed(12)
    mod = torch.randn(2, dit._n_mod, generator=g, device=DEV) * 0.3
    off = int(dit._block_array[0].mod_attn_off)
    ba = ops.precond(k["k_x_hat"], coef, k["k_a"], P["wx"], P["bx"], 128)
    hi, lo = ops.adaln(ba, mod, off, dims.eps)
    ba2, hi2, lo2 = ops.precond_adaln(k["k_x_hat"], coef, k["k_a"], P["wx"], P["bx"], 128, mod, off, dims.eps)
    assert torch.equal(ba, ba2) and torch.equal(hi, hi2) and torch.equal(lo, lo2)
    bap = pad_rows(k["ba"], 128)
    up = torch.randn(2, 128, 128, generator=g, device=DEV)
    a2t = k["a2t"].int().contiguous()
    want = ops.gather_add(bap.clone(), up, a2t, 75)
    hi, lo = ops.adaln(want, mod, off, dims.eps)
    got, hi2, lo2 = ops.upscale_adaln(bap.clone(), up, a2t, 75, mod, off, dims.eps)
    assert torch.equal(want, got) and torch.equal(hi, hi2) and torch.equal(lo, lo2)


# ------------------------------------------------------------------------------------- coordinates / physics
def test_centre_random_augmentation(env):
    from physdock_b200 import sampler as S
    _, _, _, k = env
    got = S.centre_augment_noise(k["cra_x"], k["cra_exists"], k["cra_u"], k["cra_trans"])
    rel_close("cra", got, k["cra_out"], rtol=0, atol=2e-6 * float(k["cra_x"].abs().max()))
    # with noise: x_cur + (lambda*noise)*scale, checked against the oracle
    g = torch.Generator().manual_seed(2)
    noise = torch.randn(k["cra_x"].shape, generator=g).to(DEV)
    got = S.centre_augment_noise(k["cra_x"], k["cra_exists"], k["cra_u"], k["cra_trans"], noise, 1.003, 7.25)
    want = k["cra_out"] + (1.003 * noise) * 7.25
    rel_close("cra+noise", got, want, rtol=0, atol=2e-6 * float(k["cra_x"].abs().max()))


def test_euler_update_bitwise(env):
    from physdock_b200 import sampler as S
    g = torch.Generator().manual_seed(8)
    x_hat = torch.randn(3, 75, 3, generator=g) * 4000
    x_den = torch.randn(3, 75, 3, generator=g) * 15
    aligned = torch.randn(3, 75, 3, generator=g) * 15
    w = (torch.rand(75, generator=g) > 0.7).float()
    t_hat = torch.tensor([4608.0, 37.5, 0.4])
    for eta, t_next in ((1.5, 2000.0), (1.0, 0.0)):
        want = O.euler_update(x_hat, (x_hat - x_den) / t_hat[:, None, None], t_hat, torch.tensor(t_next), eta)
        got = S.euler_update(x_hat.to(DEV), x_den.to(DEV), t_hat.to(DEV), t_next, eta)
        assert torch.equal(got.cpu(), want), float((got.cpu() - want).abs().max())
        d = O.physics_direction(x_hat, x_den, aligned, t_hat, w)
        want = O.euler_update(x_hat, d, t_hat, torch.tensor(t_next), eta)
        got = S.euler_update(x_hat.to(DEV), x_den.to(DEV), t_hat.to(DEV), t_next, eta, aligned.to(DEV), w.to(DEV))
        assert torch.equal(got.cpu(), want), float((got.cpu() - want).abs().max())


def test_weighted_rigid_align(env):
    from physdock_b200 import sampler as S
    _, _, _, k = env
    ones = torch.ones(75, device=DEV)
    for gt, out in (("wra_gt", "wra_out"), ("wra_mirror", "wra_out_mirror")):
        got = S.weighted_rigid_align(k["wra_pred"], ones, k[gt], k["wra_w"])
        rel_close(out, got, k[out], rtol=0, atol=2e-5 * float(k[out].abs().max()))
    got = S.weighted_rigid_align(k["wra_pred"], ones, k["wra_gt"][0].contiguous(), k["wra_w"])
    rel_close("wra_shared", got, k["wra_out_shared"], rtol=0, atol=2e-5 * float(k["wra_out_shared"].abs().max()))
    # a genuine rigid motion of the weighted atoms is recovered exactly, planar ligand included
    g = torch.Generator().manual_seed(4)
    Rm = O.rotation_from_uniforms(torch.rand(3, 4, generator=g)).to(DEV)
    base = torch.randn(3, 75, 3, generator=g).to(DEV) * 5
    base[..., 2] = 0.0                                    # planar: third singular value = 0
    moved = torch.einsum("bij,bkj->bki", Rm, base) + 3.0
    got = S.weighted_rigid_align(moved, ones, base, k["wra_w"])
    sel = k["wra_w"].bool()
    rel_close("planar", got[:, sel], moved[:, sel], rtol=0, atol=5e-5)


def test_template_select(env):
    from physdock_b200 import sampler as S
    g = torch.Generator().manual_seed(6)
    B, Na, n, Cn = 3, 75, 22, 9
    x_den = (torch.randn(B, Na, 3, generator=g) * 6)
    lig_idx = torch.arange(40, 40 + n)
    poses = torch.randn(Cn, n, 3, generator=g) * 4
    poses[4] = x_den[1, lig_idx] + 0.05 * torch.randn(n, 3, generator=g)     # sample 1 clearly prefers template 4
    ref_dist = torch.norm(poses[:, :, None] - poses[:, None], dim=-1)
    want_eps = O.template_epsilon(x_den[:, lig_idx], ref_dist)
    brp = torch.zeros(B, Na, 3, device=DEV)
    eps, used = S.template_select(x_den.to(DEV), lig_idx.int().to(DEV), ref_dist.to(DEV).contiguous(),
                                  poses.to(DEV).contiguous(), brp)
    rel_close("eps", eps, want_eps, rtol=0, atol=2e-6)
    assert torch.equal(used.cpu(), torch.argmin(want_eps, -1))
    assert int(used[1]) == 4
    want_brp = torch.zeros(B, Na, 3)
    want_brp[:, lig_idx] = poses[torch.argmin(want_eps, -1)]
    assert torch.equal(brp.cpu(), want_brp)
