"""Generates tests/golden/*.npz FROM THE REAL REFERENCE (imported from /root/reference through
oracle/ref_import.py).  TEST INFRASTRUCTURE.  Run here (the reference tree does not exist on the GPU
box):   python -m oracle.make_golden

The reference ships no golden vectors, so these are the pins:
  kat_modules.npz   per-module known-answer tests at small, tile-unfriendly shapes (inputs stored)
  dit_64_512.npz    full AF3DiT outputs, Nt=64 Na=512 B=4, t_hat in {4608,100,10,1,0.2}; inputs and
                    weights are regenerated from seeds (physdock_b200.synthetic) and guarded by checksums
  trace_*.npz       12-step PhysDock.sample_diffusion traces (x_hat, t_hat, x_denoised, x_next per step
                    + the recorded RNG tape) for: no physics / ref_pos alignment / template selection
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_import import build_reference_dit, import_reference  # noqa: E402
from physdock_b200.synthetic import (DiTDims, make_dit_state, make_complex, make_templates,  # noqa: E402
                                     checksum)

OUT = os.path.join(ROOT, "tests", "golden")
T_LEVELS = [4608.0, 100.0, 10.0, 1.0, 0.2]
TRACE_STEPS = 12
TRACE_B = 2


def npy(t):
    return t.detach().cpu().numpy()


def record_tape(fn, seed):
    """Runs fn() under torch.manual_seed(seed) recording every torch.rand / torch.normal draw."""
    tape = []
    orig_rand, orig_normal = torch.rand, torch.normal

    def rand(*a, **k):
        t = orig_rand(*a, **k)
        tape.append(t.detach().clone())
        return t

    def normal(*a, **k):
        t = orig_normal(*a, **k)
        tape.append(t.detach().clone())
        return t

    torch.manual_seed(seed)
    torch.rand, torch.normal = rand, normal
    try:
        out = fn()
    finally:
        torch.rand, torch.normal = orig_rand, orig_normal
    return out, tape


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    dims = DiTDims.named("medium")
    PhysDock, _, _, tu = import_reference()
    dit = build_reference_dit("medium")
    sd = make_dit_state(dims, seed=0)
    dit.load_state_dict(sd)
    sd_sum = {"sd_checksum": checksum(torch.cat([v.flatten() for v in sd.values()]))}

    # ------------------------------------------------------------------ (i) module KATs
    g = torch.Generator().manual_seed(11)
    kat = {}
    B = 2
    Sa, St = 75, 37                     # deliberately not multiples of any tile size
    t_emb = torch.randn(B, 256, generator=g)
    ba = torch.randn(B, Sa, dims.c_a, generator=g) * 1.5
    ap = torch.randn(Sa, Sa, dims.c_ap, generator=g)
    ap_mask = torch.ones(Sa, Sa)
    bs = torch.randn(B, St, dims.c_s, generator=g) * 1.5
    zz = torch.randn(St, St, dims.c_z, generator=g)
    z_mask = torch.ones(St, St)
    hole = torch.ones(St)
    hole[[3, 17]] = 0
    z_mask_holes = hole[:, None] * hole[None, :]
    with torch.no_grad():
        ab = dit.atom_dit_encoder.blocks[1]
        tb = dit.token_dit.blocks[5]
        kat.update(t_emb=t_emb, ba=ba, ap=ap, ap_mask=ap_mask, bs=bs, z=zz, z_mask=z_mask,
                   z_mask_holes=z_mask_holes)
        kat["atom_attn_out"] = ab.attention(ba, ap, t_emb, ap_mask, None)
        kat["atom_trans_out"] = ab.transition(ba, t_emb)
        kat["atom_block_out"] = ab(ba, ap, t_emb, ap_mask, None)
        kat["tok_attn_out"] = tb.attention(bs, zz, t_emb, z_mask, None)
        kat["tok_attn_holes_out"] = tb.attention(bs, zz, t_emb, z_mask_holes, None)
        kat["tok_trans_out"] = tb.transition(bs, t_emb)
        kat["tok_block_out"] = tb(bs, zz, t_emb, z_mask, None)
        xn, gate = ab.attention.norm_s(ba, t_emb)
        kat["atom_adaln_x"], kat["atom_adaln_gate"] = xn, gate
        # time embedding / precond / denoise / downscale / upscale
        t_hat = torch.tensor([4608.0, 0.37])
        x_hat = torch.randn(B, Sa, 3, generator=g) * torch.sqrt(t_hat ** 2 + 100)[:, None, None]
        a_in = torch.randn(Sa, dims.c_a, generator=g)
        ba0, tt, _ = dit.precond(x_hat, t_hat, a_in)
        kat.update(k_t_hat=t_hat, k_x_hat=x_hat, k_a=a_in, precond_ba=ba0, precond_t=tt)
        kat["denoise_out"] = dit.denoise(x_hat, t_hat, ba)
        chunk = torch.tensor([3, 0, 5, 1, 9, 2, 0, 7, 4, 6, 1, 1, 8, 3, 2, 5, 1, 1, 4, 2, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1])
        assert chunk.numel() == St
        chunk[4] += Sa - int(chunk.sum())
        assert int(chunk.sum()) == Sa and int(chunk.min()) >= 0
        a2t = torch.repeat_interleave(torch.arange(St), chunk)
        s_in = torch.randn(St, dims.c_s, generator=g)
        kat.update(chunk=chunk, a2t=a2t, k_s=s_in)
        kat["downscale_out"] = dit.downscale(ba, s_in, chunk)
        kat["upscale_out"] = dit.upscale(ba, bs, a2t)
        # coordinate ops (randoms recorded from the reference's own draws)
        x_in = torch.randn(3, Sa, 3, generator=g) * 40 + 7
        x_exists = torch.ones(Sa)
        x_exists[[0, 9, 33]] = 0
        out, tape = record_tape(lambda: tu.centre_random_augmentation(x_in, x_exists), seed=5)
        assert len(tape) == 5
        kat.update(cra_x=x_in, cra_exists=x_exists, cra_u=torch.stack(tape[:4], -1), cra_trans=tape[4],
                   cra_out=out)
        # Kabsch + template selection
        w = torch.zeros(Sa)
        w[40:62] = 1
        x_pred = torch.randn(3, Sa, 3, generator=g) * 8
        x_ref3 = torch.randn(3, Sa, 3, generator=g) * 8
        kat.update(wra_pred=x_pred, wra_gt=x_ref3, wra_w=w,
                   wra_out=tu.weighted_rigid_align(x_pred, x_ref3, w),
                   wra_out_shared=tu.weighted_rigid_align(x_pred, x_ref3[0], w))
        # reflection case: gt is a mirror image of pred on the weighted atoms
        x_mirror = x_pred.clone()
        x_mirror[..., 0] *= -1
        kat.update(wra_mirror=x_mirror, wra_out_mirror=tu.weighted_rigid_align(x_pred, x_mirror, w))
    np.savez_compressed(os.path.join(OUT, "kat_modules.npz"), **{k: npy(v) for k, v in kat.items()},
                        **sd_sum)

    # ------------------------------------------------------------------ (ii) full denoiser
    cx = make_complex(64, 512, dims, seed=1)
    g = torch.Generator().manual_seed(3)
    full = dict(sd_sum)
    full["ap_checksum"] = checksum(cx["ap"])
    full["z_checksum"] = checksum(cx["z"])
    sd64 = {k: v.double() for k, v in sd.items()}
    dit64 = build_reference_dit("medium").double()
    dit64.load_state_dict(sd64)
    cx64 = {k: (v.double() if v.is_floating_point() else v) for k, v in cx.items()}
    for t in T_LEVELS:
        x_hat = torch.randn(4, 512, 3, generator=g) * (t ** 2 + 100) ** 0.5
        t_hat = torch.full([4], t)
        with torch.no_grad():
            y = dit(cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
            y64 = dit64(cx64, x_hat.double(), t_hat.double(), cx64["a"], cx64["ap"], cx64["s"], cx64["z"])
        full[f"x_hat_checksum_{t}"] = checksum(x_hat)
        full[f"x_denoised_{t}"] = npy(y)
        full[f"x_denoised_fp64_{t}"] = npy(y64.float())
    np.savez_compressed(os.path.join(OUT, "dit_64_512.npz"), **full)

    # ------------------------------------------------------------------ (iii) sampler traces
    class RefSampler(PhysDock):     # the reference sampler with the trunk replaced by cached outputs
        def __init__(self, dit_, cond):
            nn.Module.__init__(self)
            self.dit = dit_
            self.diffusion_conditioning = cond
            self.sigma_data = 16.0

    model = RefSampler(dit, lambda batch: (cx["a"], cx["ap"], cx["s"], cx["z"]))
    tmpl = make_templates(cx, 12)
    variants = {
        "nophys": dict(align_ref_pos=False),
        "refpos": dict(align_ref_pos=True),
        "templates": dict(align_ref_pos=True, ref_mol_poses=tmpl, mmff_gamma_0_factor=6.0),
    }
    for name, kw in variants.items():
        steps = []
        orig_forward = dit.forward

        def wrapped(batch, x_hat, t_hat, a, ap, s, z, _steps=steps):
            y = orig_forward(batch, x_hat, t_hat, a, ap, s, z)
            _steps.append((x_hat.clone(), t_hat.clone(), y.clone()))
            return y

        dit.forward = wrapped
        try:
            x_final, tape = record_tape(lambda: model.sample_diffusion(
                cx, num_sample=TRACE_B, steps=TRACE_STEPS, ref_mol=None,
                karras_noise_schedule_power=1000, **kw), seed=123)
        finally:
            dit.forward = orig_forward
        assert len(steps) == TRACE_STEPS
        d = dict(sd_sum, ap_checksum=checksum(cx["ap"]), n_tape=len(tape), x_final=npy(x_final))
        for i, t in enumerate(tape):
            d[f"tape_{i}"] = npy(t)
        for i, (xh, th, xd) in enumerate(steps):
            d[f"x_hat_{i}"], d[f"t_hat_{i}"], d[f"x_denoised_{i}"] = npy(xh), npy(th), npy(xd)
        np.savez_compressed(os.path.join(OUT, f"trace_{name}.npz"), **d)
        print(name, "tape", len(tape), "final |x|", float(x_final.abs().mean()))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
