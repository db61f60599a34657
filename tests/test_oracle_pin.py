"""Pins the oracle (oracle/physdock_oracle.py) to the REAL reference:
 (1) against golden vectors generated from the real reference (always runs, also on the GPU box), and
 (2) live against /root/reference when that tree is mounted (this container only)."""
import pytest
import torch

from oracle import physdock_oracle as O
from oracle.ref_import import reference_available
from physdock_b200.synthetic import DiTDims, dit_param_shapes, make_complex, make_templates
from tests.helpers import (T, T_LEVELS, load_npz, medium_state, complex_64_512, dit_inputs_64_512,
                           checksum)


def test_fixture_inputs_regenerate_bit_identically():
    """Weights / complex / x_hat are rebuilt from seeds; the fixtures carry their checksums."""
    dims, sd, sd_sum = medium_state()
    g = load_npz("dit_64_512.npz")
    assert sd_sum == float(g["sd_checksum"])
    cx = complex_64_512()
    assert checksum(cx["ap"]) == float(g["ap_checksum"])
    assert checksum(cx["z"]) == float(g["z_checksum"])
    for t, x_hat, _ in dit_inputs_64_512():
        assert checksum(x_hat) == float(g[f"x_hat_checksum_{t}"])


def test_oracle_denoiser_matches_golden_bitwise():
    dims, sd, _ = medium_state()
    cx = complex_64_512()
    g = load_npz("dit_64_512.npz")
    for t, x_hat, t_hat in dit_inputs_64_512():
        with torch.no_grad():
            y = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        want = T(g[f"x_denoised_{t}"])
        # same torch build + same ops => bit-identical; allow a few ulp if thread count changes reductions
        assert float(O.rmsd(y, want).max()) < 2e-5, t
        # the fp32 reference's own distance from an fp64 evaluation bounds the meaningful tolerance
        assert float(O.rmsd(want, T(g[f"x_denoised_fp64_{t}"])).max()) < 2e-4


def test_oracle_module_kats():
    dims, sd, _ = medium_state()
    k = {n: T(v) for n, v in load_npz("kat_modules.npz").items()}
    eps, inf = dims.eps, dims.inf
    pa, pt = "atom_dit_encoder.blocks.1.", "token_dit.blocks.5."
    with torch.no_grad():
        got = {
            "atom_attn_out": O.dit_attention(sd, pa + "attention.", k["ba"], k["ap"], k["t_emb"], k["ap_mask"], inf, eps),
            "atom_trans_out": O.dit_transition(sd, pa + "transition.", k["ba"], k["t_emb"], eps),
            "tok_attn_out": O.dit_attention(sd, pt + "attention.", k["bs"], k["z"], k["t_emb"], k["z_mask"], inf, eps),
            "tok_attn_holes_out": O.dit_attention(sd, pt + "attention.", k["bs"], k["z"], k["t_emb"], k["z_mask_holes"], inf, eps),
            "tok_trans_out": O.dit_transition(sd, pt + "transition.", k["bs"], k["t_emb"], eps),
            "atom_adaln_x": O.ada_layer_norm_zero(sd, pa + "attention.norm_s.", k["ba"], k["t_emb"], eps)[0],
            "precond_ba": O.precond(sd, k["k_x_hat"], k["k_t_hat"], k["k_a"], 16.0)[0],
            "precond_t": O.precond(sd, k["k_x_hat"], k["k_t_hat"], k["k_a"], 16.0)[1],
            "denoise_out": O.denoise(sd, k["k_x_hat"], k["k_t_hat"], k["ba"], 16.0, eps),
            "downscale_out": O.downscale(sd, k["ba"], k["k_s"], k["chunk"]),
            "upscale_out": O.upscale(sd, k["ba"], k["bs"], k["a2t"]),
            "cra_out": O.centre_random_augmentation(k["cra_x"], k["cra_exists"], k["cra_u"], k["cra_trans"]),
            "wra_out": O.weighted_rigid_align(k["wra_pred"], k["wra_gt"], k["wra_w"]),
            "wra_out_shared": O.weighted_rigid_align(k["wra_pred"], k["wra_gt"][0], k["wra_w"]),
            "wra_out_mirror": O.weighted_rigid_align(k["wra_pred"], k["wra_mirror"], k["wra_w"]),
        }
    for name, y in got.items():
        err = float((y - k[name]).abs().max())
        scale = float(k[name].abs().max())
        assert err <= 2e-6 * max(1.0, scale), (name, err, scale)


@pytest.mark.parametrize("name", ["nophys", "refpos", "templates"])
def test_oracle_sampler_replays_reference_trace(name):
    """Replays the reference's recorded RNG tape through the oracle sampler; every step must match."""
    dims, sd, _ = medium_state()
    cx = complex_64_512()
    g = load_npz(f"trace_{name}.npz")
    tape = [T(g[f"tape_{i}"]) for i in range(int(g["n_tape"]))]
    kw = dict(nophys=dict(align_ref_pos=False), refpos=dict(align_ref_pos=True),
              templates=dict(align_ref_pos=True, ref_mol_poses=make_templates(cx, 12),
                             mmff_gamma_0_factor=6.0))[name]
    trace = []
    rng = O.ReplayRNG(tape)
    x = O.sample_diffusion(sd, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=2, steps=12,
                           karras_noise_schedule_power=1000, rng=rng, trace=trace, **kw)
    assert rng.pos == len(tape)
    for i, st in enumerate(trace):
        assert torch.equal(st["t_hat"], T(g[f"t_hat_{i}"])), i
        assert float(O.rmsd(st["x_hat"], T(g[f"x_hat_{i}"])).max()) < 1e-4, i
        assert float(O.rmsd(st["x_denoised"], T(g[f"x_denoised_{i}"])).max()) < 1e-4, i
    assert float(O.rmsd(x, T(g["x_final"])).max()) < 1e-4


@pytest.mark.parametrize("name", ["nophys", "templates"])
def test_oracle_replays_full_40_step_reference_trace(name):
    """The FULL redocking schedule (steps=40, rho=1000: 29 stochastic steps + the 11-step ODE tail), recorded from the real
    reference by oracle/make_golden_r02.py."""
    dims, sd, sd_sum = medium_state()
    cx = complex_64_512()
    g = load_npz(f"trace40_{name}.npz")
    assert sd_sum == float(g["sd_checksum"]) and checksum(cx["ap"]) == float(g["ap_checksum"])
    tape = [T(g[f"tape_{i}"]) for i in range(int(g["n_tape"]))]
    kw = dict(nophys=dict(align_ref_pos=False),
              templates=dict(align_ref_pos=True, ref_mol_poses=make_templates(cx, 12), mmff_gamma_0_factor=6.0))[name]
    trace = []
    rng = O.ReplayRNG(tape)
    x = O.sample_diffusion(sd, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=2, steps=40,
                           karras_noise_schedule_power=1000, rng=rng, trace=trace, **kw)
    assert rng.pos == len(tape) and len(trace) == 40
    assert sum(1 for st in trace if st["t_cur"] > 1.0) == 29
    for i, st in enumerate(trace):
        assert torch.equal(st["t_hat"], T(g[f"t_hat_{i}"])), i
        assert float(O.rmsd(st["x_hat"], T(g[f"x_hat_{i}"])).max()) < 1e-4, i
        assert float(O.rmsd(st["x_denoised"], T(g[f"x_denoised_{i}"])).max()) < 1e-4, i
    assert float(O.rmsd(x, T(g["x_final"])).max()) < 1e-4


def test_oracle_on_real_data_c1_fixture():
    """BASELINE.json configs[0]: real FeatureLoader output of 5SD5_HWI (Nt=64, Na=318, 29 ligand atoms, ragged chunks)."""
    from tests.helpers import c1_fixture
    dims, sd, sd_sum = medium_state()
    g, batch, cond = c1_fixture()
    assert sd_sum == float(g["sd_checksum"])
    Nt, Na = int(g["Nt"]), int(g["Na"])
    assert (Nt, Na) == (64, 318) and batch["atom_id_to_token_id"].shape == (Na,)
    assert int(batch["token_id_to_chunk_sizes"].sum()) == Na and int(batch["token_id_to_chunk_sizes"].max()) > 1
    assert int(batch["is_ligand"][batch["atom_id_to_token_id"]].sum()) == 29
    for t in T_LEVELS:
        with torch.no_grad():
            y = O.af3dit_forward(sd, batch, T(g[f"x_hat_{t}"]), torch.full([4], t), cond["a"], cond["ap"], cond["s"], cond["z"])
        assert float(O.rmsd(y, T(g[f"x_denoised_{t}"])).max()) < 2e-5, t
    tape = [T(g[f"trace_tape_{i}"]) for i in range(int(g["trace_n_tape"]))]
    trace = []
    rng = O.ReplayRNG(tape)
    x = O.sample_diffusion(sd, batch, cond["a"], cond["ap"], cond["s"], cond["z"], num_sample=4, steps=12,
                           karras_noise_schedule_power=1000, rng=rng, trace=trace, align_ref_pos=True)
    assert rng.pos == len(tape)
    for i, st in enumerate(trace):
        assert float(O.rmsd(st["x_denoised"], T(g[f"trace_x_denoised_{i}"])).max()) < 1e-4, i
    assert float(O.rmsd(x, T(g["trace_x_final"])).max()) < 1e-4


def test_schedule_facts():
    """rho=1000, 40 steps => 29 stochastic steps (t_cur>1), 17 with t_cur<=6."""
    s = O.karras_noise_schedule(40, p=1000)
    assert s.shape == (41,) and float(s[-1]) == 0.0
    assert abs(float(s[0]) - 2560.0) < 1.0 and abs(float(s[-2]) - 0.064) < 1e-3
    assert int((s[:-1] > 1.0).sum()) == 29
    assert int((s[:-1] <= 6.0).sum()) == 17


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
def test_oracle_matches_live_reference():
    from oracle.ref_import import build_reference_dit, import_reference
    import torch.nn as nn
    dims, sd, _ = medium_state()
    dit = build_reference_dit("medium")
    assert list(dit.state_dict().keys()) == list(dit_param_shapes(dims).keys())
    for k_, v in dit.state_dict().items():
        assert tuple(v.shape) == dit_param_shapes(dims)[k_], k_
    dit.load_state_dict(sd)
    cx = make_complex(24, 100, dims, seed=9, ragged=True, mask_holes=True)
    g = torch.Generator().manual_seed(4)
    x_hat = torch.randn(3, 100, 3, generator=g) * 30
    t_hat = torch.tensor([300.0, 5.0, 0.1])
    with torch.no_grad():
        y_ref = dit(cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        y = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
    assert torch.equal(y, y_ref)
    PhysDock, _, _, _ = import_reference()

    class RefSampler(PhysDock):
        def __init__(self, dit_, cond):
            nn.Module.__init__(self)
            self.dit, self.diffusion_conditioning, self.sigma_data = dit_, cond, 16.0

    m = RefSampler(dit, lambda batch: (cx["a"], cx["ap"], cx["s"], cx["z"]))
    tmpl = make_templates(cx, 7)
    for kw in [dict(align_ref_pos=False), dict(align_ref_pos=True),
               dict(align_ref_pos=True, ref_mol_poses=tmpl, mmff_gamma_0_factor=6.0)]:
        torch.manual_seed(77)
        x_ref = m.sample_diffusion(cx, num_sample=3, steps=6, ref_mol=None,
                                   karras_noise_schedule_power=1000, **kw)
        torch.manual_seed(77)
        x = O.sample_diffusion(sd, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=3, steps=6,
                               karras_noise_schedule_power=1000, **kw)
        assert torch.equal(x, x_ref), kw.keys()
