"""GPU parity of the sampler loop (PhysDock.sample_diffusion) against the reference traces and the oracle."""
import pytest
import torch

from oracle import physdock_oracle as O
from physdock_b200.synthetic import make_templates
from tests.helpers import T, load_npz, medium_state, complex_64_512, log_value

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_A = 1e-3
VARIANTS = dict(nophys=dict(align_ref_pos=False), refpos=dict(align_ref_pos=True),
                templates=dict(align_ref_pos=True, mmff_gamma_0_factor=6.0))


@pytest.fixture(scope="module")
def dit():
    from physdock_b200.dit import B200DiT
    dims, sd, _ = medium_state()
    return B200DiT.from_state_dict(sd, dims, device=DEV)


def to_dev(cx):
    return {k: v.to(DEV) for k, v in cx.items()}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_teacher_forced_trace_vs_reference(dit, name):
    """Replays the REAL reference's RNG tape (tests/golden/trace_*.npz); at every step the reference's x_hat is
    fed to the B200 denoiser and x_denoised / x_next must match within 1e-3 A RMSD."""
    from physdock_b200 import sampler as S
    dims, sd, _ = medium_state()
    cx = complex_64_512()
    g = load_npz(f"trace_{name}.npz")
    tape = [T(g[f"tape_{i}"]) for i in range(int(g["n_tape"]))]
    kw = dict(VARIANTS[name])
    if name == "templates":
        kw["ref_mol_poses"] = make_templates(cx, 12)
    # reference-side per-step quantities (x_next, chosen templates) from the oracle, which test_oracle_pin.py
    # shows reproduces these traces
    ref_trace = []
    O.sample_diffusion(sd, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=2, steps=12,
                       karras_noise_schedule_power=1000, rng=O.ReplayRNG(tape), trace=ref_trace, **kw)
    for i, st in enumerate(ref_trace):
        assert float(O.rmsd(st["x_hat"], T(g[f"x_hat_{i}"])).max()) < 1e-4
    d = to_dev(cx)
    if "ref_mol_poses" in kw:
        kw["ref_mol_poses"] = kw["ref_mol_poses"].to(DEV)
    trace = []
    x = S.sample_diffusion(dit, d, d["a"], d["ap"], d["s"], d["z"], num_sample=2, steps=12,
                           karras_noise_schedule_power=1000, rng=O.ReplayRNG(tape, DEV), trace=trace,
                           teacher=ref_trace, **kw)
    worst_den = worst_next = 0.0
    for i, (mine, ref) in enumerate(zip(trace, ref_trace)):
        assert abs(mine["t_hat"] - float(ref["t_hat"][0])) == 0.0, i
        r_den = float(O.rmsd(mine["x_denoised"].cpu(), T(g[f"x_denoised_{i}"])).max())
        r_next = float(O.rmsd(mine["x_next"].cpu(), ref["x_next"]).max())
        worst_den, worst_next = max(worst_den, r_den), max(worst_next, r_next)
        assert r_den < TOL_A, (name, i, r_den)
        assert r_next < TOL_A, (name, i, r_next)
        if ref["used_inds"] is not None:
            assert torch.equal(mine["used_inds"].cpu(), ref["used_inds"]), (name, i)
    log_value(f"trace[{name}] worst x_denoised rmsd", worst_den)
    log_value(f"trace[{name}] worst x_next rmsd", worst_next)
    assert float(O.rmsd(x.cpu(), T(g["x_final"])).max()) < TOL_A


def test_same_seed_same_device_rng_stream(dit):
    """Identical RNG state on the same device type: the oracle sampler on CUDA and the B200 sampler, both under
    torch.manual_seed(7), must consume the generator identically (teacher-forced per-step comparison)."""
    from physdock_b200 import sampler as S
    dims, sd, _ = medium_state()
    cx = to_dev(complex_64_512())
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    torch.manual_seed(7)
    ref_trace = []
    O.sample_diffusion(sd_dev, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=3, steps=10,
                       karras_noise_schedule_power=1000, align_ref_pos=True, trace=ref_trace)
    torch.manual_seed(7)
    trace = []
    S.sample_diffusion(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=3, steps=10,
                       karras_noise_schedule_power=1000, align_ref_pos=True, trace=trace, teacher=ref_trace)
    for i, (mine, ref) in enumerate(zip(trace, ref_trace)):
        r = float(O.rmsd(mine["x_next"], ref["x_next"]).max())
        assert r < TOL_A, (i, r)
    # and the generator ends in the same state
    assert torch.equal(torch.rand(4, device=DEV), torch.rand(4, device=DEV)) is False
    torch.manual_seed(7)
    O.sample_diffusion(sd_dev, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=3, steps=3,
                       karras_noise_schedule_power=1000)
    a = torch.rand(4, device=DEV)
    torch.manual_seed(7)
    S.sample_diffusion(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=3, steps=3,
                       karras_noise_schedule_power=1000)
    assert torch.equal(a, torch.rand(4, device=DEV))


def test_free_running_trajectory_stays_close(dit):
    """No teacher forcing: 12 free-running steps from the reference's tape end near the reference's x_final."""
    from physdock_b200 import sampler as S
    cx = complex_64_512()
    g = load_npz("trace_nophys.npz")
    tape = [T(g[f"tape_{i}"]) for i in range(int(g["n_tape"]))]
    d = to_dev(cx)
    x = S.sample_diffusion(dit, d, d["a"], d["ap"], d["s"], d["z"], num_sample=2, steps=12,
                           karras_noise_schedule_power=1000, rng=O.ReplayRNG(tape, DEV), align_ref_pos=False)
    r = float(O.rmsd(x.cpu(), T(g["x_final"])).max())
    log_value("free-running 12-step final rmsd", r)
    assert r < 1e-2, r


def test_physdock_module_surface(dit):
    """PhysDockB200 keeps the keyword set redocking.py:284-299 passes to sample_diffusion."""
    from physdock_b200.sampler import PhysDockB200
    cx = to_dev(complex_64_512())
    model = PhysDockB200(dit, diffusion_conditioning=lambda batch: (cx["a"], cx["ap"], cx["s"], cx["z"]))
    x = model.sample_diffusion(cx, num_sample=2, steps=4, gamma_0=0.8, gamma_min=1.0, noise_scale_lambda=1.003,
                               step_scale_eta=1.5, ode_step_scale_eta=1.0, ref_mol=None, ref_mol_poses=None,
                               use_ref_mol_poses=False, mmff_gamma_0_factor=1.0, mmff_iters=5, align_ref_pos=True,
                               karras_noise_schedule_power=1000)
    assert x.shape == (2, 512, 3) and torch.isfinite(x).all()
    from physdock_b200._lib import PdkError
    with pytest.raises(PdkError):
        model.sample_diffusion(cx, num_sample=2, steps=4, ref_mol=object())    # MMFF needs rdkit: loud, not silent


def test_sharded_sampler_single_rank_equals_plain(dit):
    """world_size 1: sample_diffusion_sharded(exact=True) consumes the generator like the plain sampler."""
    from physdock_b200 import sampler as S
    from physdock_b200.sharding import sample_diffusion_sharded
    cx = to_dev(complex_64_512())
    torch.manual_seed(11)
    a = S.sample_diffusion(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=3, steps=5,
                           karras_noise_schedule_power=1000, align_ref_pos=False)
    b = sample_diffusion_sharded(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=3, seed=11, steps=5,
                                 karras_noise_schedule_power=1000, align_ref_pos=False)
    assert torch.equal(a, b)


def test_late_step_with_pair_energy_field(dit):
    """Late steps (t_cur <= gamma_min * factor) with the opt-in GPU physics backend: x_next must equal the oracle's
    composition descend -> weighted_rigid_align -> blended direction -> Euler applied to the same x_denoised."""
    from physdock_b200 import sampler as S
    from physdock_b200.physics import PairEnergyField
    from physdock_b200.synthetic import make_ligand_field
    cx = complex_64_512()
    d = to_dev(cx)
    Na = cx["x_gt"].shape[0]
    lig = cx["is_ligand"][cx["atom_id_to_token_id"]].bool()
    rows = torch.nonzero(lig).flatten().int()
    n_lig = int(rows.numel())
    f = make_ligand_field(Na, n_lig, seed=2, missing=False)
    # the field's bonded chain sits on the complex's ligand atoms (synthetic layout: ligand tokens come last)
    assert torch.equal(rows, f["rows"])
    fld = PairEnergyField(d["a_mask"], f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"], rows=rows)
    torch.manual_seed(4)
    smp = S.DiffusionSampler(dit, d, d["a"], d["ap"], d["s"], d["z"], num_sample=3, steps=12,
                             karras_noise_schedule_power=1000, align_ref_pos=True, mmff_gamma_0_factor=6.0,
                             mmff_iters=5, physics_field=fld, physics_step=0.002)
    smp.begin()
    late = [i for i in range(12) if float(smp.sigmas[i]) <= 6.0]
    assert late, "schedule has no late step"
    i = late[0]
    # a plausible late-step state: ligand chain + globule with noise at the step's sigma
    g = torch.Generator().manual_seed(9)
    smp.x_next = (f["x0"][None] + float(smp.sigmas[i]) * torch.randn(3, Na, 3, generator=g)).to(DEV).contiguous()
    x_next = smp.step(i).cpu()
    x_hat, x_den = smp.x_hat.cpu(), smp.x_den.cpu()
    t_cur, t_next, t_hat, stochastic, _ = smp.schedule(i)
    args = (cx["a_mask"], f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"], rows)
    x_ref = O.pair_energy_descend(x_den, *args, iters=5, step=0.002, gmax=50.0)
    w = cx["a_mask"] * lig.float()
    aligned = O.weighted_rigid_align(x_den * cx["a_mask"][..., None], x_ref, w)
    th = torch.full([3], float(t_hat))
    d_cur = O.physics_direction(x_hat, x_den, aligned, th, w)
    want = O.euler_update(x_hat, d_cur, th, t_next, 1.5 if stochastic else 1.0)
    r = float(O.rmsd(x_next, want).max())
    log_value("late step with pair-energy field: x_next rmsd vs oracle composition", r)
    assert r < TOL_A, r
