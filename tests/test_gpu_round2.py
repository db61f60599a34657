"""Round-2 GPU parity tests: the full 40-step schedule, the real-data C1 fixture, the drop-in seam
(`B200DiT.from_reference` + the reference's `partial(dit, ...)` call), the hoisted conditioning, the fused Euler update
and CUDA-graph invalidation when the workspace grows."""
from functools import partial

import pytest
import torch
import torch.nn as nn

from oracle import physdock_oracle as O
from physdock_b200.synthetic import dit_param_shapes, make_templates
from tests.helpers import T, T_LEVELS, c1_fixture, complex_64_512, load_npz, log_value, medium_state

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_A = 1e-3


@pytest.fixture(scope="module")
def dit():
    from physdock_b200.dit import B200DiT
    dims, sd, _ = medium_state()
    return B200DiT.from_state_dict(sd, dims, device=DEV)


def to_dev(cx):
    return {k: v.to(DEV) for k, v in cx.items()}


# ------------------------------------------------------------------------------------------- full 40-step schedule
@pytest.mark.parametrize("name", ["nophys", "templates"])
def test_full_40_step_reference_trace_teacher_forced(dit, name):
    """steps=40, rho=1000 as redocking.py runs it: all 29 stochastic steps and the 11-step ODE tail (t_cur <= 1), replayed
    from the REAL reference's RNG tape (tests/golden/trace40_*.npz); per step x_denoised / x_next within 1e-3 A."""
    from physdock_b200 import sampler as S
    dims, sd, _ = medium_state()
    cx = complex_64_512()
    g = load_npz(f"trace40_{name}.npz")
    tape = [T(g[f"tape_{i}"]) for i in range(int(g["n_tape"]))]
    kw = dict(nophys=dict(align_ref_pos=False), templates=dict(align_ref_pos=True, mmff_gamma_0_factor=6.0))[name]
    if name == "templates":
        kw["ref_mol_poses"] = make_templates(cx, 12)
    ref_trace = []
    O.sample_diffusion(sd, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=2, steps=40,
                       karras_noise_schedule_power=1000, rng=O.ReplayRNG(tape), trace=ref_trace, **kw)
    d = to_dev(cx)
    if "ref_mol_poses" in kw:
        kw["ref_mol_poses"] = kw["ref_mol_poses"].to(DEV)
    trace = []
    x = S.sample_diffusion(dit, d, d["a"], d["ap"], d["s"], d["z"], num_sample=2, steps=40,
                           karras_noise_schedule_power=1000, rng=O.ReplayRNG(tape, DEV), trace=trace,
                           teacher=ref_trace, use_cuda_graph=(name == "nophys"), **kw)      # one trace per launch mode
    assert len(trace) == 40 and sum(1 for st in trace if st["t_cur"] > 1.0) == 29
    worst_den = worst_next = 0.0
    for i, (mine, ref) in enumerate(zip(trace, ref_trace)):
        assert abs(mine["t_hat"] - float(ref["t_hat"][0])) == 0.0, i
        r_den = float(O.rmsd(mine["x_denoised"].cpu(), T(g[f"x_denoised_{i}"])).max())
        r_next = float(O.rmsd(mine["x_next"].cpu(), ref["x_next"]).max())
        worst_den, worst_next = max(worst_den, r_den), max(worst_next, r_next)
        assert r_den < TOL_A and r_next < TOL_A, (name, i, r_den, r_next)
        if ref["used_inds"] is not None:
            assert torch.equal(mine["used_inds"].cpu(), ref["used_inds"]), (name, i)
    log_value(f"trace40[{name}] worst x_denoised rmsd", worst_den)
    log_value(f"trace40[{name}] worst x_next rmsd", worst_next)
    assert float(O.rmsd(x.cpu(), T(g["x_final"])).max()) < TOL_A


# ------------------------------------------------------------------------------------------- real data (BASELINE configs[0])
def test_c1_real_data_fixture(dit):
    """5SD5_HWI at crop 64/512 through the real FeatureLoader and the reference trunk (tests/golden/c1_5sd5.npz):
    Nt=64, Na=318 (not a multiple of any tile), real ragged chunk sizes and ligand layout."""
    from physdock_b200 import sampler as S
    g, batch, cond = c1_fixture()
    b = to_dev(batch)
    c = to_dev(cond)
    worst = 0.0
    for t in T_LEVELS:
        x_hat = T(g[f"x_hat_{t}"]).to(DEV)
        y = dit(b, x_hat, torch.full([4], t, device=DEV), c["a"], c["ap"], c["s"], c["z"])
        r = float(O.rmsd(y.cpu(), T(g[f"x_denoised_{t}"])).max())
        worst = max(worst, r)
        assert r < TOL_A, (t, r)
    log_value("c1_5sd5 denoiser worst rmsd vs reference", worst)
    # 4-sample, 12-step reference trace with ref_pos alignment, teacher-forced
    dims, sd, _ = medium_state()
    tape = [T(g[f"trace_tape_{i}"]) for i in range(int(g["trace_n_tape"]))]
    ref_trace = []
    O.sample_diffusion(sd, batch, cond["a"], cond["ap"], cond["s"], cond["z"], num_sample=4, steps=12,
                       karras_noise_schedule_power=1000, rng=O.ReplayRNG(tape), trace=ref_trace, align_ref_pos=True)
    trace = []
    x = S.sample_diffusion(dit, b, c["a"], c["ap"], c["s"], c["z"], num_sample=4, steps=12,
                           karras_noise_schedule_power=1000, rng=O.ReplayRNG(tape, DEV), trace=trace,
                           teacher=ref_trace, align_ref_pos=True)
    worst = 0.0
    for i, (mine, ref) in enumerate(zip(trace, ref_trace)):
        r_den = float(O.rmsd(mine["x_denoised"].cpu(), T(g[f"trace_x_denoised_{i}"])).max())
        r_next = float(O.rmsd(mine["x_next"].cpu(), ref["x_next"]).max())
        worst = max(worst, r_den, r_next)
        assert r_den < TOL_A and r_next < TOL_A, (i, r_den, r_next)
    log_value("c1_5sd5 12-step trace worst rmsd", worst)
    assert float(O.rmsd(x.cpu(), T(g["trace_x_final"])).max()) < TOL_A


# ------------------------------------------------------------------------------------------- the drop-in seam
class StandInAF3DiT(nn.Module):
    """A module with exactly AF3DiT's state_dict (transformers.py:178-262) -- what `model.dit` is in the reference -- for
    boxes where /root/reference is not mounted."""

    def __init__(self, sd):
        super().__init__()
        self.sigma_data = 16.0
        for key, v in sd.items():
            mod = self
            *path, leaf = key.split(".")
            for name in path:
                if name not in mod._modules:
                    mod.add_module(name, nn.Module())
                mod = mod._modules[name]
            mod.register_parameter(leaf, nn.Parameter(v.clone(), requires_grad=False))


def test_from_reference_and_reference_style_call():
    """`model.dit = B200DiT.from_reference(model.dit)`, then the reference sampler's own call
    `denoiser = partial(self.dit, batch=batch, a=a, ap=ap, s=s, z=z); denoiser(x_hat=..., t_hat=...)` (model.py:153,221)."""
    from physdock_b200.dit import B200DiT
    dims, sd, _ = medium_state()
    ref_dit = StandInAF3DiT(sd).to(DEV)
    assert list(ref_dit.state_dict().keys()) == list(dit_param_shapes(dims).keys())
    new = B200DiT.from_reference(ref_dit)
    assert next(new.parameters()).device.type == "cuda"
    assert list(new.state_dict().keys()) == list(ref_dit.state_dict().keys())
    cx = complex_64_512()
    d = to_dev(cx)
    denoiser = partial(new, batch=d, a=d["a"], ap=d["ap"], s=d["s"], z=d["z"])
    g = torch.Generator().manual_seed(5)
    outs = []
    for t in (900.0, 900.0, 2.0):        # second call replays the CUDA graph; third changes the noise level in place
        x_hat = torch.randn(3, 512, 3, generator=g) * (t ** 2 + 100) ** 0.5
        t_hat = torch.full([3], t)
        y = denoiser(x_hat=x_hat.to(DEV), t_hat=t_hat.to(DEV))
        with torch.no_grad():
            want = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        r = float(O.rmsd(y.cpu(), want).max())
        assert r < TOL_A, (t, r)
        outs.append(y)
    assert outs[0].data_ptr() != outs[1].data_ptr(), "forward must return a fresh tensor, not its static buffer"
    log_value("from_reference seam rmsd", r)
    # weights loaded later (import_state_dict, utils/import_weights.py:31-41) are picked up
    sd2 = {k: v * 1.01 for k, v in sd.items()}
    new.load_state_dict(sd2)
    y2 = denoiser(x_hat=x_hat.to(DEV), t_hat=t_hat.to(DEV))
    with torch.no_grad():
        want2 = O.af3dit_forward(sd2, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
    assert float(O.rmsd(y2.cpu(), want2).max()) < TOL_A
    assert not torch.equal(y2, outs[-1])


# ------------------------------------------------------------------------------------------- hoisted conditioning / fused Euler
def test_conditioning_table_and_fused_euler_are_bit_identical(dit):
    """`denoise_cond` with a row of the schedule's conditioning table == `denoise` computing the conditioning per call,
    bit for bit; the Euler update fused into the last kernel == pdk_euler_update, bit for bit."""
    from physdock_b200 import sampler as S
    cx = to_dev(complex_64_512())
    dit(cx, torch.zeros(1, 512, 3, device=DEV), torch.ones(1, device=DEV), cx["a"], cx["ap"], cx["s"], cx["z"])   # prepare
    t_list = torch.tensor([4608.0, 37.5, 0.4, 0.064], device=DEV)
    table = dit.conditioning_table(t_list)
    n_mod = dit.cond_width() - 8
    assert table.shape == (4, n_mod + 8)
    g = torch.Generator(device=DEV).manual_seed(1)
    for i, t in enumerate(t_list.tolist()):
        x_hat = torch.randn(3, 512, 3, generator=g, device=DEV) * (t ** 2 + 100) ** 0.5
        want = dit.denoise(x_hat, torch.full([3], t, device=DEV))
        got = dit.denoise_cond(x_hat, table[i], torch.empty_like(x_hat))
        assert torch.equal(got, want), (t, float((got - want).abs().max()))
        # per-sample rows (stride = row width) give the same bits as the shared row
        rows = table[i][None].repeat(3, 1).contiguous()
        assert torch.equal(dit.denoise_cond(x_hat, rows, torch.empty_like(x_hat)), want)
        # fused Euler: (t_next, eta) travel in entries 4, 5 of the coefficient block
        t_next, eta = 0.5 * t, 1.5
        row = table[i].clone()
        row[n_mod + 4], row[n_mod + 5] = t_next, eta
        x_next = torch.empty_like(x_hat)
        dit.denoise_cond(x_hat, row, torch.empty_like(x_hat), x_next)
        want_next = S.euler_update(x_hat, want, torch.full([3], t, device=DEV), t_next, eta)
        assert torch.equal(x_next, want_next), (t, float((x_next - want_next).abs().max()))
        ref_next = O.euler_update(x_hat.cpu(), (x_hat.cpu() - want.cpu()) / t, torch.full([3], t), torch.tensor(t_next), eta)
        assert torch.equal(x_next.cpu(), ref_next)


def test_graphs_survive_workspace_growth(dit):
    """a captured CUDA graph holds raw pointers into the workspace; a larger B reallocates it.  The small
    sampler must keep producing the same bits after a big one ran on the same complex."""
    from physdock_b200 import sampler as S
    cx = to_dev(complex_64_512())
    kw = dict(steps=6, karras_noise_schedule_power=1000, align_ref_pos=False)

    def run(n):
        torch.manual_seed(21)
        smp = S.DiffusionSampler(dit, cx, cx["a"], cx["ap"], cx["s"], cx["z"], num_sample=n, use_cuda_graph=True, **kw)
        smp.begin()
        return smp, [smp.step(i).clone() for i in range(3)]

    dit._workspace = None            # force the first workspace to be the small one
    dit._graphs.clear()
    small, first = run(2)
    ws_small = dit._workspace.numel()
    run(9)                           # grows (and reallocates) the workspace
    assert dit._workspace.numel() > ws_small
    scratch = torch.full((ws_small,), 255, dtype=torch.uint8, device=DEV)     # likely lands on the freed block
    torch.manual_seed(21)
    small.begin()
    again = [small.step(i).clone() for i in range(3)]
    del scratch
    for a, b in zip(first, again):
        assert torch.equal(a, b)


def test_forward_on_wrong_device_is_loud(dit):
    from physdock_b200._lib import PdkError
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cx = {k: v.to("cuda:1") for k, v in complex_64_512().items()}
    with pytest.raises(PdkError):
        dit.denoise(torch.zeros(1, 512, 3, device="cuda:1"), torch.ones(1, device="cuda:1"))
    del cx


def test_benchmarked_configuration_b16_256_2048_sampler_step(dit):
    """The configuration bench.py times (Nt=256, Na=2048, B=16: 4+4+4+4 / 4+3+3+3+3 attention work list, > 148-tile persistent
    GEMM loops, cta_group::2 tiles, CUDA-graph replay of pdk_dit_denoise_cond with the fused Euler update): one stochastic and
    one ODE-tail step of DiffusionSampler.step, teacher-forced against the host oracle."""
    from physdock_b200 import sampler as S
    from physdock_b200.synthetic import DiTDims, make_complex
    dims, sd, _ = medium_state()
    cx = make_complex(256, 2048, DiTDims.named("medium"), seed=1)
    d = to_dev(cx)
    torch.manual_seed(5)
    smp = S.DiffusionSampler(dit, d, d["a"], d["ap"], d["s"], d["z"], num_sample=16, steps=40,
                             karras_noise_schedule_power=1000, align_ref_pos=False, use_cuda_graph=True)
    smp.begin()
    for i in (2, 34):
        if i == 34:      # a plausible late-step state: the structure plus noise at the step's sigma
            smp.x_next.copy_(d["x_gt"][None] + float(smp.sigmas[i]) * torch.randn(16, 2048, 3, device=DEV))
        smp.step(i)
        t_cur, t_next, t_hat, stochastic, _ = smp.schedule(i)
        x_hat, x_den, x_next = smp.x_hat.cpu(), smp.x_den.cpu(), smp.x_next.cpu()
        th = torch.full([16], float(t_hat))
        with torch.no_grad():
            want = O.af3dit_forward(sd, cx, x_hat, th, cx["a"], cx["ap"], cx["s"], cx["z"])
        want_next = O.euler_update(x_hat, (x_hat - want) / th[:, None, None], th, t_next, 1.5 if stochastic else 1.0)
        r_den, r_next = float(O.rmsd(x_den, want).max()), float(O.rmsd(x_next, want_next).max())
        log_value(f"B=16 256/2048 step {i}: x_denoised rmsd", r_den)
        log_value(f"B=16 256/2048 step {i}: x_next rmsd", r_next)
        assert r_den < TOL_A and r_next < TOL_A, (i, r_den, r_next)
