"""GPU parity of the whole denoiser (AF3DiT.forward) through the drop-in module B200DiT."""
import pytest
import torch

from oracle import physdock_oracle as O
from physdock_b200.synthetic import DiTDims, make_complex
from tests.helpers import T, load_npz, medium_state, complex_64_512, dit_inputs_64_512, log_value

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_A = 1e-3      # Angstrom RMSD, the north-star budget (BASELINE.json)


@pytest.fixture(scope="module")
def dit():
    from physdock_b200.dit import B200DiT
    dims, sd, _ = medium_state()
    return B200DiT.from_state_dict(sd, dims, device=DEV)


def to_dev(cx):
    return {k: v.to(DEV) for k, v in cx.items()}


def test_denoiser_matches_reference_golden_64_512(dit):
    """Outputs of the REAL reference AF3DiT (tests/golden/dit_64_512.npz) at 5 noise levels."""
    cx = to_dev(complex_64_512())
    g = load_npz("dit_64_512.npz")
    for t, x_hat, t_hat in dit_inputs_64_512():
        y = dit(cx, x_hat.to(DEV), t_hat.to(DEV), cx["a"], cx["ap"], cx["s"], cx["z"]).cpu()
        r = float(O.rmsd(y, T(g[f"x_denoised_{t}"])).max())
        r64 = float(O.rmsd(y, T(g[f"x_denoised_fp64_{t}"])).max())
        log_value(f"dit64/512 t={t} rmsd_vs_ref_fp32", r)
        log_value(f"dit64/512 t={t} rmsd_vs_ref_fp64", r64)
        assert r < TOL_A, (t, r)
        assert torch.isfinite(y).all()


def test_denoiser_ragged_masked_complex_vs_oracle(dit):
    """Ragged chunk sizes, zero-size tokens, masked atoms/tokens, sizes far from tile multiples."""
    dims, sd, _ = medium_state()
    for (Nt, Na, seed) in ((24, 100, 9), (40, 333, 10)):
        cx = make_complex(Nt, Na, dims, seed=seed, ragged=True, mask_holes=True)
        g = torch.Generator().manual_seed(seed)
        x_hat = torch.randn(3, Na, 3, generator=g) * 30
        t_hat = torch.tensor([300.0, 5.0, 0.1])
        with torch.no_grad():
            want = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
        d = to_dev(cx)
        got = dit(d, x_hat.to(DEV), t_hat.to(DEV), d["a"], d["ap"], d["s"], d["z"]).cpu()
        # atoms that are masked out everywhere attend uniformly in the reference and are never used downstream
        live = cx["a_mask"].bool()
        r = float(O.rmsd(got[:, live], want[:, live]).max())
        log_value(f"dit ragged {Nt}/{Na} rmsd", r)
        assert r < TOL_A, (Nt, Na, r)


def test_denoiser_256_2048_vs_oracle(dit):
    """The benchmark shape (BASELINE.json configs[1]) against the CPU oracle, B=2."""
    dims, sd, _ = medium_state()
    cx = make_complex(256, 2048, dims, seed=1)
    g = torch.Generator().manual_seed(5)
    x_hat = torch.randn(2, 2048, 3, generator=g) * 50
    t_hat = torch.tensor([40.0, 0.5])
    with torch.no_grad():
        want = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
    d = to_dev(cx)
    got = dit(d, x_hat.to(DEV), t_hat.to(DEV), d["a"], d["ap"], d["s"], d["z"]).cpu()
    r = float(O.rmsd(got, want).max())
    log_value("dit256/2048 rmsd", r)
    assert r < TOL_A, r


def test_denoiser_384_3072_vs_oracle(dit):
    """BASELINE.json configs[2] shape (crop 384 / atom crop 3072) against the CPU oracle, ragged token layout, B=2."""
    dims, sd, _ = medium_state()
    cx = make_complex(384, 3072, dims, seed=4, ragged=True)
    g = torch.Generator().manual_seed(6)
    x_hat = torch.randn(2, 3072, 3, generator=g) * 30
    t_hat = torch.tensor([25.0, 1.5])
    with torch.no_grad():
        want = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
    d = to_dev(cx)
    got = dit(d, x_hat.to(DEV), t_hat.to(DEV), d["a"], d["ap"], d["s"], d["z"]).cpu()
    r = float(O.rmsd(got, want).max())
    log_value("dit384/3072 rmsd", r)
    assert r < TOL_A, r


def test_dropin_contract_and_caching(dit):
    """forward() has AF3DiT's signature; the per-complex cache is keyed on the conditioning tensors."""
    cx = to_dev(complex_64_512())
    x_hat = torch.randn(2, 512, 3, device=DEV) * 10
    t_hat = torch.full([2], 3.0, device=DEV)
    y1 = dit(cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
    keep = dit._complex_keep
    y2 = dit(batch=cx, x_hat=x_hat, t_hat=t_hat, a=cx["a"], ap=cx["ap"], s=cx["s"], z=cx["z"])
    assert dit._complex_keep is keep, "second call must reuse the cached pair bias"
    assert torch.equal(y1, y2), "denoiser must be deterministic"
    # samples are independent: evaluating sample 1 alone gives the same bits
    y_single = dit(cx, x_hat[1:], t_hat[1:], cx["a"], cx["ap"], cx["s"], cx["z"])
    assert torch.equal(y_single[0], y1[1])
    cx["ap"].mul_(1.0001)          # in-place edit bumps the version counter -> cache must refresh
    y3 = dit(cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
    assert dit._complex_keep is not keep
    assert not torch.equal(y3, y1)


def test_cpu_tensors_fail_loudly(dit):
    from physdock_b200._lib import PdkError
    cx = complex_64_512()
    with pytest.raises(PdkError):
        dit(cx, torch.zeros(1, 512, 3), torch.ones(1), cx["a"], cx["ap"], cx["s"], cx["z"])


@pytest.mark.parametrize("name", ["toy", "tiny", "small", "full"])
def test_other_model_sizes_vs_oracle(name):
    """The reference defines five model sizes that differ only in block counts (configs.py:63-93: toy 2+2/2, tiny 2+2/4,
    small 2+2/8, medium 3+3/12, full 3+3/24 atom/token DiT blocks): the handle, the modulation table and the launch plan take
    the counts from the dims.  Ragged complex, three noise levels, against the CPU oracle."""
    from physdock_b200.dit import B200DiT
    from physdock_b200.synthetic import make_dit_state
    dims = DiTDims.named(name)
    sd = make_dit_state(dims, seed=3)
    dit = B200DiT.from_state_dict(sd, dims, device=DEV)
    cx = make_complex(24, 100, dims, seed=11, ragged=True)
    g = torch.Generator().manual_seed(5)
    x_hat = torch.randn(3, 100, 3, generator=g) * 30
    t_hat = torch.tensor([300.0, 5.0, 0.1])
    with torch.no_grad():
        want = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
    d = to_dev(cx)
    got = dit(d, x_hat.to(DEV), t_hat.to(DEV), d["a"], d["ap"], d["s"], d["z"]).cpu()
    r = float(O.rmsd(got, want).max())
    log_value(f"dit model={name} 24/100 rmsd", r)
    assert r < TOL_A, (name, r)
    n_atom, n_tok = 2 * dims.no_blocks_atom, dims.no_blocks_dit
    assert dit.launches_per_denoise() == 2 + 5 * n_atom + 7 * n_tok + 6


@pytest.mark.parametrize("Nt,Na,B", [(2, 5, 1), (3, 130, 1), (2, 5, 7)])
def test_tiny_complexes_and_single_sample(dit, Nt, Na, B):
    """Smallest inputs the feature pipeline can produce (a handful of atoms: every tile is almost all padding; a token
    owning 129 atoms straddles a 128-row tile; a zero-size token) and B = 1."""
    dims, sd, _ = medium_state()
    cx = make_complex(Nt, Na, dims, seed=21, ragged=True)
    g = torch.Generator().manual_seed(Na)
    x_hat = torch.randn(B, Na, 3, generator=g) * 10
    t_hat = torch.full([B], 40.0)
    with torch.no_grad():
        want = O.af3dit_forward(sd, cx, x_hat, t_hat, cx["a"], cx["ap"], cx["s"], cx["z"])
    d = to_dev(cx)
    got = dit(d, x_hat.to(DEV), t_hat.to(DEV), d["a"], d["ap"], d["s"], d["z"]).cpu()
    r = float(O.rmsd(got, want).max())
    log_value(f"dit tiny {Nt}/{Na} B={B} rmsd", r)
    assert torch.isfinite(got).all() and r < TOL_A, (Nt, Na, B, r)
