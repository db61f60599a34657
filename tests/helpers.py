"""Shared test helpers (CPU side).  The oracle is imported here as the checker only."""
import os

import numpy as np
import torch

from physdock_b200.synthetic import DiTDims, make_dit_state, make_complex, checksum

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
T_LEVELS = [4608.0, 100.0, 10.0, 1.0, 0.2]
_cache = {}


def load_npz(name):
    if name not in _cache:
        with np.load(os.path.join(GOLDEN, name)) as f:
            _cache[name] = {k: f[k] for k in f.files}
    return _cache[name]


def T(a):
    return torch.from_numpy(np.asarray(a))


def medium_state():
    if "sd" not in _cache:
        dims = DiTDims.named("medium")
        sd = make_dit_state(dims, seed=0)
        _cache["sd"] = (dims, sd, checksum(torch.cat([v.flatten() for v in sd.values()])))
    return _cache["sd"]


def complex_64_512():
    if "cx" not in _cache:
        _cache["cx"] = make_complex(64, 512, DiTDims.named("medium"), seed=1)
    return _cache["cx"]


def dit_inputs_64_512():
    """Regenerates the x_hat / t_hat sequence of oracle/make_golden.py section (ii)."""
    g = torch.Generator().manual_seed(3)
    out = []
    for t in T_LEVELS:
        x_hat = torch.randn(4, 512, 3, generator=g) * (t ** 2 + 100) ** 0.5
        out.append((t, x_hat, torch.full([4], t)))
    return out


def rel_close(name, got, want, rtol, atol):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    assert got.shape == want.shape, f"{name}: shape {tuple(got.shape)} != {tuple(want.shape)}"
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    if bool(bad.any()):
        idx = torch.nonzero(bad)[0].tolist()
        raise AssertionError(
            f"{name}: {int(bad.sum())}/{bad.numel()} off; max abs err {float(err.max()):.3e} "
            f"(|want| max {float(want.abs().max()):.3e}); first bad at {idx}: got {float(got[tuple(idx)]):.6e} "
            f"want {float(want[tuple(idx)]):.6e}")
    _log(name, float(err.max()), float(tol.min()), float(want.abs().max()))
    return float(err.max())


def _log(name, err, tol, scale):
    """Appends (name, max error, tolerance, |want| max) to gpurun_out/parity_log.txt so tolerances can be audited."""
    d = os.path.join(os.path.dirname(GOLDEN.rstrip("/")), "..", "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_log.txt"), "a") as f:
            f.write(f"{name}\terr={err:.3e}\ttol={tol:.3e}\tscale={scale:.3e}\n")
    except OSError:
        pass


def log_value(name, value):
    _log(name, float(value), float("nan"), float("nan"))


def c1_fixture():
    """BASELINE.json configs[0] on real data (tests/golden/c1_5sd5.npz, oracle/make_golden_r02.py): the batch keys the hot
    path reads, the (fp16-stored) reference trunk outputs, reference AF3DiT outputs and a 12-step sampler trace."""
    if "c1" not in _cache:
        g = load_npz("c1_5sd5.npz")
        batch = {k[len("batch_"):]: T(g[k]) for k in g if k.startswith("batch_")}
        batch["token_bonds"] = T(g["extra_token_bonds"])
        cond = {k: T(g[k]).float() for k in ("a", "ap", "s", "z")}
        _cache["c1"] = (g, batch, cond)
    return _cache["c1"]
