"""Pair-energy physics backend: oracle self-checks on CPU, CUDA-vs-oracle parity on the GPU (through the C ABI)."""
import math

import pytest
import torch

from oracle import physdock_oracle as O            # checker only
from physdock_b200.physics import build_partner_table, PairEnergyParams
from physdock_b200.synthetic import make_ligand_field
from tests.helpers import rel_close


def _case(Na, n_lig, B, seed, missing=True):
    f = make_ligand_field(Na, n_lig, seed=seed, missing=missing)
    g = torch.Generator().manual_seed(seed + 100)
    x = f["x0"][None] + 0.3 * torch.randn(B, Na, 3, generator=g)
    return f, x


def test_partner_table_symmetric():
    partner, r0, k = build_partner_table(5, [(0, 1, 1.5, 300.0), (1, 2, 1.4, 300.0), (0, 2, 2.4, 0.0)])
    assert partner.shape[1] == 2
    for i in range(5):
        for e in range(partner.shape[1]):
            j = int(partner[i, e])
            if j >= 0:
                assert i in partner[j].tolist()
    with pytest.raises(ValueError):
        build_partner_table(3, [(0, 1, 1.0, 1.0), (0, 2, 1.0, 1.0)], width=1)


def test_oracle_gradient_matches_finite_differences():
    f, x = _case(40, 8, 1, seed=3)
    args = (f["x_exists"], f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"], f["rows"])
    xd = x.double()
    e, g = O.pair_energy_grad(xd, *args)
    h = 1e-5
    for (i, c) in [(int(f["rows"][0]), 0), (int(f["rows"][3]), 2), (int(f["rows"][7]), 1)]:
        xp, xm = xd.clone(), xd.clone()
        xp[0, i, c] += h
        xm[0, i, c] -= h
        fd = (O.pair_energy(xp, *args) - O.pair_energy(xm, *args)) / (2 * h)
        assert abs(float(fd[0]) - float(g[0, i, c])) < 1e-5 * max(1.0, abs(float(fd[0])))


def test_oracle_rows_all_equals_pair_sum():
    """rows = all atoms: the energy is each unordered pair once (checked against an explicit double loop)."""
    f, x = _case(12, 4, 1, seed=5, missing=False)
    e = O.pair_energy(x.double(), f["x_exists"], f["sigma"], f["eps"], None, None, None, None)
    p = PairEnergyParams()
    tot = 0.0
    xs = x[0].double()
    for i in range(12):
        for j in range(i + 1, 12):
            d2 = float(((xs[i] - xs[j]) ** 2).sum()) + 1e-12
            if d2 - 1e-12 >= p.cutoff ** 2:
                continue
            d = math.sqrt(d2)
            sig = 0.5 * float(f["sigma"][i] + f["sigma"][j])
            ee = math.sqrt(float(f["eps"][i]) * float(f["eps"][j]))
            s6 = (sig * sig / (d2 + p.softcore * sig * sig)) ** 3
            tot += ee * (s6 * s6 - 2 * s6) + p.clash_k * max(0.0, p.clash_scale * sig - d) ** 2
    assert abs(float(e[0]) - tot) < 1e-9 * max(1.0, abs(tot))


def test_oracle_descent_lowers_energy():
    f, x = _case(64, 12, 2, seed=7)
    args = (f["x_exists"], f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"], f["rows"])
    e0 = O.pair_energy(x, *args)
    x1 = O.pair_energy_descend(x, *args, iters=5, step=0.002)
    e1 = O.pair_energy(x1, *args)
    assert bool((e1 < e0).all())
    moved = (x1 - x).abs().sum(-1) > 0
    assert not bool(moved[:, ~f["in_rows"]].any())


@pytest.mark.gpu
@pytest.mark.parametrize("Na,n_lig,B,rows_all", [(318, 29, 4, False), (2048, 32, 16, False), (200, 10, 3, True),
                                                 (3072, 50, 2, False)])
def test_pair_energy_grad_vs_autograd(Na, n_lig, B, rows_all):
    from physdock_b200.physics import PairEnergyField
    dev = torch.device("cuda", 0)
    f, x = _case(Na, n_lig, B, seed=Na)
    rows = None if rows_all else f["rows"]
    fld = PairEnergyField(f["x_exists"].to(dev), f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"],
                          rows=rows)
    energy, grad = fld.energy_grad(x.to(dev))
    want_e, want_g = O.pair_energy_grad(x.double(), f["x_exists"], f["sigma"], f["eps"], f["partner"], f["partner_r0"],
                                        f["partner_k"], rows)
    # fp32 pair terms summed over <= Na pairs: 1e-5 of the summed magnitude
    rel_close("energy", energy, want_e, rtol=2e-5, atol=2e-5 * float(want_e.abs().max()))
    sel = torch.arange(Na) if rows_all else f["rows"].long()
    gscale = float(want_g[:, sel].abs().max())
    rel_close("grad", grad[:, sel], want_g[:, sel], rtol=2e-5, atol=2e-5 * gscale)
    if not rows_all:
        assert float(grad[:, ~f["in_rows"].to(dev)].abs().max()) == 0.0


@pytest.mark.gpu
def test_descend_vs_oracle_and_idempotent_inputs():
    from physdock_b200.physics import PairEnergyField
    dev = torch.device("cuda", 0)
    f, x = _case(512, 24, 4, seed=11)
    args = (f["x_exists"], f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"], f["rows"])
    fld = PairEnergyField(f["x_exists"].to(dev), f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"],
                          rows=f["rows"])
    xd = x.to(dev)
    keep = xd.clone()
    got = fld.descend(xd, iters=5, step=0.002, gmax=50.0)
    assert torch.equal(xd, keep)                                  # the input is not modified
    want = O.pair_energy_descend(x.double(), *args, iters=5, step=0.002, gmax=50.0)
    assert float(O.rmsd(got.cpu().double(), want).max()) < 1e-4
    assert torch.equal(got[:, ~f["in_rows"].to(dev)], xd[:, ~f["in_rows"].to(dev)])
    e0, _ = fld.energy_grad(xd)
    e0 = e0.clone()
    e1, _ = fld.energy_grad(got)
    assert bool((e1 < e0).all())
    again = fld.descend(xd, iters=5, step=0.002, gmax=50.0)       # deterministic: bit-identical on a second run
    assert torch.equal(again, got)
