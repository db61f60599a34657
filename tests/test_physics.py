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


# ------------------------------------------------------------------------------------------- parameter sources (row f4)
def test_parameters_from_reference_features_on_real_ligand():
    """physics.parameters_from_features on the real 5SD5/HWI features (tests/golden/c1_5sd5.npz): UFF (sigma, eps) by element,
    bonds from token_bonds with lengths from the reference conformer, 1-3 restraints, ligand rows."""
    from physdock_b200.physics import parameters_from_features, UFF_X_D
    from tests.helpers import c1_fixture, T
    g, batch, _ = c1_fixture()
    elements = T(g["extra_ref_element"])
    f = parameters_from_features(batch, elements=elements)
    Na = int(g["Na"])
    assert f["rows"].numel() == 29 and f["sigma"].shape == (Na,) and f["eps"].shape == (Na,)
    lig = f["rows"].long()
    assert set(elements[lig].tolist()) <= set(UFF_X_D), "every ligand element has a UFF entry"
    assert float(f["sigma"].min()) > 2.0 and float(f["sigma"].max()) < 4.2 and float(f["eps"].min()) > 0
    bonds = f["bonds"]
    assert len(bonds) >= 28, "a connected 29-atom ligand has at least 28 bonds"
    is_row = torch.zeros(Na, dtype=torch.bool)
    is_row[lig] = True
    assert all(is_row[i] and is_row[j] for i, j in bonds)
    partner, r0, k = f["partner"], f["partner_r0"], f["partner_k"]
    for i in range(Na):                                   # symmetric table; covalent bond lengths; receptor atoms carry no partners
        for e in range(partner.shape[1]):
            j = int(partner[i, e])
            if j < 0:
                continue
            assert is_row[i] and is_row[j]
            back = (partner[j] == i).nonzero().flatten()
            assert back.numel() == 1 and float(r0[j, back[0]]) == float(r0[i, e]) and float(k[j, back[0]]) == float(k[i, e])
            if float(k[i, e]) == 300.0:
                assert 1.0 < float(r0[i, e]) < 2.0, (i, j, float(r0[i, e]))
            else:
                assert 1.8 < float(r0[i, e]) < 3.2, (i, j, float(r0[i, e]))
    # the reference conformer is (by construction) a minimum of the bonded terms
    e_bond = O.pair_energy(batch["ref_pos"][None], batch["a_mask"], f["sigma"], f["eps"], partner, r0, k, f["rows"])
    assert torch.isfinite(e_bond).all()


@pytest.mark.gpu
def test_fused_descend_equals_multi_launch_and_lowers_real_ligand_energy():
    """pdk_pair_descend (one launch for all iterations) == repeated pdk_pair_energy_grad + pdk_descent_update, bit for bit, on
    the synthetic field and on the real 5SD5 ligand with feature-derived parameters; the energy goes down."""
    from physdock_b200.physics import PairEnergyField, field_from_features
    from tests.helpers import c1_fixture, T
    f, x = _case(2048, 32, 16, seed=3, missing=True)
    fld = PairEnergyField(f["x_exists"].cuda(), f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"], rows=f["rows"])
    assert fld.launches_per_descend(5) == 1
    a = fld.descend(x.cuda(), iters=5, step=0.002, gmax=50.0)
    b = fld.descend(x.cuda(), iters=5, step=0.002, gmax=50.0, fused=False)
    assert torch.equal(a, b), float((a - b).abs().max())
    g, batch, _ = c1_fixture()
    d = {k: v.cuda() for k, v in batch.items()}
    fld = field_from_features(d, elements=T(g["extra_ref_element"]))
    gen = torch.Generator().manual_seed(1)
    x0 = (batch["x_gt"][None] + 0.15 * torch.randn(4, int(g["Na"]), 3, generator=gen)).cuda()
    e0, _ = fld.energy_grad(x0)
    e0 = e0.clone()
    x1 = fld.descend(x0, iters=20, step=0.001, gmax=50.0)
    assert torch.equal(x1, fld.descend(x0, iters=20, step=0.001, gmax=50.0, fused=False))
    e1, _ = fld.energy_grad(x1)
    assert bool((e1 < e0).all()), (e0.tolist(), e1.tolist())
    notlig = torch.ones(int(g["Na"]), dtype=torch.bool)
    notlig[fld.rows.long().cpu()] = False
    assert torch.equal(x1[:, notlig.cuda()], x0[:, notlig.cuda()]), "only the ligand moves"
