"""Round driver (reference redocking.py:181-423): host bookkeeping on CPU, the full driver on the GPU."""
import pytest
import torch

from oracle import physdock_oracle as O
from physdock_b200.synthetic import make_templates
from tests.helpers import complex_64_512, medium_state, log_value


def handedness_predicate(lig_idx):
    """A chirality-like accept test that needs no RDKit: the sign of the signed volume spanned by four ligand atoms."""
    i0, i1, i2, i3 = [int(i) for i in lig_idx[:4]]

    def accept(x):
        a, b, c = x[i1] - x[i0], x[i2] - x[i0], x[i3] - x[i0]
        return float(torch.dot(a.double(), torch.linalg.cross(b.double(), c.double()))) > 0.0

    return accept


def test_oracle_rounds_bookkeeping_rules():
    """redocking.py:302-338 restated: all-rejected rounds shrink the boundary (floor 1) and the rejects fill the output."""
    g = torch.Generator().manual_seed(0)
    Na, n_lig = 30, 6
    lig = torch.zeros(Na, dtype=torch.bool)
    lig[-n_lig:] = True
    confs = torch.randn(10, n_lig, 3, generator=g)
    x_preds = [torch.randn(4, Na, 3, generator=g) for _ in range(3)]
    rounds, final, n_acc, factor = O.rounds_bookkeeping(x_preds, lig, confs, lambda x: False, 4, 8, True, 6.0)
    assert n_acc == 0 and [r["pass_flags"] for r in rounds] == [[False] * 4] * 3
    assert abs(factor - max(max(max(6.0 * 0.7, 1) * 0.7, 1) * 0.7, 1)) < 1e-12
    assert final.shape == (8, Na, 3) and torch.equal(final[0], x_preds[1][0])      # deque(maxlen=8) keeps the last 8 rejects
    assert all(len(r["used_inds"]) == 8 for r in rounds)
    # everything accepted: boundary grows, the loop stops once max_samples poses are in
    rounds, final, n_acc, factor = O.rounds_bookkeeping(x_preds, lig, confs, lambda x: True, 4, 8, True, 6.0)
    assert len(rounds) == 2 and rounds[-1]["stop"] and n_acc == 8 and abs(factor - 6.0 * 1.15 * 1.15) < 1e-12
    assert len(rounds[0]["used_inds"]) == 4 and rounds[1]["n_templates"] == 8
    # without physics_correction there is exactly one round and no predicate
    rounds, final, n_acc, factor = O.rounds_bookkeeping(x_preds[:1], lig, None, lambda x: False, 4, 8, False, 6.0)
    assert n_acc == 4 and factor == 6.0 and final.shape[0] == 4


@pytest.mark.gpu
@pytest.mark.parametrize("physics", [True, False])
def test_run_rounds_matches_oracle_bookkeeping(physics):
    from physdock_b200.dit import B200DiT
    from physdock_b200.rounds import run_rounds
    from physdock_b200.sampler import PhysDockB200
    dims, sd, _ = medium_state()
    dit = B200DiT.from_state_dict(sd, dims, device="cuda")
    cx = complex_64_512()
    d = {k: v.to("cuda") for k, v in cx.items()}
    model = PhysDockB200(dit, diffusion_conditioning=lambda batch: (batch["a"], batch["ap"], batch["s"], batch["z"]))
    lig = cx["is_ligand"][cx["atom_id_to_token_id"]].bool()
    lig_idx = torch.nonzero(lig).flatten()
    confs = make_templates(cx, 24)
    accept = handedness_predicate(lig_idx)
    w = torch.zeros(cx["x_gt"].shape[0])
    w[:200] = 1.0                                    # stand-in for the pocket-CA weights (redocking.py:198-201)
    res = run_rounds(model, d, num_augmentation_sample=4, max_samples=8, max_rounds=4, steps=8,
                     physics_correction=physics, mmff_gamma_0_factor_start=6.0, conformers=confs if physics else None,
                     accept_fn=accept, align_weights=w, ranking=True, seed=5)
    x_preds = [r.x_pred.cpu() for r in res.rounds]
    assert len(x_preds) == (len(res.rounds) if physics else 1)
    rounds, final, n_acc, factor = O.rounds_bookkeeping(x_preds, lig, confs if physics else None, accept, 4, 8, physics, 6.0)
    assert len(rounds) == len(res.rounds)
    for mine, ref in zip(res.rounds, rounds):
        assert mine.pass_flags == ref["pass_flags"] and mine.mmff_gamma_0_factor == ref["factor"]
        assert mine.n_templates == ref["n_templates"]
        if ref["used_inds"] is None:
            assert mine.used_inds is None
        else:
            assert torch.equal(mine.used_inds.cpu(), ref["used_inds"])
    assert res.n_accepted == n_acc and res.final_factor == factor
    assert torch.equal(res.accept_samples.cpu(), final)
    if physics:
        assert len(res.rounds) >= 2, "the test should exercise template hand-over"
        assert any(not f for r in res.rounds for f in r.pass_flags) and any(f for r in res.rounds for f in r.pass_flags)
    # final alignment onto the ground-truth frame + ranking vs the reference formulas on the host
    S = final.shape[0]
    want_aligned = torch.stack([O.weighted_rigid_align(cx["x_gt"][None], final[i], w)[0] for i in range(S)])
    r = float(O.rmsd(res.aligned.cpu(), want_aligned).max())
    log_value(f"run_rounds(physics={physics}) aligned poses rmsd", r)
    assert r < 1e-3
    # the pose-RMSD matrix is the GPU part (vs the reference's numpy formula on the same poses); KMeans on 8 unclustered
    # poses flips on last-bit differences of its input, so the host-side ranking rules are checked on the SAME matrix
    dist_gpu = res.rmsd_matrix.cpu().numpy()
    dist = O.pairwise_pose_rmsd(res.aligned.cpu()[:, lig].numpy())
    assert float(abs(dist_gpu - dist).max()) < 1e-6
    assert res.ranking_ids == O.rank_from_distance_matrix(dist_gpu)
