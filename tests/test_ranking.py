"""Caller-side ranking (redocking.py:302-335,357-423): oracle logic on CPU, CUDA parity on the GPU."""
import numpy as np
import pytest
import torch

from oracle import physdock_oracle as O            # checker only


def clustered_poses(S=40, n=24, k=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    base = 4.0 * torch.randn(n, 3, generator=g)
    centres = [base + 3.0 * torch.randn(1, 3, generator=g) + 1.5 * torch.randn(n, 3, generator=g) for _ in range(k)]
    sizes = [S // k + (1 if i < S % k else 0) for i in range(k)]
    poses = torch.cat([c[None] + 0.15 * torch.randn(m, n, 3, generator=g) for c, m in zip(centres, sizes)])
    owner = torch.cat([torch.full((m,), i) for i, m in enumerate(sizes)])
    return poses, owner


def test_oracle_ranking_finds_one_pose_per_cluster():
    poses, owner = clustered_poses()
    ids, dist = O.rank_poses(poses.numpy())
    assert dist.shape == (40, 40) and np.allclose(np.diag(dist), 0) and np.allclose(dist, dist.T)
    assert len(ids) == 5 and len(set(ids)) == 5
    assert sorted(int(owner[i]) for i in O.get_representatives(dist, 5)) == [0, 1, 2, 3, 4]
    # redocking.py:417-421: the global medoid goes first; if it is not a cluster representative the LAST one is dropped
    assert ids[0] == O.get_representatives(dist, 1)[0]
    assert len({int(owner[i]) for i in ids}) >= 4
    few, _ = O.rank_poses(poses[:4].numpy())
    assert few == [0, 1, 2, 3]


def test_update_mmff_factor_rules():
    from physdock_b200.ranking import update_mmff_factor
    assert update_mmff_factor(2.0, [False, True]) == pytest.approx(2.3)
    assert update_mmff_factor(2.0, [False, False]) == pytest.approx(1.4)
    assert update_mmff_factor(1.2, []) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("S,n", [(40, 24), (7, 50), (64, 3)])
def test_pairwise_rmsd_and_ranking_vs_oracle(S, n):
    from physdock_b200 import ranking
    poses, _ = clustered_poses(S=S, n=n, seed=S)
    ids, dist = ranking.rank_poses(poses.cuda())
    want_ids, want_dist = O.rank_poses(poses.numpy())
    assert np.allclose(dist.cpu().numpy(), want_dist, rtol=1e-12, atol=1e-12)
    assert ids == want_ids


@pytest.mark.gpu
def test_rank_conformer_templates_vs_oracle():
    from physdock_b200 import ranking
    g = torch.Generator().manual_seed(3)
    n, C, B = 29, 48, 16
    lig = 3.0 * torch.randn(n, 3, generator=g)
    templates = lig[None] + torch.linspace(0.05, 1.5, C)[:, None, None] * torch.randn(C, n, 3, generator=g)
    preds = lig[None] + 0.2 * torch.randn(B, n, 3, generator=g)
    got = ranking.rank_conformer_templates(preds.cuda(), templates.cuda(), None, 10).cpu()
    want = O.rank_conformer_templates(preds, templates, 10)
    assert torch.equal(got, want)
