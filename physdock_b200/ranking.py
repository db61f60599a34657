"""Per-round acceptance bookkeeping and final pose ranking of the redocking driver, the direct
caller of the sampling hot path: reference redocking.py:302-335 (template ranking, adaptive MMFF factor) and
redocking.py:357-423 (pairwise ligand RMSD matrix + KMeans representatives).

The O(S^2 n) distance matrix and the O(B C n^2) template scores run on the GPU (pdk_pairwise_rmsd, pdk_template_select's
epsilon kernel); the clustering itself stays scikit-learn's `KMeans(n_clusters, random_state=0)` exactly as the reference
calls it (redocking.py:399), on the host -- it sees an S x S matrix with S <= 40-64.
Chirality filtering (redocking.py:264-281,306-308) needs RDKit and is left to the caller (`pass_flags`).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def pairwise_rmsd(ligand_poses: torch.Tensor) -> torch.Tensor:
    """[S,n,3] fp32 CUDA -> dist [S,S] fp64 (redocking.py:391)."""
    lib = _lib.load()
    S, n, _ = ligand_poses.shape
    poses = ligand_poses.float().contiguous()
    dist = torch.empty(S, S, dtype=torch.float64, device=poses.device)
    _lib.check(lib.pdk_pairwise_rmsd(_lib.ptr(poses), _lib.ptr(dist), S, n, _lib.stream_ptr(poses.device)), "pairwise_rmsd")
    return dist


def get_representatives(distance_matrix: np.ndarray, num_clusters: int = 5) -> List[int]:
    """redocking.py:393-410: KMeans on the rows of the distance matrix, then per cluster the member with the smallest
    mean distance to the cluster's members."""
    from sklearn.cluster import KMeans
    coordinates = np.array(distance_matrix, dtype=np.float64)
    labels = KMeans(n_clusters=num_clusters, random_state=0).fit(coordinates).labels_
    reps = []
    for cluster_id in range(num_clusters):
        idx = np.where(labels == cluster_id)[0]
        avg = np.mean(distance_matrix[idx, :], axis=0)
        reps.append(int(idx[np.argmin(avg[idx])]))
    return reps


def rank_poses(ligand_poses: torch.Tensor, num_clusters: int = 5) -> Tuple[List[int], torch.Tensor]:
    """Final ranking (redocking.py:412-423): the medoid of the whole set first, then the cluster representatives.
    Returns (ids, dist)."""
    dist_t = pairwise_rmsd(ligand_poses)
    dist = dist_t.cpu().numpy()
    if len(dist) > num_clusters:
        ids = get_representatives(dist, num_clusters)
        ids_1 = get_representatives(dist, 1)[0]
        if ids_1 in ids:
            ids.remove(ids_1)
            ids = [ids_1] + ids
        else:
            ids = [ids_1] + ids[:num_clusters - 1]
    else:
        ids = list(range(len(dist)))
    return ids, dist_t


def rank_conformer_templates(ligand_poses: torch.Tensor, ref_mol_poses: torch.Tensor, ref_dist: Optional[torch.Tensor],
                             n_keep: int) -> torch.Tensor:
    """redocking.py:326-335: score every conformer template against ALL predicted ligand poses of the round
    (mean over samples of the smooth-lDDT mismatch, i.e. `epsilon.mean(dim=[-1,-2,-4])`) and keep the best `n_keep`.
    ligand_poses [B,n,3], ref_mol_poses [C,n,3] -> indices [<= n_keep] (int64, ascending score)."""
    from .sampler import template_select
    dev = ligand_poses.device
    B, n, _ = ligand_poses.shape
    if ref_dist is None:
        ref_dist = torch.norm(ref_mol_poses[:, :, None] - ref_mol_poses[:, None], dim=-1).contiguous()
    lig_idx = torch.arange(n, dtype=torch.int32, device=dev)
    scratch = torch.zeros(B, n, 3, dtype=torch.float32, device=dev)
    eps, _ = template_select(ligand_poses.float().contiguous(), lig_idx, ref_dist.float().contiguous(),
                             ref_mol_poses.float().contiguous(), scratch)
    return torch.argsort(eps.mean(dim=0))[:max(n_keep, 0)]


def update_mmff_factor(factor: float, pass_flags: Sequence[bool]) -> float:
    """redocking.py:318-322."""
    return factor * 1.15 if any(pass_flags) else max(factor * 0.7, 1)
