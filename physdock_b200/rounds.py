"""Multi-round redocking driver -- the direct caller of the sampling hot path, mirroring
reference redocking.py:181-335 (round loop, chirality accept/reject, adaptive MMFF boundary, template hand-over) and
redocking.py:339-423 (alignment of the accepted poses onto the ground-truth pocket, KMeans ranking).

What stays on the host, exactly as in the reference: the accept/reject predicate (RDKit chirality check through a PDB
text round trip, redocking.py:264-281,303-308 -- third-party, passed in as `accept_fn`; `rdkit_chirality_predicate` builds
the reference's own predicate when RDKit is importable) and the list bookkeeping.  What runs on the GPU: every
`sample_diffusion` round, the conformer-template ranking (pdk_template_select's epsilon kernel), the final weighted Kabsch
alignment (pdk_rigid_align) and the pose-RMSD matrix (pdk_pairwise_rmsd).  File output (PDB/SDF writers) is the caller's.

Sharded mode (one process per GPU): every round each rank samples its share of `num_augmentation_sample`; ONE
`all_gather_into_tensor` per round brings all poses to every rank, and every rank runs the same (deterministic)
bookkeeping, so no further exchange is needed.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import torch

from . import _lib
from .ranking import rank_conformer_templates, rank_poses, update_mmff_factor


@dataclass
class RoundRecord:
    recycle_id: int
    mmff_gamma_0_factor: float                 # the factor this round was sampled with
    x_pred: torch.Tensor                       # [num_augmentation_sample, Na, 3] (device)
    pass_flags: List[bool]
    n_templates: int                           # ligand + reference templates handed to this round
    used_inds: Optional[torch.Tensor] = None   # conformers chosen as reference templates for the NEXT round


@dataclass
class RoundsResult:
    accept_samples: torch.Tensor               # [S, Na, 3], S <= max_samples (accepted first, then rejected fill-ins)
    n_accepted: int
    aligned: Optional[torch.Tensor]            # accept_samples aligned onto x_gt's pocket frame (redocking.py:341-342)
    ranking_ids: Optional[List[int]]           # redocking.py:412-423
    rmsd_matrix: Optional[torch.Tensor]
    final_factor: float
    rounds: List[RoundRecord] = field(default_factory=list)


def rdkit_chirality_predicate(ref_mol, ref_pos_pdb_block: str, pdb_block_of: Callable[[torch.Tensor], str]):
    """The reference's accept test (redocking.py:231-238,264-281): the chiral centres RDKit perceives in the predicted ligand
    (through a PDB block) must equal those of the reference conformer.  Needs RDKit; `pdb_block_of(x_cpu)` is the caller's
    `feature_loader.write_pdb_block(x_cpu, infer_meta_data=..., ligand_only=True)`."""
    try:
        from rdkit import Chem
    except Exception as e:  # pragma: no cover - rdkit is absent from this image
        raise _lib.PdkError("rdkit is not importable: pass your own accept_fn (or None to accept every pose)") from e
    init = {i[0]: i[1] for i in Chem.FindMolChiralCenters(ref_mol)}                                    # pragma: no cover
    gt = {i[0]: i[1] for i in Chem.FindMolChiralCenters(Chem.MolFromPDBBlock(ref_pos_pdb_block, sanitize=False))}  # pragma: no cover
    centres = {k: v for k, v in gt.items() if k in init}                                               # pragma: no cover

    def accept(x_cpu: torch.Tensor) -> bool:                                                           # pragma: no cover
        try:
            new = {i[0]: i[1] for i in Chem.FindMolChiralCenters(Chem.MolFromPDBBlock(pdb_block_of(x_cpu), sanitize=False))}
        except Exception:
            return False
        return all(c in new and new[c] == v for c, v in centres.items())

    return accept                                                                                      # pragma: no cover


def pocket_alignment_weights(batch: Dict[str, torch.Tensor], use_pocket: bool = True) -> torch.Tensor:
    """redocking.py:198-201 (`align_mode == "pocket_ca"`): protein atoms of the pocket residues."""
    w = (batch["s_mask"] * batch["is_protein"])[batch["atom_id_to_token_id"].long()]
    if use_pocket:
        w = batch["pocket_res_feat"][batch["atom_id_to_token_id"].long()] * w
    return w


@torch.no_grad()
def run_rounds(model, batch: Dict[str, torch.Tensor], num_augmentation_sample: int = 20, max_samples: int = 40,
               max_rounds: int = 5, steps: int = 40, physics_correction: bool = False,
               mmff_gamma_0_factor_start: float = 6.0, mmff_iters: int = 5, karras_noise_schedule_power: float = 1000,
               conformers: Optional[torch.Tensor] = None, accept_fn: Optional[Callable[[torch.Tensor], bool]] = None,
               ref_mol=None, ref_mol_num_error: bool = False, use_x_gt_ligand_as_ref_pos: bool = False,
               align_weights: Optional[torch.Tensor] = None, ranking: bool = True, num_clusters: int = 5,
               sharded: bool = False, seed: Optional[int] = None, **sample_kw) -> RoundsResult:
    """One system of redocking.py:163-423 given `model` (a PhysDockB200: `.sample_diffusion`, `.dit`,
    `.diffusion_conditioning`) and the feature dict `batch` on the GPU.

    conformers [C, n_lig, 3]: the RDKit `EmbedMultipleConfs` pool of redocking.py:241-258 (generated by the caller; needed
    when physics_correction is on).  accept_fn(x_cpu [Na,3]) -> bool: the chirality predicate (None: accept everything,
    i.e. `pass_flag = True`, the reference's behaviour without physics_correction).  sample_kw: extra keywords for
    `sample_diffusion` (e.g. physics_field=..., mmff_fn=...).
    """
    dev = batch["x_gt"].device
    is_ligand_atom = batch["is_ligand"][batch["atom_id_to_token_id"]].bool()
    x_gt = batch["x_gt"][None]
    if physics_correction and conformers is None and not use_x_gt_ligand_as_ref_pos:
        raise _lib.PdkError("physics_correction needs the conformer pool (redocking.py:241-258 builds it with RDKit "
                            "EmbedMultipleConfs): pass conformers=[C, n_lig, 3]")
    ref_mol_poses = None if conformers is None else conformers.to(dev).float()[:, :int(is_ligand_atom.sum())].contiguous()
    ref_dist = None if ref_mol_poses is None else torch.norm(ref_mol_poses[:, :, None] - ref_mol_poses[:, None], dim=-1).contiguous()

    accept_samples: List[torch.Tensor] = []
    reject_samples: deque = deque([], maxlen=max_samples)
    ligand_templates: List[torch.Tensor] = []
    reference_templates: List[torch.Tensor] = []
    factor = mmff_gamma_0_factor_start
    records: List[RoundRecord] = []
    for recycle_id in range(max_rounds):
        if recycle_id > 0 and not physics_correction:
            break
        if recycle_id >= 1 and "batch_msa_feat" in batch:
            batch["msa_feat"] = batch["batch_msa_feat"][recycle_id]                    # redocking.py:187-188
        if use_x_gt_ligand_as_ref_pos:
            templates = x_gt[:, is_ligand_atom]
        elif recycle_id > 0:
            templates = torch.stack(ligand_templates + reference_templates, dim=0)
        else:
            templates = None
        kw = dict(num_sample=num_augmentation_sample, steps=steps, mmff_gamma_0_factor=factor,
                  align_ref_pos=recycle_id > 0, ref_mol=ref_mol if not ref_mol_num_error else None,
                  ref_mol_poses=templates, use_ref_mol_poses=recycle_id != 0 and physics_correction,
                  ode_step_scale_eta=1.0 if not ref_mol_num_error else 1.5, mmff_iters=mmff_iters,
                  karras_noise_schedule_power=karras_noise_schedule_power, **sample_kw)
        if sharded:
            from .sharding import sample_diffusion_sharded
            a, ap, s, z = model.diffusion_conditioning(batch)
            n = kw.pop("num_sample")
            x_pred = sample_diffusion_sharded(model.dit, batch, a, ap, s, z, num_sample=n,
                                              seed=(0 if seed is None else seed) + recycle_id, exact=True, **kw)
        else:
            if seed is not None:
                torch.manual_seed(seed + recycle_id)
            x_pred = model.sample_diffusion(batch, **kw)
        x_pred_cpu = x_pred.cpu()

        pass_flags = []
        for x, x_cpu in zip(x_pred, x_pred_cpu):                                        # redocking.py:302-317
            pass_flag = bool(accept_fn(x_cpu)) if (physics_correction and accept_fn is not None) else True
            pass_flags.append(pass_flag)
            if pass_flag:
                ligand_templates.append(x[is_ligand_atom])
                accept_samples.append(x)
            else:
                reject_samples.append(x)
        rec = RoundRecord(recycle_id, factor, x_pred, pass_flags, 0 if templates is None else templates.shape[0])
        records.append(rec)
        if physics_correction:
            factor = update_mmff_factor(factor, pass_flags)                             # redocking.py:318-322
            if len(accept_samples) >= max_samples:
                break
            # rank the conformer pool against ALL poses of this round, keep the best as next round's reference templates
            rec.used_inds = rank_conformer_templates(x_pred[:, is_ligand_atom], ref_mol_poses, ref_dist,
                                                     max_samples - len(ligand_templates))     # redocking.py:326-335
            reference_templates = [ref_mol_poses[i] for i in rec.used_inds.tolist()]

    n_accepted = len(accept_samples)
    if len(accept_samples) < num_augmentation_sample:                                    # redocking.py:337-338
        accept_samples = accept_samples + list(reject_samples)
    final = torch.stack(accept_samples[:max_samples], dim=0)

    aligned = ids = dist = None
    if align_weights is not None:
        # redocking.py:341-342: weighted_rigid_align(x_gt, x, weights) = x moved into x_gt's frame, one pose at a time
        from .sampler import weighted_rigid_align
        S = final.shape[0]
        aligned = weighted_rigid_align(x_gt.expand(S, -1, -1).contiguous(), torch.ones_like(batch["a_mask"]).float(),
                                       final, align_weights.to(dev).float())
    if ranking:
        poses = (aligned if aligned is not None else final)[:, is_ligand_atom]
        ids, dist = rank_poses(poses, num_clusters)
    return RoundsResult(final, n_accepted, aligned, ids, dist, factor, records)
