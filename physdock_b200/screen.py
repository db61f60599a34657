"""Virtual-screening driver around the sampling hot path (reference screening.py:100-340).

The reference loops over a SMILES library in ONE process: featurise the receptor + ligand (`remove_ligand=True, smi=...`,
screening.py:106-114), move the tensors to the GPU, run the 102 M-parameter trunk and then `sample_diffusion` -- strictly one
after the other, the trunk being recomputed for every ligand because ligand tokens take part in the pair stack
(diffusion_conditioning.py:232-238; no exact receptor cache is possible).  Measured on a B200 at Nt=256 / Na=2048
(profiles/r02_trunk_vs_sampling.txt): trunk 131-188 ms per ligand in PyTorch eager against 78 ms for the B200-native sampling
of 8 poses x 40 steps, i.e. the out-of-scope trunk is now ~63 % of a ligand's GPU time.

What this driver does about it without touching the trunk's arithmetic:
  * ligands are sharded round-robin over the ranks (`sharding.shard_ligands`; all poses of a ligand stay on one GPU);
  * inside a rank the NEXT ligand is featurised, staged through pinned memory and run through the trunk on a side stream
    while the CURRENT ligand is being sampled (`pipeline.prefetch_complexes`), so a ligand costs max(trunk, sampling)
    instead of their sum once the pipeline is full;
  * results are collected with ONE `all_gather_object` per library (poses are 24.6 KB per sample).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

from .pipeline import prefetch_complexes
from .sharding import shard_ligands, world


@torch.no_grad()
def screen_ligands(model, ligands: Sequence[Any], featurise: Callable[[Any], Optional[Dict[str, torch.Tensor]]],
                   num_sample: int = 8, steps: int = 40, karras_noise_schedule_power: float = 1000, overlap: bool = True,
                   gather: bool = True, on_result: Optional[Callable[[int, Any, torch.Tensor], None]] = None,
                   **sample_kw) -> List[Optional[torch.Tensor]]:
    """Samples `num_sample` poses for every ligand of the library.

    model: a PhysDockB200 (`.diffusion_conditioning(batch) -> (a, ap, s, z)`, `.sample_diffusion(batch, conditioning=...)`);
    featurise(ligand) -> CPU feature dict (the reference's `feature_loader.load(..., remove_ligand=True, smi=ligand)`), or
    None when featurisation fails (the reference prints and skips, screening.py:115-118).
    Returns, on every rank when `gather`, the list of pose tensors [num_sample, Na_i, 3] (CPU) in library order."""
    rank, ws = world()
    dev = next(model.dit.parameters()).device
    mine = shard_ligands(len(ligands), rank, ws)

    def systems():
        for i in mine:
            yield i, featurise(ligands[i])

    def sample(i, batch, cond):
        x = model.sample_diffusion(batch, num_sample=num_sample, steps=steps,
                                   karras_noise_schedule_power=karras_noise_schedule_power, conditioning=cond, **sample_kw)
        x_cpu = x.cpu()
        if on_result is not None:
            on_result(i, ligands[i], x_cpu)
        return x_cpu

    local: List[Tuple[int, torch.Tensor]] = []
    if overlap:
        for i, batch, cond in prefetch_complexes(systems(), dev, conditioning_fn=model.diffusion_conditioning, depth=1):
            local.append((i, sample(i, batch, cond)))
    else:           # the reference's serial order (kept for A/B timing)
        for i, tensors in systems():
            if tensors is None:
                continue
            batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in tensors.items()}
            local.append((i, sample(i, batch, model.diffusion_conditioning(batch))))
    if not gather or ws == 1:
        parts = [local]
    else:
        import torch.distributed as dist
        parts = [None] * ws
        dist.all_gather_object(parts, local)
    out: List[Optional[torch.Tensor]] = [None] * len(ligands)
    for part in parts:
        for i, x in part:
            out[i] = x
    return out
