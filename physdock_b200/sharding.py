"""Multi-GPU sample sharding: one process per GPU, independent samples, no data-path
collective; ONE all_gather of the final coordinates so rank 0 can rank/cluster them
(reference: single process, redocking.py:300,357-447).

Parity with a single-GPU run of `num_sample` samples: every rank can draw the random tensors for ALL samples
from the same seed and keep its slice (`ShardedRNG`), so the union over ranks is bit-identical to one rank
drawing everything.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(num_sample: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the sample axis owned by `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(num_sample, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_samples(x_local: torch.Tensor) -> torch.Tensor:
    """all_gather along the sample axis (ragged shards allowed).  NCCL on GPUs, gloo in the CPU tests."""
    rank, ws = world()
    if ws == 1:
        return x_local
    n = torch.tensor([x_local.shape[0]], device=x_local.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n)
    counts = [int(c) for c in counts]
    m = max(counts)
    pad = torch.zeros((m,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


class ShardedRNG:
    """Draws each random tensor for all `num_sample` samples and returns this rank's rows, so a sharded run
    consumes the generator exactly like a single-process run (leading dimension must be the sample axis)."""

    def __init__(self, base_rng, num_sample: int, rank: int, world_size: int):
        self.base, self.n = base_rng, num_sample
        self.lo, self.hi = shard_range(num_sample, rank, world_size)

    def _cut(self, fn, shape):
        shape = tuple(shape)
        assert shape[0] == self.hi - self.lo, (shape, self.lo, self.hi)
        return fn((self.n,) + shape[1:])[self.lo:self.hi].contiguous()

    def rand(self, shape):
        return self._cut(self.base.rand, shape)

    def normal(self, shape):
        return self._cut(self.base.normal, shape)


def sample_diffusion_sharded(dit, batch, a, ap, s, z, num_sample: int, seed: int = 0, exact: bool = True, **kw):
    """Runs `sample_diffusion` for this rank's share of `num_sample` samples and returns ALL samples on every
    rank.  exact=True reproduces the single-process random stream (ShardedRNG); exact=False seeds each rank
    with seed+rank (cheaper: no redundant draws)."""
    from .sampler import DeviceRNG, sample_diffusion
    rank, ws = world()
    lo, hi = shard_range(num_sample, rank, ws)
    dev = batch["x_gt"].device
    if exact:
        torch.manual_seed(seed)
        rng = ShardedRNG(DeviceRNG(dev), num_sample, rank, ws)
    else:
        torch.manual_seed(seed + rank)
        rng = DeviceRNG(dev)
    if kw.get("ref_mol_poses") is not None:
        pass    # templates are shared by all samples: nothing to shard
    x_local = sample_diffusion(dit, batch, a, ap, s, z, num_sample=hi - lo, rng=rng, **kw)
    return gather_samples(x_local)


# ---------------------------------------------------------------------------------------------------------------------
# screening (reference screening.py:106-119 loops over the ligand library in ONE process)
# ---------------------------------------------------------------------------------------------------------------------
def shard_ligands(n_ligands: int, rank: int, world_size: int):
    """Round-robin split of a ligand library: rank r screens ligands r, r+ws, r+2ws, ...; all `num_sample` poses of one
    ligand stay on one GPU (they share that ligand's trunk outputs and pair-bias cache)."""
    return list(range(rank, n_ligands, world_size))


def screen_library(n_ligands: int, sample_ligand, n_atoms_of=None):
    """Runs `sample_ligand(ligand_index) -> x [num_sample, n_atoms, 3]` for this rank's ligands and returns, on every
    rank, the list of all ligands' coordinates in library order.  No collective on the data path; ONE all_gather_object
    of the (small) final coordinates per library.  `n_atoms_of` is unused (kept for symmetric signatures)."""
    rank, ws = world()
    mine = shard_ligands(n_ligands, rank, ws)
    local = [(i, sample_ligand(i).detach().cpu()) for i in mine]
    if ws == 1:
        gathered = [local]
    else:
        gathered = [None] * ws
        dist.all_gather_object(gathered, local)
    out = [None] * n_ligands
    for part in gathered:
        for i, x in part:
            out[i] = x
    return out
