"""Multi-GPU sample sharding: one process per GPU, independent samples, no data-path
collective; ONE all_gather of the final coordinates so rank 0 can rank/cluster them
(reference: single process, redocking.py:300,357-447).

Parity with a single-GPU run of `num_sample` samples: every rank can draw the random tensors for ALL samples
from the same seed and keep its slice (`ShardedRNG`), so the union over ranks is bit-identical to one rank
drawing everything.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

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


def warm_communicator(device=None) -> None:
    """Creates the communicator (NCCL: lazy, ~ms on first use) outside any timed or latency-critical region."""
    rank, ws = world()
    if ws == 1:
        return
    t = torch.zeros(1, device=device if device is not None else "cpu")
    dist.all_reduce(t)
    if t.is_cuda:
        torch.cuda.synchronize(t.device)


def gather_samples(x_local: torch.Tensor, counts: Optional[Sequence[int]] = None) -> torch.Tensor:
    """all_gather along the sample axis (ragged and empty shards allowed).  NCCL on GPUs, gloo in the CPU tests.

    `counts` = samples per rank when the caller knows them on the host (`shard_range` does): ONE
    `all_gather_into_tensor` of the padded shards, no count exchange and no host sync.  Without it the counts are
    exchanged first (one extra small collective and a device->host read)."""
    rank, ws = world()
    if ws == 1:
        return x_local
    if counts is None:
        n = torch.tensor([x_local.shape[0]], device=x_local.device, dtype=torch.int64)
        allc = torch.empty(ws, device=x_local.device, dtype=torch.int64)
        dist.all_gather_into_tensor(allc, n)
        counts = [int(c) for c in allc.tolist()]
    counts = list(counts)
    assert len(counts) == ws and counts[rank] == x_local.shape[0], (counts, rank, tuple(x_local.shape))
    m = max(counts)
    tail = tuple(x_local.shape[1:])
    if m == 0:
        return x_local.new_zeros((0,) + tail)
    if all(c == m for c in counts):
        send = x_local.contiguous()
    else:
        send = x_local.new_zeros((m,) + tail)
        send[: x_local.shape[0]] = x_local
    out = x_local.new_empty((ws * m,) + tail)
    dist.all_gather_into_tensor(out, send)
    if all(c == m for c in counts):
        return out
    out = out.view((ws, m) + tail)
    return torch.cat([out[r, :c] for r, c in enumerate(counts)], dim=0)


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
    counts = [shard_range(num_sample, r, ws)[1] - shard_range(num_sample, r, ws)[0] for r in range(ws)]
    if hi > lo:
        x_local = sample_diffusion(dit, batch, a, ap, s, z, num_sample=hi - lo, rng=rng, **kw)
    else:       # num_sample < world_size: this rank has nothing to sample but still takes part in the collective
        x_local = torch.zeros(0, batch["x_gt"].shape[-2], 3, dtype=torch.float32, device=dev)
    return gather_samples(x_local, counts)


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
