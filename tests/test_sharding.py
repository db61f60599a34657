"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from physdock_b200.sharding import ShardedRNG, gather_samples, shard_range


def test_shard_range_covers_all_samples():
    for n in (1, 5, 16, 40, 41):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


class _CpuRNG:
    def rand(self, shape):
        return torch.rand(list(shape))

    def normal(self, shape):
        return torch.normal(0, 1, size=tuple(shape))


def _worker(rank, ws, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    n = 5                                    # ragged: 3 + 2
    lo, hi = shard_range(n, rank, ws)
    torch.manual_seed(42)
    rng = ShardedRNG(_CpuRNG(), n, rank, ws)
    x0 = rng.normal((hi - lo, 7, 3))         # the per-rank slice of one global draw
    u = rng.rand((hi - lo,))
    local = x0 * 2 + u[:, None, None]        # stand-in for the per-sample computation
    full = gather_samples(local)             # counts exchanged
    counts = [shard_range(n, r, ws)[1] - shard_range(n, r, ws)[0] for r in range(ws)]
    full_known = gather_samples(local, counts)      # host-known counts: one collective
    assert torch.equal(full, full_known)
    # equal shards (no padding path) and an empty shard (num_sample < world_size)
    eq = gather_samples(torch.full((2, 3), float(rank)), [2, 2])
    assert torch.equal(eq, torch.tensor([[0.] * 3] * 2 + [[1.] * 3] * 2))
    lo1, hi1 = shard_range(1, rank, ws)
    one = gather_samples(torch.full((hi1 - lo1, 4, 3), 7.0), [1, 0])
    assert one.shape == (1, 4, 3) and bool((one == 7.0).all())
    none = gather_samples(torch.zeros(0, 4, 3), [0, 0])
    assert none.shape == (0, 4, 3)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    ws, port = 2, 29571
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.manual_seed(42)
    x0 = torch.normal(0, 1, size=(5, 7, 3))
    u = torch.rand([5])
    assert torch.equal(full, x0 * 2 + u[:, None, None])


def _screen_worker(rank, ws, port, q):
    from physdock_b200.sharding import screen_library, shard_ligands
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    seen = []

    def sample_ligand(i):                    # stand-in for trunk + sample_diffusion of ligand i (8 poses, ragged atom counts)
        seen.append(i)
        g = torch.Generator().manual_seed(1000 + i)
        return torch.randn(8, 20 + i, 3, generator=g)

    out = screen_library(7, sample_ligand)
    assert seen == shard_ligands(7, rank, ws)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_screening_covers_library_in_order():
    """BASELINE.json configs[3] (screening, ligands sharded round-robin, 8 samples per ligand stay on one rank)."""
    from physdock_b200.sharding import shard_ligands
    assert shard_ligands(7, 0, 2) == [0, 2, 4, 6] and shard_ligands(7, 1, 2) == [1, 3, 5]
    assert sorted(sum((shard_ligands(1000, r, 8) for r in range(8)), [])) == list(range(1000))
    ws, port = 2, 29573
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_screen_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(out) == 7
    for i, x in enumerate(out):
        g = torch.Generator().manual_seed(1000 + i)
        assert torch.equal(x, torch.randn(8, 20 + i, 3, generator=g))


class _FakeModel:
    """Stand-in for PhysDockB200 on CPU: trunk = identity on a feature, sampler = deterministic function of it."""

    class _Dit(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))

    def __init__(self):
        self.dit = self._Dit()
        self.trunk_calls = []

    def diffusion_conditioning(self, batch):
        self.trunk_calls.append(int(batch["lig"]))
        return (batch["feat"] * 2,)

    def sample_diffusion(self, batch, num_sample, steps, karras_noise_schedule_power, conditioning, **kw):
        return conditioning[0][None].repeat(num_sample, 1, 1) + steps


def _screen_ligands_worker(rank, ws, port, q):
    from physdock_b200.screen import screen_ligands
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    model = _FakeModel()

    def featurise(i):
        if i == 3:
            return None                       # featurisation failure: skipped, like screening.py:115-118
        return {"feat": torch.full((4 + i, 3), float(i)), "lig": torch.tensor(i)}

    for overlap in (True, False):
        model.trunk_calls.clear()
        out = screen_ligands(model, list(range(7)), featurise, num_sample=8, steps=40, overlap=overlap)
        assert model.trunk_calls == [i for i in range(rank, 7, ws) if i != 3]
        assert out[3] is None
        for i, x in enumerate(out):
            if i != 3:
                assert x.shape == (8, 4 + i, 3) and bool((x == 2 * i + 40).all())
    if rank == 0:
        q.put("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_screen_ligands_overlap_equals_serial():
    """physdock_b200.screen.screen_ligands (BASELINE.json configs[3]): round-robin ligand shards, prefetch pipeline == serial."""
    ws, port = 2, 29575
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_screen_ligands_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    assert q.get(timeout=120) == "ok"
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
