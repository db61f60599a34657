"""Feature hand-off: ordering / skipping / error propagation on CPU, staging + overlap on the GPU."""
import time

import pytest
import torch

from physdock_b200.pipeline import PinnedStager, prefetch_complexes


def _systems(n, fail_at=None, none_at=()):
    for i in range(n):
        if fail_at is not None and i == fail_at:
            raise RuntimeError(f"featurisation of system {i} blew up")
        if i in none_at:
            yield f"sys{i}", None
            continue
        g = torch.Generator().manual_seed(i)
        yield f"sys{i}", {"x": torch.randn(5, 3, generator=g), "idx": torch.arange(4) + i, "name": f"sys{i}"}


def test_prefetch_order_skips_and_errors_on_cpu():
    got = list(prefetch_complexes(_systems(6, none_at=(2,)), "cpu", conditioning_fn=lambda b: (b["x"] * 2,), depth=2))
    assert [m for m, _, _ in got] == ["sys0", "sys1", "sys3", "sys4", "sys5"]
    for meta, batch, cond in got:
        i = int(meta[3:])
        g = torch.Generator().manual_seed(i)
        x = torch.randn(5, 3, generator=g)
        assert torch.equal(batch["x"], x) and torch.equal(cond[0], x * 2) and batch["name"] == meta
    it = prefetch_complexes(_systems(5, fail_at=3), "cpu", depth=1)
    assert [next(it)[0] for _ in range(3)] == ["sys0", "sys1", "sys2"]
    with pytest.raises(RuntimeError, match="system 3"):
        next(it)
    # abandoning the generator early stops the producer thread
    it = prefetch_complexes(_systems(1000), "cpu", depth=1)
    next(it)
    it.close()


@pytest.mark.gpu
def test_pinned_staging_and_trunk_overlap_on_gpu():
    dev = torch.device("cuda", 0)
    st = PinnedStager(dev, depth=2)
    batches = [b for _, b in _systems(5)]
    staged = [st.stage(b) for b in batches]            # more batches than slots: buffers are reused safely
    for b, s in zip(batches, staged):
        d = s.wait()
        assert d["x"].is_cuda and torch.equal(d["x"].cpu(), b["x"]) and torch.equal(d["idx"].cpu(), b["idx"])
        assert d["name"] == b["name"] and s.h2d_bytes == 5 * 3 * 4 + 4 * 8
    assert len(st._slots[0]["buffers"]) == 2, "pinned buffers are cached per (name, shape, dtype)"

    def trunk(b):                                      # stand-in for diffusion_conditioning: enough work to overlap
        y = b["x"].sum() + torch.zeros(2048, 2048, device=dev)
        for _ in range(20):
            y = y @ y * 1e-4
        return (y,)

    def consume(batch, cond):                          # stand-in for sampling
        z = torch.ones(2048, 2048, device=dev)
        for _ in range(20):
            z = z @ z * 1e-4
        return float(cond[0].sum() + z.sum())

    serial = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for meta, b in _systems(6):
        d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
        serial.append(consume(d, trunk(d)))
    t_serial = time.perf_counter() - t0
    piped = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for meta, d, cond in prefetch_complexes(_systems(6), dev, conditioning_fn=trunk, depth=1):
        piped.append(consume(d, cond))
    t_piped = time.perf_counter() - t0
    assert piped == serial, "same results as the serial hand-off"
    print(f"serial {t_serial * 1e3:.1f} ms, pipelined {t_piped * 1e3:.1f} ms")
