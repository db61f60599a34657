"""Feature hand-off between the reference's CPU featurisation and the B200 sampler.

The reference moves every tensor of `FeatureLoader.load`'s output to the GPU with a blocking `.to(device)` from pageable
memory, one system at a time, and only then runs the trunk and the sampler (redocking.py:156-160, screening.py:119-123;
feature_loader.py:1004-1173 builds ~40 tensors, 30-60 MB at crop 256/2048).  With the sampling step at ~2.6 ms that serial
chain (featurise -> copy -> trunk -> sample) leaves the GPU idle between systems.  This module keeps the reference's
DataLoader (worker processes do the featurisation) and overlaps the rest:

  PinnedStager        every tensor goes through a reusable page-locked buffer and ONE async copy stream;
  prefetch_complexes  a producer thread stages system k+1 and runs its once-per-complex trunk (the reference's PyTorch
                      `diffusion_conditioning`, out of scope for kernels) on a side stream while system k is being sampled
                      on the main stream; the consumer only waits on CUDA events.

On a CPU device the same code runs without pinning / streams (used by the host-logic tests).
"""
from __future__ import annotations

import queue
import threading
from dataclasses import dataclass
from typing import Any, Callable, Dict, Iterable, Iterator, Optional, Tuple

import torch


@dataclass
class StagedBatch:
    tensors: Dict[str, torch.Tensor]          # on the target device
    event: Optional["torch.cuda.Event"]       # recorded on the copy stream after the last copy (None on CPU)
    h2d_bytes: int

    def wait(self, stream: Optional["torch.cuda.Stream"] = None) -> Dict[str, torch.Tensor]:
        """Makes `stream` (default: the current stream) wait for the copies; does not block the host."""
        if self.event is not None:
            (stream or torch.cuda.current_stream(next(iter(self.tensors.values())).device)).wait_event(self.event)
        return self.tensors


class PinnedStager:
    """Host -> device staging of feature dicts through reusable pinned buffers on a dedicated copy stream.

    `depth` = number of batches that may be in flight (a pinned buffer is only rewritten after the copy that read it has
    completed: each slot carries its event)."""

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.depth = max(1, depth)
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._slots = [dict(buffers={}, event=None) for _ in range(self.depth)]
        self._next = 0
        self.bytes_staged = 0

    def _pinned(self, slot, name: str, t: torch.Tensor) -> torch.Tensor:
        key = (name, tuple(t.shape), t.dtype)
        buf = slot["buffers"].get(key)
        if buf is None:
            buf = slot["buffers"][key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        return buf

    def stage(self, tensors: Dict[str, Any]) -> StagedBatch:
        """Non-tensor entries are passed through; tensors already on the device are left where they are."""
        if not self.cuda:
            return StagedBatch({k: v for k, v in tensors.items()}, None, 0)
        slot = self._slots[self._next]
        self._next = (self._next + 1) % self.depth
        if slot["event"] is not None:
            slot["event"].synchronize()               # the copies that read this slot's pinned buffers are done
        out, nbytes = {}, 0
        with torch.cuda.stream(self.stream):
            for k, v in tensors.items():
                if not torch.is_tensor(v) or v.is_cuda:
                    out[k] = v
                    continue
                src = v.contiguous()
                pin = self._pinned(slot, k, src)
                pin.copy_(src)                        # pageable -> pinned (host memcpy)
                out[k] = pin.to(self.device, non_blocking=True)
                nbytes += src.numel() * src.element_size()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        slot["event"] = ev
        self.bytes_staged += nbytes
        return StagedBatch(out, ev, nbytes)


_END = object()


def prefetch_complexes(systems: Iterable[Tuple[Any, Optional[Dict[str, torch.Tensor]]]], device,
                       conditioning_fn: Optional[Callable[[Dict[str, torch.Tensor]], tuple]] = None, depth: int = 1,
                       stager: Optional[PinnedStager] = None) -> Iterator[Tuple[Any, Dict[str, torch.Tensor], Optional[tuple]]]:
    """Yields (meta, device_batch, conditioning) in order.  `systems` yields (meta, cpu_tensor_dict) -- e.g. the reference's
    DataLoader over system pickles (redocking.py:110-115) or SMILES (screening.py:100-119); entries whose dict is None
    (featurisation failed, redocking.py:157-158) are skipped.  While the caller works on system k, a producer thread stages
    system k+1 .. k+depth and, when `conditioning_fn` is given, runs the trunk for them on a side stream.  Exceptions of
    the producer are re-raised in the consumer at the position they occurred."""
    device = torch.device(device)
    cuda = device.type == "cuda"
    stager = stager or PinnedStager(device, depth=depth + 1)
    side = torch.cuda.Stream(device=device) if cuda else None
    q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
    stop = threading.Event()

    def put(item) -> bool:
        while not stop.is_set():
            try:
                q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def produce():
        try:
            if cuda:
                torch.cuda.set_device(device)
            for meta, tensors in systems:
                if stop.is_set():
                    return
                if tensors is None:
                    continue
                staged = stager.stage(tensors)
                cond, ev = None, staged.event
                if conditioning_fn is not None:
                    if cuda:
                        with torch.cuda.stream(side):
                            staged.wait(side)
                            with torch.no_grad():
                                cond = conditioning_fn(staged.tensors)
                            ev = torch.cuda.Event()
                            ev.record(side)
                    else:
                        with torch.no_grad():
                            cond = conditioning_fn(staged.tensors)
                if not put((meta, staged.tensors, cond, ev, None)):
                    return
        except BaseException as e:  # noqa: BLE001 -- handed to the consumer
            put((None, None, None, None, e))
        finally:
            put(_END)

    th = threading.Thread(target=produce, name="pdk-prefetch", daemon=True)
    th.start()
    try:
        while True:
            item = q.get()
            if item is _END:
                return
            meta, batch, cond, ev, err = item
            if err is not None:
                raise err
            if ev is not None:
                cur = torch.cuda.current_stream(device)
                cur.wait_event(ev)
                for t in list(batch.values()) + list(cond or ()):
                    if torch.is_tensor(t) and t.is_cuda:
                        t.record_stream(cur)          # allocated on the copy / side stream, consumed on this one
            yield meta, batch, cond
    finally:
        stop.set()
        th.join(timeout=5.0)
