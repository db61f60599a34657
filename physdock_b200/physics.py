"""Pair-energy physics backend (host side): device-resident replacement of the reference's per-step RDKit MMFF94 call
(`get_next_step_pos`, PhysDock/models/model.py:26-52, used at model.py:252-261).

The reference hands the denoised ligand to RDKit on the CPU for `mmff_iters` minimiser iterations.  RDKit is not part
of the reference tree, so its arithmetic is not reproducible here (parity unpinned); this module is the
opt-in backend BASELINE.json's north_star names: a per-atom-pair soft-core LJ + clash + bond/restraint energy of the
ligand in the field of every atom of the crop, its analytic coordinate gradient, and `mmff_iters` clamped
gradient-descent steps on the ligand atoms -- all on the GPU, no host round trip, CUDA-graph capturable.
Functional form: physdock_b200/csrc/physics.cu; oracle: oracle/physdock_oracle.py:pair_energy (autograd).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


@dataclass
class PairEnergyParams:
    clash_k: float = 10.0        # kcal/mol/A^2-like weight of the clash penalty
    clash_scale: float = 0.6     # clash when d < clash_scale * sigma_ij
    cutoff: float = 10.0         # nonbonded cutoff, Angstrom
    softcore: float = 0.1        # LJ soft-core constant (keeps the energy finite on noisy coordinates)


def build_partner_table(n_atoms: int, bonds: Sequence[Tuple[int, int, float, float]], width: Optional[int] = None):
    """Symmetric per-atom partner table from a list of (i, j, r0, k): partner/r0/k [n_atoms, E] (CPU tensors).

    Every listed pair is excluded from the nonbonded terms; k = 0 lists a pure exclusion (e.g. a 1-3 pair).
    """
    slots = [[] for _ in range(n_atoms)]
    for i, j, r0, k in bonds:
        if i == j:
            raise ValueError("bond with i == j")
        slots[i].append((j, r0, k))
        slots[j].append((i, r0, k))
    need = max([len(s) for s in slots] + [1])
    E = need if width is None else width
    if need > E:
        raise ValueError(f"partner table width {E} < {need} partners of one atom")
    partner = torch.full((n_atoms, E), -1, dtype=torch.int32)
    r0t = torch.zeros(n_atoms, E, dtype=torch.float32)
    kt = torch.zeros(n_atoms, E, dtype=torch.float32)
    for i, s in enumerate(slots):
        for e, (j, r0, k) in enumerate(s):
            partner[i, e], r0t[i, e], kt[i, e] = j, r0, k
    return partner, r0t, kt


class PairEnergyField:
    """Energy / gradient / descent of the `rows` atoms (the ligand) in the field of all `Na` atoms.

    All tensors live on one CUDA device; every method only enqueues kernels on the current stream.
    """

    def __init__(self, x_exists: torch.Tensor, sigma: torch.Tensor, eps: torch.Tensor,
                 partner: Optional[torch.Tensor] = None, partner_r0: Optional[torch.Tensor] = None,
                 partner_k: Optional[torch.Tensor] = None, rows: Optional[torch.Tensor] = None,
                 params: Optional[PairEnergyParams] = None):
        dev = x_exists.device
        if dev.type != "cuda":
            raise _lib.PdkError("PairEnergyField needs CUDA tensors (no CPU fallback)")
        self.dev, self.Na = dev, x_exists.numel()
        self.params = params or PairEnergyParams()
        self.x_exists = x_exists.float().contiguous()
        self.sigma, self.eps = sigma.to(dev).float().contiguous(), eps.to(dev).float().contiguous()
        if partner is None:
            self.E, self.partner, self.p_r0, self.p_k = 0, None, None, None
        else:
            self.E = partner.shape[1]
            self.partner = partner.to(dev).int().contiguous()
            self.p_r0, self.p_k = partner_r0.to(dev).float().contiguous(), partner_k.to(dev).float().contiguous()
        if rows is None:
            self.rows, self.in_rows, self.n_rows = None, None, self.Na
        else:
            self.rows = rows.to(dev).int().contiguous()
            self.n_rows = self.rows.numel()
            self.in_rows = torch.zeros(self.Na, dtype=torch.uint8, device=dev)
            self.in_rows[self.rows.long()] = 1
        self._ws = {}

    def _buffers(self, B: int):
        if B not in self._ws:
            self._ws[B] = (torch.empty(B, self.n_rows, dtype=torch.float32, device=self.dev),
                           torch.empty(B, dtype=torch.float32, device=self.dev),
                           torch.zeros(B, self.Na, 3, dtype=torch.float32, device=self.dev))
        return self._ws[B]

    def energy_grad(self, x: torch.Tensor, want_energy: bool = True):
        """x [B,Na,3] -> (energy [B], grad [B,Na,3]); grad rows of non-row atoms are zero.  Buffers are reused."""
        lib = _lib.load()
        B = x.shape[0]
        e_row, energy, grad = self._buffers(B)
        p = self.params
        _lib.check(lib.pdk_pair_energy_grad(_lib.ptr(x), _lib.ptr(self.x_exists), _lib.ptr(self.sigma), _lib.ptr(self.eps),
                                            _lib.ptr(self.partner), _lib.ptr(self.p_r0), _lib.ptr(self.p_k), self.E,
                                            _lib.ptr(self.rows), _lib.ptr(self.in_rows), self.n_rows, p.clash_k,
                                            p.clash_scale, p.cutoff, p.softcore, _lib.ptr(e_row),
                                            _lib.ptr(energy) if want_energy else None, _lib.ptr(grad), B, self.Na,
                                            _lib.stream_ptr(self.dev)), "pair_energy_grad")
        return energy, grad

    def descend(self, x: torch.Tensor, iters: int = 5, step: float = 0.01, gmax: float = 50.0) -> torch.Tensor:
        """`iters` steps of x <- x - step * clamp(grad E, +-gmax) on the row atoms; returns a new tensor."""
        lib = _lib.load()
        B = x.shape[0]
        cur = x.contiguous()
        bufs = [torch.empty_like(cur), torch.empty_like(cur)]
        for it in range(iters):
            _, grad = self.energy_grad(cur, want_energy=False)
            out = bufs[it & 1]
            _lib.check(lib.pdk_descent_update(_lib.ptr(cur), _lib.ptr(grad), _lib.ptr(self.in_rows), step, gmax,
                                              _lib.ptr(out), B, self.Na, _lib.stream_ptr(self.dev)), "descent_update")
            cur = out
        return cur if iters > 0 else cur.clone()

    def launches_per_descend(self, iters: int) -> int:
        return 2 * iters
