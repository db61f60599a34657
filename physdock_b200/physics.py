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

    FUSED_MAX_ROWS = 4096       # one CTA per sample keeps the rows' gradients in shared memory

    def _fused_ok(self) -> bool:
        return self.rows is not None and self.n_rows <= self.FUSED_MAX_ROWS and self.Na <= 8192

    def descend(self, x: torch.Tensor, iters: int = 5, step: float = 0.01, gmax: float = 50.0,
                fused: Optional[bool] = None) -> torch.Tensor:
        """`iters` steps of x <- x - step * clamp(grad E, +-gmax) on the row atoms; returns a new tensor.
        ONE launch for all iterations when the rows are a ligand-sized subset (pdk_pair_descend); `fused=False` forces the
        2-launches-per-iteration path (same bits, used by the tests)."""
        lib = _lib.load()
        B = x.shape[0]
        cur = x.contiguous()
        if iters > 0 and (self._fused_ok() if fused is None else fused):
            out = torch.empty_like(cur)
            p = self.params
            _lib.check(lib.pdk_pair_descend(_lib.ptr(cur), _lib.ptr(self.x_exists), _lib.ptr(self.sigma), _lib.ptr(self.eps),
                                            _lib.ptr(self.partner), _lib.ptr(self.p_r0), _lib.ptr(self.p_k), self.E,
                                            _lib.ptr(self.rows), _lib.ptr(self.in_rows), self.n_rows, p.clash_k, p.clash_scale,
                                            p.cutoff, p.softcore, iters, step, gmax, _lib.ptr(out), B, self.Na,
                                            _lib.stream_ptr(self.dev)), "pair_descend")
            return out
        bufs = [torch.empty_like(cur), torch.empty_like(cur)]
        for it in range(iters):
            _, grad = self.energy_grad(cur, want_energy=False)
            out = bufs[it & 1]
            _lib.check(lib.pdk_descent_update(_lib.ptr(cur), _lib.ptr(grad), _lib.ptr(self.in_rows), step, gmax,
                                              _lib.ptr(out), B, self.Na, _lib.stream_ptr(self.dev)), "descent_update")
            cur = out
        return cur if iters > 0 else cur.clone()

    def launches_per_descend(self, iters: int) -> int:
        return (1 if iters > 0 else 0) if self._fused_ok() else 2 * iters


# ---------------------------------------------------------------------------------------------------------------------
# Parameter sources.  The functional form (csrc/physics.cu) is 12-6 Lennard-Jones with per-atom (sigma, eps), harmonic
# bonds and 1-3 restraints; it is NOT MMFF94 (buffered 14-7, stretch-bend, torsions, charges), so parity against the
# reference's RDKit step stays UNPINNED whichever source fills the tables.
# ---------------------------------------------------------------------------------------------------------------------
# Universal Force Field nonbonded parameters (Rappe et al., J. Am. Chem. Soc. 114 (1992) 10024, table 1): x_I = vdW distance
# in Angstrom (position of the pair minimum), D_I = well depth in kcal/mol.  12-6 LJ has its minimum at 2^(1/6) sigma.
UFF_X_D = {1: (2.886, 0.044), 5: (4.083, 0.180), 6: (3.851, 0.105), 7: (3.660, 0.069), 8: (3.500, 0.060), 9: (3.364, 0.050),
           11: (2.983, 0.030), 12: (3.021, 0.111), 14: (4.295, 0.402), 15: (4.147, 0.305), 16: (4.035, 0.274),
           17: (3.947, 0.227), 19: (3.812, 0.035), 20: (3.399, 0.238), 25: (2.961, 0.013), 26: (2.912, 0.013),
           27: (2.872, 0.014), 28: (2.834, 0.015), 29: (3.495, 0.005), 30: (2.763, 0.124), 34: (4.205, 0.291),
           35: (4.189, 0.251), 53: (4.500, 0.339)}
UFF_DEFAULT = (3.851, 0.105)                      # unknown elements are treated as carbon (stated fallback)
BOND_K, ANGLE13_K = 300.0, 60.0                   # kcal/mol/A^2: generic stretch constant / 1-3 restraint (stated defaults)


def uff_sigma_eps(atomic_numbers: torch.Tensor):
    """Per-atom (sigma, eps) from the UFF table by atomic number (1-based); sigma = x_I / 2^(1/6)."""
    z = atomic_numbers.long().tolist()
    xs = torch.tensor([UFF_X_D.get(int(k), UFF_DEFAULT)[0] for k in z], dtype=torch.float32)
    ds = torch.tensor([UFF_X_D.get(int(k), UFF_DEFAULT)[1] for k in z], dtype=torch.float32)
    return xs / 2.0 ** (1.0 / 6.0), ds


def bonded_terms_from_geometry(bond_pairs: Sequence[Tuple[int, int]], ref_pos: torch.Tensor, bond_k: float = BOND_K,
                               angle_k: float = ANGLE13_K):
    """(i, j, r0, k) for every bond (r0 = its length in the reference conformer) and every 1-3 pair two bonds apart
    (r0 = their distance in the reference conformer: fixes the valence angle)."""
    nbr = {}
    for i, j in bond_pairs:
        nbr.setdefault(i, set()).add(j)
        nbr.setdefault(j, set()).add(i)
    out, seen = [], set()
    for i, j in bond_pairs:
        key = (min(i, j), max(i, j))
        if key not in seen:
            seen.add(key)
            out.append((key[0], key[1], float(torch.norm(ref_pos[key[0]] - ref_pos[key[1]])), bond_k))
    for c, ns in nbr.items():
        ns = sorted(ns)
        for u in range(len(ns)):
            for v in range(u + 1, len(ns)):
                key = (ns[u], ns[v])
                if key not in seen:
                    seen.add(key)
                    out.append((key[0], key[1], float(torch.norm(ref_pos[key[0]] - ref_pos[key[1]])), angle_k))
    return out


def parameters_from_features(batch, elements: Optional[torch.Tensor] = None):
    """RDKit-free parameter source built from the REFERENCE'S OWN feature tensors (feature_loader.py:146-161,970-998):
      * element of every atom = argmax of the 128-wide one-hot at `ref_feat[:, 4:132]` (atomic number - 1), or `elements`
        ([Na] atomic numbers) when the caller has them -> UFF (sigma, eps);
      * ligand bonds = `token_bonds` between ligand tokens (ligand atoms are one atom per token), bond lengths and 1-3
        distances from `ref_pos`, the reference conformer the sampler itself aligns to (model.py:183,245);
      * movable rows = the ligand atoms; the receptor only contributes nonbonded terms.
    Returns CPU tensors: dict(sigma, eps, partner, partner_r0, partner_k, rows, bonds).
    """
    a2t = batch["atom_id_to_token_id"].long().cpu()
    is_lig_tok = batch["is_ligand"].cpu() > 0
    is_lig_atom = is_lig_tok[a2t]
    Na = a2t.numel()
    if elements is None:
        elements = batch["ref_feat"][:, 4:132].cpu().argmax(-1) + 1
    sigma, eps = uff_sigma_eps(elements.cpu())
    ref_pos = batch["ref_pos"].float().cpu()
    tok2atom = {}
    for at in torch.nonzero(is_lig_atom).flatten().tolist():
        tok2atom.setdefault(int(a2t[at]), []).append(at)
    tb = batch["token_bonds"].cpu() > 0
    pairs = []
    for ti, tj in torch.nonzero(torch.triu(tb, diagonal=1)).tolist():
        if bool(is_lig_tok[ti]) and bool(is_lig_tok[tj]) and len(tok2atom.get(ti, [])) == 1 and len(tok2atom.get(tj, [])) == 1:
            pairs.append((tok2atom[ti][0], tok2atom[tj][0]))
    partner, r0, k = build_partner_table(Na, bonded_terms_from_geometry(pairs, ref_pos))
    rows = torch.nonzero(is_lig_atom).flatten().int()
    return dict(sigma=sigma, eps=eps, partner=partner, partner_r0=r0, partner_k=k, rows=rows, bonds=pairs)


def field_from_features(batch, params: Optional[PairEnergyParams] = None, elements: Optional[torch.Tensor] = None,
                        device=None) -> "PairEnergyField":
    """`parameters_from_features` as a device-resident PairEnergyField (the ligand moves in the field of the whole crop)."""
    f = parameters_from_features(batch, elements)
    dev = device if device is not None else batch["a_mask"].device
    return PairEnergyField(batch["a_mask"].to(dev), f["sigma"], f["eps"], f["partner"], f["partner_r0"], f["partner_k"],
                           rows=f["rows"], params=params)


def field_from_rdkit_mol(batch, ref_mol, params: Optional[PairEnergyParams] = None, device=None) -> "PairEnergyField":
    """MMFF94-typed parameter source when RDKit is importable (rdkit==2024.3.3, enviroment.yaml:33): ligand (sigma, eps) from
    `MMFFGetMoleculeProperties(...).GetMMFFVdWParams(i, i)` (R*_ii, eps_ii), bond constants from `GetMMFFBondStretchParams`
    (k = 143.9325 / 2 * kb kcal/mol/A^2, r0), 1-3 distances from `GetMMFFAngleBendParams` (theta0, law of cosines); receptor
    atoms keep the UFF table.  Still the 12-6 / harmonic functional form: parity vs MMFFOptimizeMolecule UNPINNED."""
    try:
        from rdkit.Chem import AllChem
    except Exception as e:  # pragma: no cover - rdkit is absent from this image
        raise _lib.PdkError("rdkit is not importable: use field_from_features (UFF table + reference-conformer geometry)") from e
    import math                                                                       # pragma: no cover
    base = field_from_features(batch, params=params, device="cpu" if device is None else device)   # pragma: no cover
    rows = base.rows.cpu().long()                                                    # pragma: no cover
    props = AllChem.MMFFGetMoleculeProperties(ref_mol, mmffVariant="MMFF94")         # pragma: no cover
    sigma, eps = base.sigma.cpu().clone(), base.eps.cpu().clone()                    # pragma: no cover
    n = min(ref_mol.GetNumAtoms(), rows.numel())                                     # pragma: no cover
    for i in range(n):                                                               # pragma: no cover
        vdw = props.GetMMFFVdWParams(i, i)
        if vdw:
            sigma[rows[i]], eps[rows[i]] = vdw[2] / 2.0 ** (1.0 / 6.0), vdw[3]
    terms = []                                                                       # pragma: no cover
    for bnd in ref_mol.GetBonds():                                                   # pragma: no cover
        i, j = bnd.GetBeginAtomIdx(), bnd.GetEndAtomIdx()
        bs = props.GetMMFFBondStretchParams(ref_mol, i, j)
        if bs and i < n and j < n:
            terms.append((int(rows[i]), int(rows[j]), float(bs[2]), 143.9325 / 2.0 * float(bs[1])))
    for atom in ref_mol.GetAtoms():                                                  # pragma: no cover
        c = atom.GetIdx()
        nb = [a.GetIdx() for a in atom.GetNeighbors()]
        for u in range(len(nb)):
            for v in range(u + 1, len(nb)):
                ab = props.GetMMFFAngleBendParams(ref_mol, nb[u], c, nb[v])
                b1, b2 = props.GetMMFFBondStretchParams(ref_mol, nb[u], c), props.GetMMFFBondStretchParams(ref_mol, c, nb[v])
                if ab and b1 and b2 and max(nb[u], nb[v], c) < n:
                    th = math.radians(float(ab[2]))
                    r13 = math.sqrt(b1[2] ** 2 + b2[2] ** 2 - 2 * b1[2] * b2[2] * math.cos(th))
                    terms.append((int(rows[nb[u]]), int(rows[nb[v]]), r13, ANGLE13_K))
    partner, r0, k = build_partner_table(base.Na, terms)                             # pragma: no cover
    dev = device if device is not None else batch["a_mask"].device                   # pragma: no cover
    return PairEnergyField(batch["a_mask"].to(dev), sigma, eps, partner, r0, k, rows=rows.int(), params=params)   # pragma: no cover
