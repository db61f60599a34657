"""Seeded synthetic inputs for the sampling hot path.

There are no trained weights in this image (`params/params.pt` is a Zenodo download) and the
reference pins no results, so every test and the benchmark use:

  * `make_dit_state(dims, seed)`   -- a state dict with the *exact* keys/shapes of the reference
    denoiser `AF3DiT` (PhysDock/models/layers/transformers.py:178-203): fan-in scaled normal weights,
    and -- because the reference zero-initialises every bias (primitives/linear.py:116-118) -- biases
    and norm affine terms perturbed so those code paths carry signal.
  * `make_complex(Nt, Na, seed)`   -- the trunk outputs `a, ap, s, z` plus the `batch` entries the
    sampler reads (feature_loader.py:970-998 contract; inference never pads so masks are all-ones
    unless `ragged=True`).

Everything is generated on the CPU from `torch.Generator().manual_seed(seed)` so the GPU box and this
container see bit-identical inputs (same torch build), then moved to the requested device.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, asdict
from typing import Dict, Optional

import math

import torch


@dataclass(frozen=True)
class DiTDims:
    """Dimension table of the denoiser (PhysDock/configs.py:59-88)."""
    c_a: int = 128
    c_ap: int = 16
    c_s: int = 512
    c_z: int = 128
    no_blocks_atom: int = 3
    no_blocks_dit: int = 12
    sigma_data: float = 16.0
    inf: float = 1e9
    eps: float = 1e-8
    c_t: int = 256          # time embedding width (timestep_embeddings.py:157-160)
    c_hidden: int = 32      # per-head width (attentions.py:223)

    @staticmethod
    def named(model_name: str = "medium") -> "DiTDims":
        table = {"toy": (2, 2), "tiny": (2, 4), "small": (2, 8), "medium": (3, 12), "full": (3, 24)}
        na, nd = table[model_name]
        return DiTDims(no_blocks_atom=na, no_blocks_dit=nd)

    def ffn_hidden(self, c: int) -> int:
        """feed_forward.py:18-25: int(2*4c/3) rounded up to a multiple of 128."""
        h = int(2 * (4 * c) / 3)
        return 128 * ((h + 127) // 128)

    def as_dict(self):
        return asdict(self)


def _block_shapes(prefix: str, c: int, c_pair: int, dims: DiTDims) -> "OrderedDict[str, tuple]":
    hid = dims.ffn_hidden(c)
    p = prefix
    return OrderedDict([
        (f"{p}.attention.norm_s.linear.weight", (3 * c, dims.c_t)),
        (f"{p}.attention.norm_s.linear.bias", (3 * c,)),
        (f"{p}.attention.norm_z.weight", (c_pair,)),
        (f"{p}.attention.norm_z.bias", (c_pair,)),
        (f"{p}.attention.linear_q.weight", (c, c)),
        (f"{p}.attention.linear_k.weight", (c, c)),
        (f"{p}.attention.linear_v.weight", (c, c)),
        (f"{p}.attention.linear_z.weight", (c // dims.c_hidden, c_pair)),
        (f"{p}.attention.norm_q.weight", (dims.c_hidden,)),
        (f"{p}.attention.norm_k.weight", (dims.c_hidden,)),
        (f"{p}.attention.linear_o.weight", (c, c)),
        (f"{p}.attention.linear_o.bias", (c,)),
        (f"{p}.transition.ffn_norm.linear.weight", (3 * c, dims.c_t)),
        (f"{p}.transition.ffn_norm.linear.bias", (3 * c,)),
        (f"{p}.transition.feed_forward.w1.weight", (hid, c)),
        (f"{p}.transition.feed_forward.w2.weight", (c, hid)),
        (f"{p}.transition.feed_forward.w3.weight", (hid, c)),
    ])


def dit_param_shapes(dims: DiTDims = DiTDims()) -> "OrderedDict[str, tuple]":
    """Key -> shape, in the reference module's registration order (transformers.py:192-203)."""
    out: "OrderedDict[str, tuple]" = OrderedDict()
    out["linear_x.weight"] = (dims.c_a, 3)
    out["linear_x.bias"] = (dims.c_a,)
    out["linear_downscale.weight"] = (dims.c_s, dims.c_a)
    out["linear_downscale.bias"] = (dims.c_s,)
    out["linear_upscale.weight"] = (dims.c_a, dims.c_s)
    out["linear_upscale.bias"] = (dims.c_a,)
    out["time_embedder.timestep_embedder.linear_1.weight"] = (dims.c_t, dims.c_t)
    out["time_embedder.timestep_embedder.linear_1.bias"] = (dims.c_t,)
    out["time_embedder.timestep_embedder.linear_2.weight"] = (dims.c_t, dims.c_t)
    out["time_embedder.timestep_embedder.linear_2.bias"] = (dims.c_t,)
    for i in range(dims.no_blocks_atom):
        out.update(_block_shapes(f"atom_dit_encoder.blocks.{i}", dims.c_a, dims.c_ap, dims))
    for i in range(dims.no_blocks_dit):
        out.update(_block_shapes(f"token_dit.blocks.{i}", dims.c_s, dims.c_z, dims))
    for i in range(dims.no_blocks_atom):
        out.update(_block_shapes(f"atom_dit_decoder.blocks.{i}", dims.c_a, dims.c_ap, dims))
    out["norm_r.weight"] = (dims.c_a,)
    out["norm_r.bias"] = (dims.c_a,)
    out["linear_r.weight"] = (3, dims.c_a)
    return out


def make_dit_state(dims: DiTDims = DiTDims(), seed: int = 0, device="cpu",
                   dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights.  One CPU generator, keys visited in registration order."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, shape in dit_param_shapes(dims).items():
        if k.endswith(".weight") and len(shape) == 2:
            fan_in = shape[1]
            std = (1.0 / fan_in) ** 0.5
            if ".norm_s.linear." in k or ".ffn_norm.linear." in k:
                std *= 0.5          # keeps (1+scale) and gate O(1) like a trained AdaLN-Zero
            v = torch.randn(shape, generator=g) * std
        elif k.endswith(".weight"):     # LayerNorm / RMSNorm gains
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                           # every bias (zero in the reference's init): perturb
            v = 0.02 * torch.randn(shape, generator=g)
            if ".norm_s.linear.bias" in k or ".ffn_norm.linear.bias" in k:
                c = shape[0] // 3
                v[2 * c:] += 0.5        # gate offset so the residual branches are not ~0
        sd[k] = v.to(dtype).to(device)
    return sd


def token_layout(Nt: int, Na: int, ragged: bool = False, seed: int = 0):
    """token_id_to_chunk_sizes / is_ligand for Nt tokens over Na atoms.

    Default = the benchmark layout of SURVEY.md section 8d: Nt/8 single-atom ligand tokens at the end, the
    remaining tokens share the other atoms as evenly as possible (9 atoms each at 256/2048).
    `ragged=True` draws variable chunk sizes (1..14) and inserts a few zero-size (UNK) tokens, which the
    reference maps to a zero row in `downscale` (transformers.py:205-212).
    """
    n_lig = max(1, Nt // 8)
    n_res = Nt - n_lig
    atoms_res = Na - n_lig
    assert atoms_res >= n_res > 0
    if not ragged:
        base, rem = divmod(atoms_res, n_res)
        sizes = [base + (1 if i < rem else 0) for i in range(n_res)]
    else:
        g = torch.Generator().manual_seed(seed + 7)
        w = torch.randint(1, 15, (n_res,), generator=g).double()
        zero = torch.randperm(n_res, generator=g)[: max(1, n_res // 16)]
        w[zero] = 0
        raw = torch.floor(w / w.sum() * atoms_res).long()
        raw[w > 0] = torch.clamp(raw[w > 0], min=1)
        # fix the total on the largest chunk
        raw[torch.argmax(raw)] += atoms_res - int(raw.sum())
        assert int(raw.min()) >= 0 and int(raw.sum()) == atoms_res
        sizes = raw.tolist()
    chunk = torch.tensor(sizes + [1] * n_lig, dtype=torch.int64)
    is_lig = torch.cat([torch.zeros(n_res), torch.ones(n_lig)]).float()
    assert int(chunk.sum()) == Na
    return chunk, is_lig


def make_complex(Nt: int, Na: int, dims: DiTDims = DiTDims(), seed: int = 1, device="cpu",
                 ragged: bool = False, mask_holes: bool = False) -> Dict[str, torch.Tensor]:
    """Trunk outputs + batch entries consumed by the hot path.

    Returns a dict with `a [Na,c_a]`, `ap [Na,Na,c_ap]`, `s [Nt,c_s]`, `z [Nt,Nt,c_z]` (all N(0,1), the
    order they are drawn in is fixed) and the batch keys of SURVEY.md Appendix B that the sampler and the
    denoiser read.  `mask_holes=True` zeroes a few atoms/tokens in the masks (the reference supports it
    through `gen_attn_mask`, tensor_utils.py:642-646, even though inference never produces it).
    """
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(Na, dims.c_a, generator=g)
    ap = torch.randn(Na, Na, dims.c_ap, generator=g)
    s = torch.randn(Nt, dims.c_s, generator=g)
    z = torch.randn(Nt, Nt, dims.c_z, generator=g)
    chunk, is_lig = token_layout(Nt, Na, ragged=ragged, seed=seed)
    atom2tok = torch.repeat_interleave(torch.arange(Nt), chunk)
    a_mask = torch.ones(Na)
    s_mask = torch.ones(Nt)
    if mask_holes:
        a_mask[torch.randperm(Na, generator=g)[: max(1, Na // 50)]] = 0
        a_mask[atom2tok >= Nt - max(1, Nt // 8)] = 1        # never hide ligand atoms
        s_mask[torch.randperm(Nt - max(1, Nt // 8), generator=g)[: max(1, Nt // 40)]] = 0
    x_gt = 10.0 * torch.randn(Na, 3, generator=g)
    ref_pos = 2.0 * torch.randn(Na, 3, generator=g)
    out = dict(
        a=a, ap=ap, s=s, z=z,
        ap_mask=a_mask[:, None] * a_mask[None, :],
        z_mask=s_mask[:, None] * s_mask[None, :],
        a_mask=a_mask, x_exists=a_mask.clone(), s_mask=s_mask,
        token_id_to_chunk_sizes=chunk, atom_id_to_token_id=atom2tok,
        is_ligand=is_lig, x_gt=x_gt, ref_pos=ref_pos,
    )
    return {k: v.to(device) for k, v in out.items()}


def make_templates(batch: Dict[str, torch.Tensor], n_templates: int = 40, seed: int = 5,
                   jitter: float = 0.3) -> torch.Tensor:
    """Synthetic conformer templates `ref_mol_poses [C, n_lig, 3]` = ligand x_gt + jitter*N(0,1)
    (stand-in for RDKit EmbedMultipleConfs, redocking.py:241-258, which is absent here)."""
    g = torch.Generator().manual_seed(seed)
    lig = batch["is_ligand"].cpu()[batch["atom_id_to_token_id"].cpu()].bool()
    base = batch["x_gt"].cpu()[lig]
    t = base[None] + jitter * torch.randn(n_templates, base.shape[0], 3, generator=g)
    return t.to(batch["x_gt"].device)


def make_ligand_field(Na: int, n_lig: int, seed: int = 0, missing: bool = True) -> Dict[str, torch.Tensor]:
    """Synthetic input of the pair-energy physics backend (physics.py): a globule of `Na - n_lig` "protein" atoms
    (jittered 1.9 A lattice points nearest the origin) and a bonded chain of `n_lig` ligand atoms (the LAST atoms, as in
    PhysDock crops where ligand tokens come last) threaded through it.  Returns x0 [Na,3], x_exists, per-atom sigma/eps
    (three pseudo-elements), the symmetric partner table of the ligand (1-2 bonds k=300 r0=1.5, 1-3 restraints k=40,
    one pure exclusion), rows (ligand atom indices, int32) and in_rows (bool)."""
    from .physics import build_partner_table
    g = torch.Generator().manual_seed(seed)
    n_prot = Na - n_lig
    side = int(math.ceil((2.2 * n_prot) ** (1 / 3))) + 2
    ax = (torch.arange(side, dtype=torch.float32) - (side - 1) / 2) * 1.9
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1).reshape(-1, 3)
    order = torch.argsort((grid ** 2).sum(-1), stable=True)
    prot = grid[order[:n_prot]] + 0.25 * torch.randn(n_prot, 3, generator=g)
    lig = torch.zeros(n_lig, 3)
    direction = torch.tensor([1.0, 0.3, -0.2])
    for i in range(1, n_lig):
        step = direction + 0.9 * torch.randn(3, generator=g)
        lig[i] = lig[i - 1] + 1.5 * step / step.norm()
    lig = lig - lig.mean(0) + torch.tensor([0.4, -0.3, 0.2])
    x0 = torch.cat([prot, lig], dim=0)
    kind = torch.randint(0, 3, (Na,), generator=g)
    sigma = torch.tensor([3.4, 3.1, 3.8])[kind]
    eps = torch.tensor([0.09, 0.17, 0.05])[kind]
    x_exists = torch.ones(Na)
    if missing and n_prot > 8:
        x_exists[torch.randperm(n_prot, generator=g)[:max(1, n_prot // 50)]] = 0
    bonds = [(n_prot + i, n_prot + i + 1, 1.5, 300.0) for i in range(n_lig - 1)]
    bonds += [(n_prot + i, n_prot + i + 2, 2.45, 40.0) for i in range(n_lig - 2)]
    if n_lig >= 5:
        bonds.append((n_prot, n_prot + 4, 0.0, 0.0))        # exclusion without a restraint
    partner, r0, k = build_partner_table(Na, bonds)
    rows = torch.arange(n_prot, Na, dtype=torch.int32)
    in_rows = torch.zeros(Na, dtype=torch.bool)
    in_rows[n_prot:] = True
    return dict(x0=x0, x_exists=x_exists, sigma=sigma, eps=eps, partner=partner, partner_r0=r0, partner_k=k, rows=rows,
                in_rows=in_rows)


def checksum(t: torch.Tensor) -> float:
    """Order-independent fingerprint used by the golden fixtures to detect input drift."""
    t = t.detach().double().cpu().flatten()
    return float((t * torch.cos(torch.arange(t.numel(), dtype=torch.float64) * 0.37)).sum())
