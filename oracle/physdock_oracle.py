"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU (PyTorch fp32, optionally fp64) restatement of PhysDock's reverse-diffusion sampling hot path,
written as plain functions over a state dict.  Each function cites the reference file:line it follows
(paths relative to the reference repo root).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module, and only as the *checker*; the product
package `physdock_b200/` never does (it fails loudly when its CUDA library is missing).

Pinning: `tests/test_oracle_pin.py` checks every function here against the real reference imported from
/root/reference (oracle/ref_import.py) when that tree is present, and against the golden vectors that
`oracle/make_golden.py` generated FROM THE REAL REFERENCE and committed under tests/golden/.  The
reference itself ships no tests or golden vectors.

Parity unpinned (stated, not hidden): the RDKit MMFF94 step `get_next_step_pos` (models/model.py:26-52)
lives in un-vendored rdkit==2024.3.3 (enviroment.yaml:33), which is not installable here; the oracle
exposes the branch through a user-supplied `mmff_fn` hook and has no pinned vectors for it.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# schedule + RNG stream
# ----------------------------------------------------------------------------------------------
def karras_noise_schedule(num_steps: int = 200, sigma_data: float = 16, s_max: float = 160,
                          s_min: float = 4 * 10e-4, p: float = 7) -> Tensor:
    """models/model.py:117-129.  Note s_min = 4*10e-4 = 4e-3 (sic)."""
    step_indices = torch.arange(num_steps, dtype=torch.float32)
    t_steps = sigma_data * (s_max ** (1 / p) + step_indices / (num_steps - 1) * (
            s_min ** (1 / p) - s_max ** (1 / p))) ** p
    return torch.cat([t_steps, torch.zeros_like(t_steps[:1])])


class TorchRNG:
    """Draws random tensors with the same torch calls, shapes, dtypes and ORDER as the reference
    (x0: model.py:148; per step: tensor_utils.py:549-557 x2, :582; then model.py:77 iff t_cur>gamma_min).
    A recorded stream can be replayed into another implementation (teacher forcing / device parity)."""

    def __init__(self, device="cpu", dtype=torch.float32, record: bool = False):
        self.device, self.dtype = device, dtype
        self.tape: Optional[List[Tensor]] = [] if record else None

    def _rec(self, t):
        if self.tape is not None:
            self.tape.append(t.detach().cpu().clone())
        return t

    def rand(self, shape):
        return self._rec(torch.rand(list(shape), device=self.device, dtype=torch.float32))

    def normal(self, shape):
        return self._rec(torch.normal(0, 1, size=tuple(shape), dtype=self.dtype, device=self.device,
                                      requires_grad=False))


class ReplayRNG:
    """Replays a recorded tape (list of CPU tensors) in order, on any device."""

    def __init__(self, tape: List[Tensor], device="cpu"):
        self.tape, self.pos, self.device = tape, 0, device

    def _next(self, shape):
        t = self.tape[self.pos]
        self.pos += 1
        assert list(t.shape) == list(shape), (t.shape, shape)
        return t.to(self.device)

    rand = _next
    normal = _next


# ----------------------------------------------------------------------------------------------
# a5 / a4: coordinate augmentation and noise
# ----------------------------------------------------------------------------------------------
def uniform_sphere_point(u_phi: Tensor, u_theta: Tensor) -> Tensor:
    """utils/tensor_utils.py:545-562 with the two uniform draws injected."""
    phi = u_phi * 2 * torch.pi
    theta = torch.acos(u_theta * 2 - 1)
    return torch.stack([torch.cos(phi) * torch.sin(theta), torch.sin(phi) * torch.sin(theta),
                        torch.cos(theta)], dim=-1)


def rotation_from_uniforms(u: Tensor) -> Tensor:
    """utils/tensor_utils.py:565-573.  u [..., 4] = (phi0, theta0, phi1, theta1) uniforms in [0,1)."""
    uniform_e0 = uniform_sphere_point(u[..., 0], u[..., 1])
    uniform_e1 = uniform_sphere_point(u[..., 2], u[..., 3])
    e1 = uniform_e1 - uniform_e0 * (uniform_e1 * uniform_e0).sum(dim=-1, keepdim=True)
    e1 = e1 / torch.norm(e1, dim=-1, keepdim=True)
    e0 = uniform_e0
    e2 = torch.cross(e0, e1, dim=-1)
    return torch.stack([e0, e1, e2], dim=-2)


def centre_random_augmentation(x: Tensor, x_exists: Tensor, u: Tensor, trans: Tensor,
                               s: float = 1.0) -> Tensor:
    """utils/tensor_utils.py:576-586 with the randoms injected: u [B,4] uniforms, trans [B,3] normals."""
    mean = torch.sum(x * x_exists[None, :, None], dim=-2, keepdim=True) / torch.sum(x_exists)
    x_aug = x - mean
    R = rotation_from_uniforms(u)
    x_aug = torch.einsum("...ij,...kj->...ki", R, x_aug)
    return x_aug + (s * trans)[..., None, :]


def diffuse(x_cur: Tensor, t_hat: Tensor, t_cur, noise: Tensor, noise_scale_lambda: float = 1.0) -> Tensor:
    """models/model.py:70-85 (t_cur given) with the normal draw injected."""
    ksi = noise_scale_lambda * noise * torch.sqrt(t_hat ** 2 - t_cur ** 2)[..., None, None]
    return x_cur + ksi


# ----------------------------------------------------------------------------------------------
# a7..a15: the denoiser AF3DiT
# ----------------------------------------------------------------------------------------------
def timestep_proj(timesteps: Tensor, dim: int = 256) -> Tensor:
    """primitives/timestep_embeddings.py:35-86 with flip_sin_to_cos=True, downscale_freq_shift=0
    (:160): [cos | sin] of t * exp(-ln(1e4) k / 128)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32,
                                               device=timesteps.device)
    exponent = exponent / (half - 0)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)


def time_embedding(sd: State, timestep: Tensor) -> Tensor:
    """primitives/timestep_embeddings.py:156-166 + :127-141 (Linear, SiLU, Linear)."""
    p = "time_embedder.timestep_embedder."
    h = timestep_proj(timestep).to(timestep.dtype)
    h = F.linear(h, sd[p + "linear_1.weight"], sd[p + "linear_1.bias"])
    h = F.silu(h)
    return F.linear(h, sd[p + "linear_2.weight"], sd[p + "linear_2.bias"])


def precond(sd: State, x_hat: Tensor, t_hat: Tensor, a: Tensor, sigma_data: float):
    """layers/transformers.py:218-226.  The time-embedding input is t_hat * c_noise (sic)."""
    c_in = 1 / (torch.sqrt(t_hat[:, None, None] ** 2 + sigma_data ** 2))
    c_noise = torch.log(t_hat / sigma_data) / 4.0
    ba = F.linear(x_hat * c_in, sd["linear_x.weight"], sd["linear_x.bias"]) + a[None]
    t = time_embedding(sd, t_hat * c_noise)
    return ba, t


def ada_layer_norm_zero(sd: State, p: str, x: Tensor, t: Tensor, eps: float):
    """primitives/adaptive_layer_norm_zero.py:11-21: shift, scale, gate = chunk3(Linear(SiLU(t)))."""
    shift, scale, gate = F.linear(F.silu(t[..., None, :]), sd[p + "linear.weight"],
                                  sd[p + "linear.bias"]).chunk(3, dim=-1)
    x = F.layer_norm(x, (x.shape[-1],), None, None, eps) * (1 + scale) + shift
    return x, gate


def rms_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """primitives/rms_norm.py:14-19."""
    return (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)) * w


def gen_attn_mask(mask: Tensor, neg_inf: float) -> Tensor:
    """utils/tensor_utils.py:642-646."""
    attn_mask = torch.zeros_like(mask)
    attn_mask[mask == 0] = neg_inf
    return attn_mask


def pair_bias(sd: State, p: str, z: Tensor, z_mask: Tensor, inf: float) -> Tensor:
    """primitives/attentions.py:246,254-255: Linear_nobias(LayerNorm_affine(z)) -> [1,H,S,S] + mask.
    nn.LayerNorm default eps 1e-5 (attentions.py:232 passes no eps)."""
    c_z = z.shape[-1]
    z_norm = F.layer_norm(z, (c_z,), sd[p + "norm_z.weight"], sd[p + "norm_z.bias"], 1e-5)
    bias = F.linear(z_norm, sd[p + "linear_z.weight"]).permute([2, 0, 1])[None]
    return bias + gen_attn_mask(z_mask.type_as(bias), -inf)[None, None]


def dit_attention(sd: State, p: str, bs: Tensor, z: Tensor, t: Tensor, z_mask: Tensor,
                  inf: float, eps: float, bias: Optional[Tensor] = None) -> Tensor:
    """primitives/attentions.py:240-265 (beta is always None, transformers.py:225)."""
    B, S, c = bs.shape
    D = 32
    H = c // D
    bs_norm, gate = ada_layer_norm_zero(sd, p + "norm_s.", bs, t, eps)
    q = F.linear(bs_norm, sd[p + "linear_q.weight"]).reshape([B, S, H, D]).transpose(-2, -3)
    k = F.linear(bs_norm, sd[p + "linear_k.weight"]).reshape([B, S, H, D]).transpose(-2, -3)
    v = F.linear(bs_norm, sd[p + "linear_v.weight"]).reshape([B, S, H, D]).transpose(-2, -3)
    q = rms_norm(q, sd[p + "norm_q.weight"], eps)
    k = rms_norm(k, sd[p + "norm_k.weight"], eps)
    if bias is None:
        bias = pair_bias(sd, p, z, z_mask, inf)
    o = F.scaled_dot_product_attention(q, k, v, bias.to(q.dtype), dropout_p=0, scale=None).transpose(-2, -3)
    o = o.reshape([B, S, -1])
    return F.linear(o, sd[p + "linear_o.weight"], sd[p + "linear_o.bias"]) * gate


def dit_transition(sd: State, p: str, x: Tensor, t: Tensor, eps: float) -> Tensor:
    """primitives/transitions.py:21-30 + feed_forward.py:30-31 (SwiGLU, no biases)."""
    x_norm, gate = ada_layer_norm_zero(sd, p + "ffn_norm.", x, t, eps)
    f = p + "feed_forward."
    h = F.silu(F.linear(x_norm, sd[f + "w1.weight"])) * F.linear(x_norm, sd[f + "w3.weight"])
    return F.linear(h, sd[f + "w2.weight"]) * gate


def dit_stack(sd: State, name: str, n_blocks: int, bs: Tensor, z: Tensor, t: Tensor, z_mask: Tensor,
              inf: float, eps: float, biases: Optional[List[Tensor]] = None) -> Tensor:
    """layers/transformers.py:149-175: bs += Attn; bs += Transition, n_blocks times."""
    for i in range(n_blocks):
        p = f"{name}.blocks.{i}."
        bs = bs + dit_attention(sd, p + "attention.", bs, z, t, z_mask, inf, eps,
                                None if biases is None else biases[i])
        bs = bs + dit_transition(sd, p + "transition.", bs, t, eps)
    return bs


def downscale(sd: State, ba: Tensor, s: Tensor, token_id_to_chunk_sizes: Tensor) -> Tensor:
    """layers/transformers.py:205-212: cumsum over atoms, gather chunk ends, diff => per-token sum,
    divided by (chunk_size + 1e-3), plus s."""
    ba_cumsum = torch.cumsum(F.silu(F.linear(ba, sd["linear_downscale.weight"],
                                             sd["linear_downscale.bias"])), dim=-2)
    inds = torch.cumsum(token_id_to_chunk_sizes, dim=-1) - 1
    value = ba_cumsum[:, inds, :]
    x = torch.cat([value[:, 0:1, :], torch.diff(value, dim=-2)], dim=-2)
    x = x / (token_id_to_chunk_sizes[None, :, None] + 1e-3)
    return x + s[None]


def upscale(sd: State, ba: Tensor, bs: Tensor, atom_id_to_token_id: Tensor) -> Tensor:
    """layers/transformers.py:214-216."""
    return ba + F.linear(bs, sd["linear_upscale.weight"], sd["linear_upscale.bias"])[:, atom_id_to_token_id]


def denoise(sd: State, x_hat: Tensor, t_hat: Tensor, ba: Tensor, sigma_data: float, eps: float) -> Tensor:
    """layers/transformers.py:228-233."""
    c_skip = (sigma_data ** 2 / (sigma_data ** 2 + t_hat ** 2))[:, None, None]
    c_out = (sigma_data * t_hat / torch.sqrt(sigma_data ** 2 + t_hat ** 2))[:, None, None]
    r = F.linear(F.layer_norm(ba, (ba.shape[-1],), sd["norm_r.weight"], sd["norm_r.bias"], eps),
                 sd["linear_r.weight"])
    return c_skip * x_hat + c_out * r


def count_blocks(sd: State, name: str) -> int:
    n = 0
    while f"{name}.blocks.{n}.attention.linear_q.weight" in sd:
        n += 1
    return n


def af3dit_forward(sd: State, batch: Dict[str, Tensor], x_hat: Tensor, t_hat: Tensor, a: Tensor,
                   ap: Tensor, s: Tensor, z: Tensor, sigma_data: float = 16.0, inf: float = 1e9,
                   eps: float = 1e-8, trace: Optional[dict] = None) -> Tensor:
    """layers/transformers.py:235-262."""
    ap_mask, z_mask = batch["ap_mask"], batch["z_mask"]
    n_atom = count_blocks(sd, "atom_dit_encoder")
    n_tok = count_blocks(sd, "token_dit")
    ba, t = precond(sd, x_hat, t_hat, a, sigma_data)
    if trace is not None:
        trace["t"], trace["ba0"] = t, ba
    ba = dit_stack(sd, "atom_dit_encoder", n_atom, ba, ap, t, ap_mask, inf, eps)
    if trace is not None:
        trace["ba_enc"] = ba
    bs = downscale(sd, ba, s, batch["token_id_to_chunk_sizes"])
    if trace is not None:
        trace["bs0"] = bs
    bs = dit_stack(sd, "token_dit", n_tok, bs, z, t, z_mask, inf, eps)
    if trace is not None:
        trace["bs_out"] = bs
    ba = upscale(sd, ba, bs, batch["atom_id_to_token_id"])
    ba = dit_stack(sd, "atom_dit_decoder", count_blocks(sd, "atom_dit_decoder"), ba, ap, t, ap_mask, inf, eps)
    if trace is not None:
        trace["ba_dec"] = ba
    return denoise(sd, x_hat, t_hat, ba, sigma_data, eps)


# ----------------------------------------------------------------------------------------------
# a16 / a17: physics guidance (template selection, Kabsch projection)
# ----------------------------------------------------------------------------------------------
def template_epsilon(ligand_pos: Tensor, ref_mol_poses_dist: Tensor) -> Tensor:
    """models/model.py:231-239: smooth-lDDT-like mismatch [B, C] between the denoised ligand's distance
    matrix and each template's."""
    ligand_dist = torch.norm(ligand_pos[:, :, None] - ligand_pos[:, None], dim=-1)
    delta = (ligand_dist[:, None] - ref_mol_poses_dist[None]).abs()
    epsilon = 0.25 * (torch.sigmoid(-0.5 + delta) + torch.sigmoid(-1 + delta) + torch.sigmoid(-2 + delta)
                      + torch.sigmoid(-4 + delta))
    return epsilon.mean(dim=[-1, -2])


def template_select(ligand_pos: Tensor, ref_mol_poses_dist: Tensor) -> Tensor:
    """models/model.py:240: argmin over templates."""
    return torch.argmin(template_epsilon(ligand_pos, ref_mol_poses_dist), dim=-1)


def weighted_rigid_align(x_pred: Tensor, x_gt: Tensor, weights: Tensor) -> Tensor:
    """utils/tensor_utils.py:724-778.  Returns x_gt rotated/translated onto x_pred's frame."""
    x_pred, x_gt, weights = x_pred.float(), x_gt.float(), weights.float()
    if len(x_gt.shape) == 2:
        x_gt = x_gt[..., None, :, :]
    wsum = torch.sum(weights[..., None, :], dim=-1, keepdim=True)
    mu_pred = torch.sum(x_pred * weights[..., None, :, None], dim=-2) / wsum
    mu_gt = torch.sum(x_gt * weights[..., None, :, None], dim=-2) / wsum
    x_pred_hat = x_pred - mu_pred[..., None, :]
    x_gt_hat = x_gt - mu_gt[..., None, :]
    outer = torch.einsum("...ij,...ik->...ijk", x_gt_hat, x_pred_hat)
    H = torch.sum(outer * weights[..., None, :, None, None], dim=-3)
    U, _, Vh = torch.linalg.svd(H)
    Fm = torch.eye(3, device=H.device)
    Fm[-1, -1] = -1
    R = torch.matmul(U, Vh)
    R_reflection = torch.matmul(U, Fm).matmul(Vh)
    R = torch.where((torch.det(R) < 0)[..., None, None], R_reflection, R)
    R = torch.transpose(R, -1, -2)
    return torch.einsum("...ij,...kj->...ki", R, x_gt_hat) + mu_pred[..., None, :]


def physics_direction(x_hat: Tensor, x_denoised: Tensor, aligned: Tensor, t_hat: Tensor,
                      weights: Tensor) -> Tensor:
    """models/model.py:247-250 (and :258-261): blend of the denoiser direction and the ligand projection."""
    d_ligand = (x_hat - aligned) / t_hat[..., None, None] * weights[None, :, None]
    return ((x_hat - x_denoised) / t_hat[..., None, None]) * (1 - weights[None, :, None]) + d_ligand


def euler_update(x_hat: Tensor, d_cur: Tensor, t_hat: Tensor, t_next, eta: float) -> Tensor:
    """models/model.py:264,278-281."""
    dt = (t_next - t_hat)[..., None, None]
    return x_hat + eta * dt * d_cur


# ----------------------------------------------------------------------------------------------
# a1 / a2: the sampler
# ----------------------------------------------------------------------------------------------
def sample_diffusion(sd: State, batch: Dict[str, Tensor], a: Tensor, ap: Tensor, s: Tensor, z: Tensor,
                     num_sample: int = 5, steps: int = 200, gamma_0: float = 0.8, gamma_min: float = 1.0,
                     noise_scale_lambda: float = 1.003, step_scale_eta: float = 1.5,
                     ode_step_scale_eta: float = 1.0, ref_mol_poses: Optional[Tensor] = None,
                     mmff_gamma_0_factor: float = 1.0, align_ref_pos: bool = True,
                     karras_noise_schedule_power: float = 7, sigma_data: float = 16.0,
                     rng=None, mmff_fn: Optional[Callable] = None, trace: Optional[list] = None,
                     denoiser: Optional[Callable] = None, max_steps: Optional[int] = None) -> Tensor:
    """models/model.py:157-282 with the trunk outputs (a, ap, s, z) passed in (prepare_solver :144 is
    out of scope) and the RNG injected.  `mmff_fn(x_lig [B,n,3]) -> [B,n,3]` stands for
    get_next_step_pos(ref_mol, ., mmff_iters) (:26-52); None == `ref_mol=None`.
    `trace` (a list) receives one dict per step.  `denoiser(x_hat, t_hat)` overrides the oracle denoiser
    (used to teacher-force another implementation through the same control flow)."""
    with torch.no_grad():
        x_exists = batch["a_mask"]
        device, dtype = batch["x_gt"].device, batch["x_gt"].dtype
        rng = rng or TorchRNG(device, dtype)
        is_ligand_atom = batch["is_ligand"][batch["atom_id_to_token_id"]].bool()
        batch_ref_pos = batch["ref_pos"][None].repeat([num_sample, 1, 1])
        ref_mol_poses_dist = None
        if ref_mol_poses is not None:
            ref_mol_poses = ref_mol_poses.to(device)
            ref_mol_poses_dist = torch.norm(ref_mol_poses[:, :, None] - ref_mol_poses[:, None], dim=-1)
        if denoiser is None:
            def denoiser(x_hat, t_hat):
                return af3dit_forward(sd, batch, x_hat, t_hat, a, ap, s, z, sigma_data)
        # prepare_solver, model.py:147-148
        sigmas = karras_noise_schedule(num_steps=steps, p=karras_noise_schedule_power).to(device).to(dtype)
        x_next = sigmas[0] * rng.normal((num_sample, *batch["x_gt"].shape[-2:]))
        for i, (t_cur, t_next) in enumerate(zip(sigmas[:-1], sigmas[1:])):
            if max_steps is not None and i >= max_steps:
                break
            u = torch.stack([rng.rand((num_sample,)) for _ in range(4)], dim=-1)
            trans = rng.normal((num_sample, 3))
            x_cur = centre_random_augmentation(x_next, x_exists, u, trans)
            if t_cur > gamma_min:
                t_hat = torch.full([num_sample], fill_value=t_cur * (gamma_0 + 1), device=device, dtype=dtype)
                noise = rng.normal(x_cur.shape)
                x_hat = diffuse(x_cur, t_hat, t_cur, noise, noise_scale_lambda)
            else:
                t_hat = torch.full([num_sample], fill_value=t_cur, device=device, dtype=dtype)
                x_hat = x_cur
            x_denoised = denoiser(x_hat, t_hat)
            used_inds = None
            if align_ref_pos and t_cur > gamma_min * mmff_gamma_0_factor:
                weights = x_exists * batch["is_ligand"][batch["atom_id_to_token_id"]]
                if ref_mol_poses is not None:
                    used_inds = template_select(x_denoised[:, is_ligand_atom], ref_mol_poses_dist)
                    batch_ref_pos[:, is_ligand_atom] = ref_mol_poses[used_inds]
                aligned = weighted_rigid_align(x_denoised * x_exists[..., None], batch_ref_pos, weights)
                d_cur = physics_direction(x_hat, x_denoised, aligned, t_hat, weights)
            elif mmff_fn is not None and t_cur <= gamma_min * mmff_gamma_0_factor:
                weights = x_exists * batch["is_ligand"][batch["atom_id_to_token_id"]]
                x_ref = x_denoised.clone()
                x_ref[:, is_ligand_atom] = mmff_fn(x_denoised[:, is_ligand_atom])
                aligned = weighted_rigid_align(x_denoised * x_exists[..., None], x_ref, weights)
                d_cur = physics_direction(x_hat, x_denoised, aligned, t_hat, weights)
            else:
                d_cur = (x_hat - x_denoised) / t_hat[..., None, None]
            eta = step_scale_eta if t_cur > gamma_min else ode_step_scale_eta
            x_next = euler_update(x_hat, d_cur, t_hat, t_next, eta)
            if trace is not None:
                trace.append(dict(i=i, t_cur=float(t_cur), t_next=float(t_next), t_hat=t_hat.clone(),
                                  x_cur=x_cur, x_hat=x_hat, x_denoised=x_denoised, x_next=x_next,
                                  used_inds=used_inds))
    return x_next


# ----------------------------------------------------------------------------------------------
# pair-energy physics backend (extension; NOT reference arithmetic)
# ----------------------------------------------------------------------------------------------
def pair_energy(x: Tensor, x_exists: Tensor, sigma: Tensor, eps: Tensor, partner: Optional[Tensor],
                partner_r0: Optional[Tensor], partner_k: Optional[Tensor], rows: Optional[Tensor],
                clash_k: float = 10.0, clash_scale: float = 0.6, cutoff: float = 10.0,
                softcore: float = 0.1) -> Tensor:
    """Restatement of the functional form DEFINED in physdock_b200/csrc/physics.cu (header comment) as dense
    differentiable tensor algebra; its gradient comes from autograd.  This is the checker of the opt-in physics backend
    that replaces get_next_step_pos (models/model.py:26-52, RDKit MMFF94, parity unpinned): the reference has no
    arithmetic for it (its lj/bond losses are stubs, models/loss_module.py:284-308).

    x [B,Na,3]; rows [n] (None = all atoms); partner tables [Na,E] (-1 = empty).  Returns energy [B].
    """
    B, Na, _ = x.shape
    dt = x.dtype
    if rows is None:
        rows = torch.arange(Na)
    rows = rows.long()
    in_rows = torch.zeros(Na, dtype=torch.bool)
    in_rows[rows] = True
    xi = x[:, rows]                                                     # [B,n,3]
    diff = xi[:, :, None, :] - x[:, None, :, :]                         # [B,n,Na,3]
    r2 = (diff ** 2).sum(-1)
    d2 = r2 + 1e-12
    d = torch.sqrt(d2)
    ex = x_exists.to(dt)
    sig = 0.5 * (sigma[rows][:, None] + sigma[None, :]).to(dt)          # [n,Na]
    se = torch.sqrt(eps.clamp(min=0).to(dt))
    e_ij = se[rows][:, None] * se[None, :]
    nb = (ex[rows][:, None] * ex[None, :]).bool()
    excl = torch.zeros(len(rows), Na, dtype=torch.bool)
    excl[torch.arange(len(rows)), rows] = True
    if partner is not None and partner.numel() > 0:
        pr = partner[rows].long()                                       # [n,E]
        valid = pr >= 0
        ri = torch.arange(len(rows))[:, None].expand_as(pr)
        excl[ri[valid], pr[valid]] = True
    nb = nb & ~excl
    nbf = (nb[None] & (r2.detach() < cutoff * cutoff)).to(dt)           # the cutoff is a hard mask (no switching)
    sig2 = sig ** 2
    u = sig2 / (d2 + softcore * sig2)
    s6 = u ** 3
    e_lj = e_ij * (s6 * s6 - 2.0 * s6)
    pen = torch.clamp(clash_scale * sig - d, min=0.0)
    e_nb = (e_lj + clash_k * pen * pen) * nbf
    w = torch.where(in_rows, 0.5, 1.0).to(dt)[None, None, :]
    energy = (w * e_nb).sum(dim=(1, 2))
    if partner is not None and partner.numel() > 0:
        pr = partner[rows].long()
        valid = (pr >= 0)
        prc = pr.clamp(min=0)
        xj = x[:, prc]                                                  # [B,n,E,3]
        db = torch.sqrt(((xi[:, :, None, :] - xj) ** 2).sum(-1) + 1e-12)
        k = partner_k[rows].to(dt) * valid.to(dt) * ex[rows][:, None] * ex[prc]
        wb = torch.where(in_rows[prc], 0.5, 1.0).to(dt)
        energy = energy + (wb * k * (db - partner_r0[rows].to(dt)) ** 2).sum(dim=(1, 2))
    return energy


def pair_energy_grad(x: Tensor, *args, **kw):
    """(energy [B], dE/dx [B,Na,3]) by autograd; the backend only reports the rows' gradient."""
    xg = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        e = pair_energy(xg, *args, **kw)
        g, = torch.autograd.grad(e.sum(), xg)
    return e.detach(), g


def pair_energy_descend(x: Tensor, x_exists, sigma, eps, partner, partner_r0, partner_k, rows, iters: int = 5,
                        step: float = 0.01, gmax: float = 50.0, **kw) -> Tensor:
    """`iters` clamped gradient-descent steps on the row atoms (the backend's stand-in for MMFFOptimizeMolecule(maxIters))."""
    Na = x.shape[1]
    mask = torch.zeros(Na, dtype=x.dtype)
    mask[(torch.arange(Na) if rows is None else rows.long())] = 1
    cur = x.clone()
    for _ in range(iters):
        _, g = pair_energy_grad(cur, x_exists, sigma, eps, partner, partner_r0, partner_k, rows, **kw)
        cur = cur - step * g.clamp(-gmax, gmax) * mask[None, :, None]
    return cur


# ----------------------------------------------------------------------------------------------
# caller-side ranking (redocking.py), numpy as in the reference
# ----------------------------------------------------------------------------------------------
def pairwise_pose_rmsd(pred) -> "np.ndarray":
    """redocking.py:391.  pred [S,n,3] -> [S,S] float64."""
    import numpy as np
    pred = np.asarray(pred, dtype=np.float64)
    return np.sqrt(np.mean(np.linalg.norm(pred[:, None] - pred[None], axis=-1) ** 2, axis=-1))


def get_representatives(distance_matrix, num_clusters: int = 5):
    """redocking.py:393-410 (sklearn KMeans, random_state=0; medoid per cluster)."""
    import numpy as np
    from sklearn.cluster import KMeans
    num_elements = len(distance_matrix)
    coordinates = np.zeros((num_elements, num_elements))
    for i in range(num_elements):
        coordinates[i] = distance_matrix[i]
    kmeans = KMeans(n_clusters=num_clusters, random_state=0)
    kmeans.fit(coordinates)
    labels = kmeans.labels_
    representatives_indices = []
    for cluster_id in range(num_clusters):
        cluster_indices = np.where(labels == cluster_id)[0]
        avg_distances = np.mean(distance_matrix[cluster_indices, :], axis=0)
        representatives_indices.append(cluster_indices[np.argmin(avg_distances[cluster_indices])])
    return [int(i) for i in representatives_indices]


def rank_poses(pred, num_clusters: int = 5):
    """redocking.py:412-423."""
    dist = pairwise_pose_rmsd(pred)
    return rank_from_distance_matrix(dist, num_clusters), dist


def rank_from_distance_matrix(dist, num_clusters: int = 5):
    """redocking.py:412-423 given the pose-RMSD matrix of redocking.py:391."""
    if len(dist) > num_clusters:
        ids = get_representatives(dist, num_clusters)
        ids_1 = get_representatives(dist, 1)[0]
        if ids_1 in ids:
            ids.remove(ids_1)
            ids = [ids_1] + ids
        else:
            ids = [ids_1] + ids[:4]
    else:
        ids = list(range(len(dist)))
    return ids


def rank_conformer_templates(x_pred_lig: Tensor, ref_mol_poses: Tensor, n_keep: int) -> Tensor:
    """redocking.py:326-332."""
    ref_mol_poses_dist = torch.norm(ref_mol_poses[:, :, None] - ref_mol_poses[:, None], dim=-1)
    ligand_dist = torch.norm(x_pred_lig[:, :, None] - x_pred_lig[:, None], dim=-1)
    delta = (ligand_dist[:, None] - ref_mol_poses_dist[None]).abs()
    epsilon = 0.25 * (torch.sigmoid(-0.5 + delta) + torch.sigmoid(-1 + delta) + torch.sigmoid(
        -2 + delta) + torch.sigmoid(-4 + delta))
    epsilon = epsilon.mean(dim=[-1, -2, -4])
    return torch.argsort(epsilon)[:n_keep]


def rounds_bookkeeping(x_preds, is_ligand_atom: Tensor, ref_mol_poses: Optional[Tensor], accept_fn,
                       num_augmentation_sample: int, max_samples: int, physics_correction: bool,
                       mmff_gamma_0_factor_start: float):
    """The per-system round loop of redocking.py:163-338 with the model call replaced by the recorded predictions
    `x_preds` (one [B,Na,3] CPU tensor per executed round): accept/reject lists (:302-317), adaptive boundary (:318-322),
    early stop (:323-324), conformer ranking for the next round's reference templates (:326-335), reject fill-in (:337-338).
    Returns (per-round dicts, final accepted stack [<= max_samples, Na, 3], n_accepted, final factor)."""
    from collections import deque
    accept_samples, reject_samples = [], deque([], maxlen=max_samples)
    ligand_templates, reference_templates = [], []
    mmff_gamma_0_factor = mmff_gamma_0_factor_start
    out = []
    for recycle_id, x_pred in enumerate(x_preds):
        assert recycle_id == 0 or physics_correction
        n_templates = len(ligand_templates) + len(reference_templates) if recycle_id > 0 else 0
        factor_used = mmff_gamma_0_factor
        pass_flags = []
        for x in x_pred:
            pass_flag = bool(accept_fn(x)) if (physics_correction and accept_fn is not None) else True
            pass_flags.append(pass_flag)
            if pass_flag:
                ligand_templates.append(x[is_ligand_atom])
                accept_samples.append(x)
            else:
                reject_samples.append(x)
        used_inds = None
        stop = False
        if physics_correction:
            if any(pass_flags):
                mmff_gamma_0_factor = mmff_gamma_0_factor * 1.15
            else:
                mmff_gamma_0_factor = max(mmff_gamma_0_factor * 0.7, 1)
            if len(accept_samples) >= max_samples:
                stop = True
            else:
                used_inds = rank_conformer_templates(x_pred[:, is_ligand_atom], ref_mol_poses,
                                                     max_samples - len(ligand_templates))
                reference_templates = [ref_mol_poses[i] for i in used_inds]
        out.append(dict(recycle_id=recycle_id, factor=factor_used, pass_flags=pass_flags, n_templates=n_templates,
                        used_inds=used_inds, stop=stop))
        if stop:
            break
    n_accepted = len(accept_samples)
    if len(accept_samples) < num_augmentation_sample:
        accept_samples = accept_samples + [_ for _ in reject_samples]
    return out, torch.stack(accept_samples[:max_samples], dim=0), n_accepted, mmff_gamma_0_factor


def rmsd(a: Tensor, b: Tensor) -> Tensor:
    """Per-sample RMSD in Angstrom between two coordinate sets [B,N,3] (the parity metric)."""
    return ((a.double() - b.double()) ** 2).sum(-1).mean(-1).sqrt()
