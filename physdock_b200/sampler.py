"""The reverse-diffusion sampler on B200 -- mirror of `PhysDock.sample_diffusion`
(reference PhysDock/models/model.py:157-282) and of the module surface `PhysDock` exposes to
redocking.py:284-299 / screening.py:294.

Per step (model.py:211-281) the device work is, with no host<->device sync:
    pdk_centre_augment    (centre_random_augmentation + diffuse, fused)
    pdk_dit_denoise_cond  (AF3DiT; its last kernel also writes the Euler update when the step has no physics guidance)
    [pdk_template_select + pdk_rigid_align + pdk_euler_update]      (physics guidance, RDKit-free part)
Everything the denoiser derives from the noise level alone (time embedding, the 36 AdaLN-Zero modulations, c_in / c_skip /
c_out) is computed for the WHOLE schedule when the sampler is built (`B200DiT.conditioning_table`): all samples of a step
share one noise level, known before the loop starts.
The noise schedule lives on the host exactly as in the reference (`karras_noise_schedule` runs on the CPU,
model.py:147), so the `t_cur > gamma_min` branches are taken on host floats instead of syncing on a device
scalar each step (model.py:213).  Random numbers are drawn with the same torch calls, shapes, dtypes and
order as the reference, on the coordinates' device, so a run is reproducible against the reference sampler
on the same device type under the same `torch.manual_seed`.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from .dit import B200DiT


def karras_noise_schedule(num_steps: int = 200, sigma_data: float = 16, s_max: float = 160,
                          s_min: float = 4 * 10e-4, p: float = 7) -> torch.Tensor:
    """model.py:117-129, verbatim arithmetic (CPU fp32)."""
    step_indices = torch.arange(num_steps, dtype=torch.float32)
    t_steps = sigma_data * (s_max ** (1 / p) + step_indices / (num_steps - 1) * (
            s_min ** (1 / p) - s_max ** (1 / p))) ** p
    return torch.cat([t_steps, torch.zeros_like(t_steps[:1])])


class DeviceRNG:
    """The reference's random draws (model.py:77,148; tensor_utils.py:555-557,582) on `device`."""

    def __init__(self, device, dtype=torch.float32):
        self.device, self.dtype = device, dtype

    def rand(self, shape):
        return torch.rand(list(shape), device=self.device, dtype=torch.float32)

    def normal(self, shape):
        return torch.normal(0, 1, size=tuple(shape), dtype=self.dtype, device=self.device, requires_grad=False)


# ---------------------------------------------------------------------------------- op wrappers
def centre_augment_noise(x: torch.Tensor, x_exists: torch.Tensor, u4: torch.Tensor, trans: torch.Tensor,
                         noise: Optional[torch.Tensor] = None, noise_scale_lambda: float = 1.0,
                         noise_scale: float = 0.0, s: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """centre_random_augmentation (tensor_utils.py:576-586) [+ diffuse (model.py:70-85) when noise is given]."""
    lib = _lib.load()
    B, Na, _ = x.shape
    x, u4, trans = x.float().contiguous(), u4.float().contiguous(), trans.float().contiguous()
    x_exists = x_exists.float().contiguous()
    if noise is not None:
        noise = noise.float().contiguous()
    out = torch.empty_like(x) if out is None else out
    _lib.check(lib.pdk_centre_augment(_lib.ptr(x), _lib.ptr(x_exists), _lib.ptr(u4), _lib.ptr(trans), _lib.ptr(noise),
                                      float(noise_scale_lambda), float(noise_scale), float(s), _lib.ptr(out), B, Na,
                                      _lib.stream_ptr(x.device)), "pdk_centre_augment")
    return out


def euler_update(x_hat, x_den, t_hat, t_next: float, eta: float, aligned=None, weights=None, out=None):
    """model.py:247-250,263-264,278-281."""
    lib = _lib.load()
    B, Na, _ = x_hat.shape
    out = torch.empty_like(x_hat) if out is None else out
    if aligned is not None:
        aligned, weights = aligned.float().contiguous(), weights.float().contiguous()
    _lib.check(lib.pdk_euler_update(_lib.ptr(x_hat.contiguous()), _lib.ptr(x_den.contiguous()), _lib.ptr(aligned),
                                    _lib.ptr(weights), _lib.ptr(t_hat.float().contiguous()), float(t_next), float(eta),
                                    _lib.ptr(out), B, Na, _lib.stream_ptr(x_hat.device)), "pdk_euler_update")
    return out


def template_select(x_den, lig_idx, ref_dist, ref_poses, batch_ref_pos):
    """model.py:231-241.  Returns (eps [B,C], used [B]); batch_ref_pos is updated in place."""
    lib = _lib.load()
    B, Na, _ = x_den.shape
    Cn, n = ref_poses.shape[0], ref_poses.shape[1]
    eps = torch.empty(B, Cn, dtype=torch.float32, device=x_den.device)
    used = torch.empty(B, dtype=torch.int64, device=x_den.device)
    _lib.check(lib.pdk_template_select(_lib.ptr(x_den.contiguous()), _lib.ptr(lig_idx), _lib.ptr(ref_dist),
                                       _lib.ptr(ref_poses), _lib.ptr(eps), _lib.ptr(used), _lib.ptr(batch_ref_pos),
                                       B, Na, n, Cn, _lib.stream_ptr(x_den.device)), "pdk_template_select")
    return eps, used


def weighted_rigid_align(x_den, x_exists, x_gt, weights, out=None):
    """weighted_rigid_align(x_den * x_exists[..., None], x_gt, weights) (tensor_utils.py:724-778)."""
    lib = _lib.load()
    B, Na, _ = x_den.shape
    out = torch.empty_like(x_den) if out is None else out
    x_gt = x_gt.float().contiguous()
    _lib.check(lib.pdk_rigid_align(_lib.ptr(x_den.contiguous()), _lib.ptr(x_exists.float().contiguous()), _lib.ptr(x_gt),
                                   1 if x_gt.dim() == 3 else 0, _lib.ptr(weights.float().contiguous()), _lib.ptr(out),
                                   B, Na, _lib.stream_ptr(x_den.device)), "pdk_rigid_align")
    return out


def _rdkit_conformers(ref_mol, num_confs: int = 512):
    """The conformer pool of model.py:188-203 (RDKit EmbedMultipleConfs, enforceChirality=True).  Third-party, un-vendored:
    only reachable where rdkit is importable."""
    try:
        import copy
        from rdkit.Chem import AllChem
    except Exception as e:  # pragma: no cover - rdkit is absent from this image
        raise _lib.PdkError("use_ref_mol_poses=True without ref_mol_poses needs RDKit EmbedMultipleConfs (model.py:185-203): "
                            "pass ref_mol_poses, or conformer_fn=... returning [C, n_lig, 3]") from e
    mol = copy.deepcopy(ref_mol)                                                   # pragma: no cover
    cids = AllChem.EmbedMultipleConfs(mol, numConfs=num_confs, enforceChirality=True)   # pragma: no cover
    n = mol.GetNumAtoms()                                                          # pragma: no cover
    out = torch.zeros(num_confs, n, 3)                                             # pragma: no cover
    for i, cid in enumerate(cids):                                                 # pragma: no cover
        conf = mol.GetConformer(cid)
        for j in range(n):
            pos = conf.GetAtomPosition(j)
            out[i, j, 0], out[i, j, 1], out[i, j, 2] = pos.x, pos.y, pos.z
    return out                                                                     # pragma: no cover


def _rdkit_mmff(ref_mol, mmff_iters):
    """get_next_step_pos (model.py:26-52): RDKit MMFF94 on the ligand alone, on the CPU, per sample.
    Third-party, un-vendored (rdkit==2024.3.3, enviroment.yaml:33): parity unpinned, see DESIGN.md."""
    try:
        from rdkit.Chem import AllChem
        from rdkit.Geometry import Point3D
    except Exception as e:  # pragma: no cover - rdkit is absent from this image
        raise _lib.PdkError("ref_mol was given but rdkit is not importable: the MMFF94 step of "
                            "PhysDock/models/model.py:252-261 needs rdkit==2024.3.3; pass ref_mol=None "
                            "to sample without it") from e

    def fn(lig_pos: torch.Tensor) -> torch.Tensor:   # pragma: no cover
        conf = ref_mol.GetConformer()
        out = []
        for pos in lig_pos.cpu().tolist():
            for i in range(conf.GetNumAtoms()):
                conf.SetAtomPosition(i, Point3D(*pos[i]))
            AllChem.MMFFOptimizeMolecule(ref_mol, mmffVariant="MMFF94", maxIters=mmff_iters,
                                         ignoreInterfragInteractions=True)
            conf = ref_mol.GetConformer()
            out.append([[conf.GetAtomPosition(i).x, conf.GetAtomPosition(i).y, conf.GetAtomPosition(i).z]
                        for i in range(ref_mol.GetNumAtoms())])
        return torch.tensor(out, device=lig_pos.device, dtype=lig_pos.dtype)

    return fn


# ---------------------------------------------------------------------------------- the sampler
class DiffusionSampler:
    """State + one-step driver of the reverse-diffusion loop (model.py:211-281) for one prepared complex.

    `sample_diffusion` below is `begin(); for i: step(i)`.  bench.py drives `step` (device-resident inputs)
    and `step_from_host` (pinned host buffers in, coordinates out) directly.
    """

    def __init__(self, dit: B200DiT, batch: Dict[str, torch.Tensor], a, ap, s, z, num_sample: int = 5,
                 steps: int = 200, gamma_0: float = 0.8, gamma_min: float = 1.0, noise_scale_lambda: float = 1.003,
                 step_scale_eta: float = 1.5, ode_step_scale_eta: float = 1.0, ref_mol=None,
                 ref_mol_poses: Optional[torch.Tensor] = None, use_ref_mol_poses: bool = False,
                 mmff_gamma_0_factor: float = 1.0, mmff_iters: int = 5, align_ref_pos: bool = True,
                 karras_noise_schedule_power: float = 7, rng=None, mmff_fn: Optional[Callable] = None,
                 use_cuda_graph: Optional[bool] = None, physics_field=None, physics_step: float = 0.002,
                 physics_gmax: float = 50.0, conformer_fn: Optional[Callable] = None):
        dev = batch["x_gt"].device
        # CUDA-graph replay of the denoiser saves ~3.5 % per step (0.065 ms at B=8..16) but capturing the 121 launches costs
        # ~15 ms: None = auto = launch eagerly (one C call enqueues all kernels, with PDL) during the first trajectory of this
        # sampler and capture only when it is reused for a second one.  A one-shot `sample_diffusion` call (a redocking round,
        # a screening ligand) therefore never pays for a capture; bench.py, which times the steady state, passes True.
        self.use_cuda_graph = use_cuda_graph
        self._steps_run = 0
        if dev.type != "cuda":
            raise _lib.PdkError("sample_diffusion needs the batch on a CUDA device (no CPU fallback)")
        self.dit, self.dev, self.B, self.Na = dit, dev, num_sample, batch["x_gt"].shape[-2]
        self.gamma_0, self.gamma_min, self.lam = gamma_0, gamma_min, noise_scale_lambda
        self.eta_s, self.eta_d = step_scale_eta, ode_step_scale_eta
        self.align_ref_pos, self.mmff_factor = align_ref_pos, mmff_gamma_0_factor
        self.rng = rng or DeviceRNG(dev, torch.float32)
        self.x_exists = batch["a_mask"].to(dev).float().contiguous()
        lig_atom_f = batch["is_ligand"].to(dev)[batch["atom_id_to_token_id"].to(dev)].float()
        self.is_ligand_atom = lig_atom_f.bool()
        self.lig_idx = torch.nonzero(self.is_ligand_atom).flatten().int().contiguous()
        self.weights = (self.x_exists * lig_atom_f).contiguous()
        self.batch_ref_pos = batch["ref_pos"].to(dev).float()[None].repeat([self.B, 1, 1]).contiguous()
        self.ref_mol_poses, self.ref_dist = None, None
        if ref_mol_poses is None and use_ref_mol_poses:
            # model.py:185-203: 512 RDKit conformers; `conformer_fn() -> [C, n_lig, 3]` is the host hook for it (like mmff_fn)
            ref_mol_poses = conformer_fn() if conformer_fn is not None else _rdkit_conformers(ref_mol, 512)
            ref_mol_poses = ref_mol_poses[:, :int(self.is_ligand_atom.sum())]
        if ref_mol_poses is not None:
            self.ref_mol_poses = ref_mol_poses.to(dev).float().contiguous()
            if self.ref_mol_poses.shape[1] != self.lig_idx.numel():     # the reference swallows this (try/except)
                raise _lib.PdkError("ref_mol_poses atom count != number of ligand atoms")
            self.ref_dist = torch.norm(self.ref_mol_poses[:, :, None] - self.ref_mol_poses[:, None], dim=-1).contiguous()
        if ref_mol is not None and mmff_fn is None:
            mmff_fn = _rdkit_mmff(ref_mol, mmff_iters)
        self.mmff_fn = mmff_fn
        # opt-in device-resident replacement of the MMFF step (physics.PairEnergyField); takes precedence over mmff_fn
        self.physics_field, self.mmff_iters = physics_field, mmff_iters
        self.physics_step, self.physics_gmax = physics_step, physics_gmax
        dit._pack()
        sig = dit._complex_signature(batch, a, ap, s, z)
        if sig != dit._complex_sig:
            dit.prepare_complex(batch, a, ap, s, z)
            dit._complex_sig = sig
        self.steps = steps
        self.sigmas = karras_noise_schedule(num_steps=steps, p=karras_noise_schedule_power)     # host, fp32
        shape = (self.B, self.Na, 3)
        # persistent buffers: every kernel argument of a step is a pointer into these (CUDA-graph replay)
        self.x_next = torch.empty(shape, dtype=torch.float32, device=dev)
        self.x_hat, self.x_den, self.aligned = (torch.empty_like(self.x_next) for _ in range(3))
        self.t_hat_dev = torch.empty(self.B, dtype=torch.float32, device=dev)
        self.last_used = None
        self._sched_cache: Dict[int, tuple] = {}
        self._copy_stream = None
        self._sched_floats: Dict[int, tuple] = {}
        # conditioning of the whole schedule (2 launches, once): row i = [modulations | c_in c_skip c_out t_hat | t_next eta . .]
        t_hats = torch.stack([self.schedule(i)[2].float() for i in range(steps)]).to(dev)
        self.cond_table = dit.conditioning_table(t_hats)
        n_mod = dit.cond_width() - 8
        extra = torch.zeros(steps, 2, dtype=torch.float32)
        for i in range(steps):
            _, t_next, _, stochastic, _ = self.schedule(i)
            extra[i, 0], extra[i, 1] = t_next, (self.eta_s if stochastic else self.eta_d)
        self.cond_table[:, n_mod + 4:n_mod + 6] = extra.to(dev)
        self.cond_cur = torch.empty(dit.cond_width(), dtype=torch.float32, device=dev)

    def begin(self) -> torch.Tensor:
        """x_0 = sigma_0 * N(0,1)   (prepare_solver, model.py:148)."""
        self.x_next.copy_(self.sigmas[0].to(self.dev) * self.rng.normal((self.B, self.Na, 3)))
        return self.x_next

    def schedule(self, i: int):
        """Host-side scalars of step i: (t_cur, t_next, t_hat, stochastic, noise_scale), model.py:213-220.
        Computed once per step index (a handful of 0-dim CPU tensor ops = tens of microseconds of host time that the
        end-to-end path, which synchronises every step, would otherwise pay on every call)."""
        hit = self._sched_cache.get(i)
        if hit is None:
            hit = self._sched_cache[i] = self._schedule(i)
        return hit

    def _schedule(self, i: int):
        t_cur, t_next = self.sigmas[i], self.sigmas[i + 1]
        stochastic = bool(t_cur > self.gamma_min)
        if stochastic:
            t_hat = t_cur * (self.gamma_0 + 1)                       # 0-dim fp32 multiply, as model.py:215
            noise_scale = float(torch.sqrt(t_hat ** 2 - t_cur ** 2))  # model.py:81
        else:
            t_hat, noise_scale = t_cur, 0.0
        return t_cur, t_next, t_hat, stochastic, noise_scale

    def draw(self, i: int):
        """The step's random tensors in the reference's order (tensor_utils.py:549-557 x2, :582; model.py:77)."""
        stochastic = self.schedule(i)[3]
        u4 = torch.stack([self.rng.rand((self.B,)) for _ in range(4)], dim=-1)
        trans = self.rng.normal((self.B, 3))
        noise = self.rng.normal((self.B, self.Na, 3)) if stochastic else None
        return u4, trans, noise

    def launches_per_step(self, i: int) -> int:
        """Kernels of THIS library enqueued by step(i) (torch's own random draws / copies not counted)."""
        thr = self.gamma_min * self.mmff_factor
        t_cur = self.sigmas[i]
        early, late = bool(t_cur > thr), bool(t_cur <= thr)
        n = 1 + self.dit.launches_per_denoise(cond=True)
        if self.align_ref_pos and early:
            n += (2 if self.ref_mol_poses is not None else 0) + 2              # template eps + pick, rigid align, euler
        elif late and self.physics_field is not None:
            n += self.physics_field.launches_per_descend(self.mmff_iters) + 2
        elif late and self.mmff_fn is not None:
            n += 2
        return n

    def step(self, i: int, randoms=None, teacher_x_hat: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One iteration of the loop at schedule index i; returns x_next (a persistent buffer, updated in place by the
        following step: clone it to keep it)."""
        t_cur, t_next, t_hat, stochastic, noise_scale = self.schedule(i)
        fl = self._sched_floats.get(i)
        if fl is None:       # python floats of the step's scalars + the two branch conditions, computed once per index
            thr = self.gamma_min * self.mmff_factor
            fl = self._sched_floats[i] = (float(t_hat), float(t_next), bool(t_cur > thr), bool(t_cur <= thr))
        t_hat_f, t_next_f, early, late = fl
        u4, trans, noise = randoms if randoms is not None else self.draw(i)
        centre_augment_noise(self.x_next, self.x_exists, u4, trans, noise, self.lam, noise_scale, out=self.x_hat)
        if teacher_x_hat is not None:
            self.x_hat.copy_(teacher_x_hat)
        # does this step blend a physics direction into d_cur (model.py:223-261)?  If not, the denoiser's last kernel also
        # writes the Euler update (x_next is dead once centre_augment has consumed it, so it is updated in place).
        guided = (self.align_ref_pos and early) or (late and (self.physics_field is not None or self.mmff_fn is not None))
        fused_next = None if guided else self.x_next
        graph = self.use_cuda_graph if self.use_cuda_graph is not None else self._steps_run >= self.steps
        self._steps_run += 1
        if graph:
            self.cond_cur.copy_(self.cond_table[i])          # one 166 KB device copy selects the step's conditioning
            self.dit.denoise_cond_graphed(self.x_hat, self.cond_cur, self.x_den, fused_next)
        else:
            self.dit.denoise_cond(self.x_hat, self.cond_table[i], self.x_den, fused_next)
        self.last_used = None
        if not guided:
            return self.x_next
        self.t_hat_dev.fill_(t_hat_f)
        physics = False
        if self.align_ref_pos and early:
            if self.ref_mol_poses is not None:
                _, self.last_used = template_select(self.x_den, self.lig_idx, self.ref_dist, self.ref_mol_poses,
                                                    self.batch_ref_pos)
            weighted_rigid_align(self.x_den, self.x_exists, self.batch_ref_pos, self.weights, out=self.aligned)
            physics = True
        elif self.physics_field is not None and late:
            # model.py:252-261 with get_next_step_pos replaced by mmff_iters descent steps on the pair energy (GPU)
            x_ref = self.physics_field.descend(self.x_den, iters=self.mmff_iters, step=self.physics_step,
                                               gmax=self.physics_gmax)
            weighted_rigid_align(self.x_den, self.x_exists, x_ref, self.weights, out=self.aligned)
            physics = True
        elif self.mmff_fn is not None and late:
            x_ref = self.x_den.clone()
            x_ref[:, self.is_ligand_atom] = self.mmff_fn(self.x_den[:, self.is_ligand_atom])
            weighted_rigid_align(self.x_den, self.x_exists, x_ref, self.weights, out=self.aligned)
            physics = True
        eta = self.eta_s if stochastic else self.eta_d
        euler_update(self.x_hat, self.x_den, self.t_hat_dev, t_next_f, eta, self.aligned if physics else None,
                     self.weights if physics else None, out=self.x_next)
        return self.x_next

    def upload_randoms(self, i: int, u4_host, trans_host, noise_host):
        """H2D of step i's random tensors (pinned host memory) on a separate copy stream, so that it overlaps the
        previous step's kernels; returns the device tensors and the event the compute stream must wait for."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        with torch.cuda.stream(self._copy_stream):
            u4 = u4_host.to(self.dev, non_blocking=True)
            trans = trans_host.to(self.dev, non_blocking=True)
            noise = noise_host.to(self.dev, non_blocking=True) if self.schedule(i)[3] else None
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return u4, trans, noise, ev

    def step_from_host(self, i: int, x_host: torch.Tensor, u4_host, trans_host, noise_host, out_host: torch.Tensor,
                       uploaded=None):
        """Same step with every per-step input coming from pinned host memory and the result read back
        (bench.py's end-to-end leg): H2D(x, randoms) -> step -> D2H(x_next).  `uploaded` = the result of an earlier
        `upload_randoms(i, ...)` (the randoms do not depend on the previous step, only x does)."""
        self.x_next.copy_(x_host, non_blocking=True)
        if uploaded is None:
            uploaded = self.upload_randoms(i, u4_host, trans_host, noise_host)
        u4, trans, noise, ev = uploaded
        torch.cuda.current_stream(self.dev).wait_event(ev)
        for t in (u4, trans, noise):
            if t is not None:
                t.record_stream(torch.cuda.current_stream(self.dev))
        x = self.step(i, randoms=(u4, trans, noise))
        out_host.copy_(x, non_blocking=True)
        return out_host


@torch.no_grad()
def sample_diffusion(dit: B200DiT, batch: Dict[str, torch.Tensor], a, ap, s, z, num_sample: int = 5,
                     steps: int = 200, gamma_0: float = 0.8, gamma_min: float = 1.0,
                     noise_scale_lambda: float = 1.003, step_scale_eta: float = 1.5,
                     ode_step_scale_eta: float = 1.0, ref_mol=None, ref_mol_poses: Optional[torch.Tensor] = None,
                     use_ref_mol_poses: bool = False, mmff_gamma_0_factor: float = 1.0, mmff_iters: int = 5,
                     align_ref_pos: bool = True, karras_noise_schedule_power: float = 7, rng=None,
                     mmff_fn: Optional[Callable] = None, trace: Optional[List[dict]] = None,
                     teacher: Optional[List[dict]] = None, max_steps: Optional[int] = None, physics_field=None,
                     physics_step: float = 0.002, physics_gmax: float = 50.0,
                     conformer_fn: Optional[Callable] = None, use_cuda_graph: Optional[bool] = None) -> torch.Tensor:
    """`PhysDock.sample_diffusion` (model.py:157-282) given the trunk outputs (a, ap, s, z).

    Extra hooks (not in the reference): `rng` (object with rand/normal), `mmff_fn`, `conformer_fn` (host hooks for the two
    RDKit calls, MMFFOptimizeMolecule and EmbedMultipleConfs), `physics_field` (a
    physics.PairEnergyField: GPU replacement of the MMFF step), `trace` (list that
    receives per-step tensors), `teacher` (per-step dicts with an `x_hat` to feed the denoiser instead of the
    free-running one: teacher-forced parity).
    """
    smp = DiffusionSampler(dit, batch, a, ap, s, z, num_sample=num_sample, steps=steps, gamma_0=gamma_0,
                           gamma_min=gamma_min, noise_scale_lambda=noise_scale_lambda, step_scale_eta=step_scale_eta,
                           ode_step_scale_eta=ode_step_scale_eta, ref_mol=ref_mol, ref_mol_poses=ref_mol_poses,
                           use_ref_mol_poses=use_ref_mol_poses, mmff_gamma_0_factor=mmff_gamma_0_factor,
                           mmff_iters=mmff_iters, align_ref_pos=align_ref_pos,
                           karras_noise_schedule_power=karras_noise_schedule_power, rng=rng, mmff_fn=mmff_fn,
                           physics_field=physics_field, physics_step=physics_step, physics_gmax=physics_gmax,
                           conformer_fn=conformer_fn, use_cuda_graph=use_cuda_graph)
    x = smp.begin()
    for i in range(steps):
        if max_steps is not None and i >= max_steps:
            break
        th = None if teacher is None else teacher[i]["x_hat"].to(smp.dev)
        x = smp.step(i, teacher_x_hat=th)
        if trace is not None:
            sch = smp.schedule(i)
            trace.append(dict(i=i, t_cur=float(sch[0]), t_hat=float(sch[2]), x_hat=smp.x_hat.clone(),
                              x_denoised=smp.x_den.clone(), x_next=x.clone(),
                              used_inds=None if smp.last_used is None else smp.last_used.clone()))
    return x


class PhysDockB200(nn.Module):
    """Mirror of the reference model container `PhysDock` (models/model.py:55-68) for inference.

    `diffusion_conditioning` is the once-per-complex trunk (out of scope here: pass the reference's PyTorch
    module, or any callable batch -> (a, ap, s, z)); `dit` is the B200 denoiser.  Built from a reference
    model with `PhysDockB200.from_reference(model)`; `sample_diffusion` keeps the keyword set used at
    redocking.py:284-299.
    """

    def __init__(self, dit: B200DiT, diffusion_conditioning: Optional[Callable] = None, sigma_data: float = 16.0):
        super().__init__()
        self.dit = dit
        self.diffusion_conditioning = diffusion_conditioning
        self.sigma_data = sigma_data

    @classmethod
    def from_reference(cls, ref_model: nn.Module) -> "PhysDockB200":
        return cls(B200DiT.from_reference(ref_model.dit), ref_model.diffusion_conditioning,
                   getattr(ref_model, "sigma_data", 16.0))

    karras_noise_schedule = staticmethod(karras_noise_schedule)

    @torch.no_grad()
    def sample_diffusion(self, batch, num_sample: int = 5, steps: int = 200, gamma_0: float = 0.8,
                         gamma_min: float = 1.0, noise_scale_lambda: float = 1.003, step_scale_eta: float = 1.5,
                         ode_step_scale_eta=1.0, ref_mol=None, ref_mol_poses=None, use_ref_mol_poses=False,
                         mmff_gamma_0_factor=1.0, mmff_iters=5, align_ref_pos=True, karras_noise_schedule_power=7,
                         conditioning=None, **hooks) -> torch.Tensor:
        if conditioning is None:
            if self.diffusion_conditioning is None:
                raise _lib.PdkError("no trunk: pass conditioning=(a, ap, s, z) or set diffusion_conditioning")
            conditioning = self.diffusion_conditioning(batch)       # model.py:144
        a, ap, s, z = conditioning
        return sample_diffusion(self.dit, batch, a, ap, s, z, num_sample=num_sample, steps=steps, gamma_0=gamma_0,
                                gamma_min=gamma_min, noise_scale_lambda=noise_scale_lambda,
                                step_scale_eta=step_scale_eta, ode_step_scale_eta=ode_step_scale_eta, ref_mol=ref_mol,
                                ref_mol_poses=ref_mol_poses, use_ref_mol_poses=use_ref_mol_poses,
                                mmff_gamma_0_factor=mmff_gamma_0_factor, mmff_iters=mmff_iters,
                                align_ref_pos=align_ref_pos, karras_noise_schedule_power=karras_noise_schedule_power,
                                **hooks)
