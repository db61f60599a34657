"""B200DiT -- drop-in replacement for the reference denoiser `AF3DiT`
(PhysDock/models/layers/transformers.py:178-262).

Same constructor arguments, same `forward(batch, x_hat, t_hat, a, ap, s, z) -> x_denoised` contract, same
`state_dict()` keys/shapes (so PhysDock/utils/import_weights.py:31-41 keeps working), but the arithmetic
runs in the hand-written sm_100a kernels behind include/physdock_b200.h:

    model.dit = B200DiT.from_reference(model.dit)        # the whole drop-in

What is cached and when:
  * weights  -> re-laid-out once per parameter version (`_pack`): q|k|v concatenated, w1/w3 interleaved in blocks of 16,
    every matrix split into fp16 hi/lo planes, all 36 AdaLN-Zero linears concatenated, LayerNorm(z) affine
    folded into linear_z.  `load_state_dict`, `.to()` / `.cuda()` / `.float()` bump a version counter (O(1) check per
    call); the full per-parameter scan runs once per complex.  After modifying a parameter IN PLACE call
    `mark_weights_changed()`.
  * complex  -> `prepare_complex` runs once per (a, ap, s, z, masks): the pair-bias of every block
    ([6,4,Sa,Sa] + [12,16,St,St] fp32) that the reference recomputes in every block of every step.
  * launches -> `forward` replays the ~120 kernels of one denoiser call from a CUDA graph (inputs are copied into
    persistent buffers); the sampler uses `denoise_cond_graphed` with the conditioning of the whole schedule
    precomputed (`conditioning_table`).
There is no CPU / PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from .synthetic import DiTDims, dit_param_shapes


def _split_planes(w: torch.Tensor):
    """fp32 -> (hi, lo) fp16 planes: hi = fp16(w), lo = fp16(w - hi)."""
    w = w.detach().float().contiguous()
    hi = w.half()
    lo = (w - hi.float()).half()
    return hi.contiguous(), lo.contiguous()


def _interleave16(w1: torch.Tensor, w3: torch.Tensor) -> torch.Tensor:
    """[hid,c],[hid,c] -> [2*hid,c] with rows in blocks of 16: w1[0:16], w3[0:16], w1[16:32], ... so that one
    32-column accumulator chunk of the GEMM epilogue holds h1 and h3 of the same 16 hidden channels."""
    hid, c = w1.shape
    return torch.stack([w1.reshape(hid // 16, 16, c), w3.reshape(hid // 16, 16, c)], dim=1).reshape(2 * hid, c)


class B200DiT(nn.Module):
    def __init__(self, c_a: int = 128, c_ap: int = 16, c_s: int = 512, c_z: int = 128, inf: float = 1e9,
                 eps: float = 1e-8, no_blocks_atom: int = 3, no_blocks_dit: int = 12, sigma_data: float = 16.0):
        super().__init__()
        self.dims = DiTDims(c_a=c_a, c_ap=c_ap, c_s=c_s, c_z=c_z, no_blocks_atom=no_blocks_atom,
                            no_blocks_dit=no_blocks_dit, sigma_data=float(sigma_data), inf=float(inf),
                            eps=float(eps))
        self.sigma_data = sigma_data
        # parameters registered under the reference's exact dotted names, in its registration order
        for key, shape in dit_param_shapes(self.dims).items():
            mod: nn.Module = self
            *path, leaf = key.split(".")
            for name in path:
                if name not in mod._modules:
                    mod.add_module(name, nn.Module())
                mod = mod._modules[name]
            mod.register_parameter(leaf, nn.Parameter(torch.zeros(shape), requires_grad=False))
        self._handle = None
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._pack_sig = None
        self._pack_version = -1
        self._weights_version = 0
        self._complex_sig = None
        self._complex_keep = None
        self._complex_token = 0
        self._workspace: Optional[torch.Tensor] = None
        self._graphs: Dict[tuple, object] = {}
        self._block_array = None
        self._fwd_static: Dict[tuple, tuple] = {}
        self._bias_bufs = None
        self.use_cuda_graph = True

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_reference(cls, ref_dit: nn.Module) -> "B200DiT":
        """Builds a B200DiT carrying the weights (and device) of a reference `AF3DiT` instance."""
        sd = ref_dit.state_dict()
        n_atom = len({k.split(".")[2] for k in sd if k.startswith("atom_dit_encoder.blocks.")})
        n_tok = len({k.split(".")[2] for k in sd if k.startswith("token_dit.blocks.")})
        new = cls(c_a=sd["linear_x.weight"].shape[0], c_ap=sd["atom_dit_encoder.blocks.0.attention.norm_z.weight"].shape[0],
                  c_s=sd["linear_downscale.weight"].shape[0], c_z=sd["token_dit.blocks.0.attention.norm_z.weight"].shape[0],
                  no_blocks_atom=n_atom, no_blocks_dit=n_tok, sigma_data=getattr(ref_dit, "sigma_data", 16.0))
        new.load_state_dict(sd)
        return new.to(next(iter(sd.values())).device)

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], dims: DiTDims = DiTDims(), device="cuda") -> "B200DiT":
        new = cls(c_a=dims.c_a, c_ap=dims.c_ap, c_s=dims.c_s, c_z=dims.c_z, inf=dims.inf, eps=dims.eps,
                  no_blocks_atom=dims.no_blocks_atom, no_blocks_dit=dims.no_blocks_dit, sigma_data=dims.sigma_data)
        new.load_state_dict(sd)
        return new.to(device)

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().pdk_dit_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def mark_weights_changed(self) -> None:
        """Call after modifying a parameter in place (p.data.copy_(...)): forces a re-pack on the next call."""
        self._weights_version += 1

    def _apply(self, fn, *a, **kw):          # .to() / .cuda() / .float() / .half(): parameters are replaced
        out = super()._apply(fn, *a, **kw)
        self._weights_version += 1
        return out

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)
        self._weights_version += 1
        return out

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _pack_fast(self) -> None:
        """O(1) per-call check (the 319-tensor signature scan costs ~100 us of host time): re-pack only when the version
        counter moved.  `prepare_complex` still runs the full scan once per complex."""
        if self._packed is None or self._pack_version != self._weights_version:
            self._pack()

    def _blocks_in_order(self):
        d = self.dims
        return ([("atom_dit_encoder", i, d.c_a) for i in range(d.no_blocks_atom)] +
                [("token_dit", i, d.c_s) for i in range(d.no_blocks_dit)] +
                [("atom_dit_decoder", i, d.c_a) for i in range(d.no_blocks_atom)])

    def _fold_pair_bias(self, sd, stacks, c_pair):
        """LayerNorm(z) affine folded into linear_z for all blocks of the given stacks:
        Linear(LN(x)) = sum_c (W[h,c] g[c]) xhat[c] + sum_c W[h,c] b[c]   (attentions.py:246,254)."""
        ws, bs = [], []
        for stack, i in stacks:
            p = f"{stack}.blocks.{i}.attention."
            W = sd[p + "linear_z.weight"].double()
            g, b = sd[p + "norm_z.weight"].double(), sd[p + "norm_z.bias"].double()
            ws.append(W * g[None, :])
            bs.append(W @ b)
        wfold = torch.cat(ws, 0)                    # [L*H, c_pair]
        return wfold.t().contiguous().float(), torch.cat(bs, 0).float().contiguous()

    def _pack(self):
        sig = self._signature()
        if self._packed is not None and sig == self._pack_sig:
            self._pack_version = self._weights_version
            return
        lib = _lib.load()
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.PdkError("B200DiT must live on a CUDA device (no CPU fallback)")
        d = self.dims
        sd = {k: v.detach().float() for k, v in self.state_dict().items()}
        P: Dict[str, torch.Tensor] = {}
        half = 128
        exponent = -math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32)   # timestep_embeddings.py:62-66
        P["freq"] = torch.exp(exponent / (half - 0)).to(dev)
        te = "time_embedder.timestep_embedder."
        for n, k in (("te_w1", "linear_1.weight"), ("te_b1", "linear_1.bias"), ("te_w2", "linear_2.weight"),
                     ("te_b2", "linear_2.bias")):
            P[n] = sd[te + k].contiguous()
        # every AdaLayerNormZero.linear of the model, concatenated along the output dimension
        mods_w, mods_b, offs, off = [], [], {}, 0
        for stack, i, c in self._blocks_in_order():
            for sub in ("attention.norm_s", "transition.ffn_norm"):
                p = f"{stack}.blocks.{i}.{sub}.linear."
                mods_w.append(sd[p + "weight"])
                mods_b.append(sd[p + "bias"])
                offs[(stack, i, sub)] = off
                off += 3 * c
        P["wmod"], P["bmod"] = torch.cat(mods_w, 0).contiguous(), torch.cat(mods_b, 0).contiguous()
        P["wmod_h"], P["wmod_l"] = _split_planes(P["wmod"])     # the 36 modulations run as one tensor-core GEMM
        n_mod = off
        P["wx"], P["bx"] = sd["linear_x.weight"].contiguous(), sd["linear_x.bias"].contiguous()
        P["wdown_h"], P["wdown_l"] = _split_planes(sd["linear_downscale.weight"])
        P["bdown"] = sd["linear_downscale.bias"].contiguous()
        P["wup_h"], P["wup_l"] = _split_planes(sd["linear_upscale.weight"])
        P["bup"] = sd["linear_upscale.bias"].contiguous()
        P["norm_r_w"], P["norm_r_b"] = sd["norm_r.weight"].contiguous(), sd["norm_r.bias"].contiguous()
        P["wr"] = sd["linear_r.weight"].contiguous()
        atom_stacks = [("atom_dit_encoder", i) for i in range(d.no_blocks_atom)] + \
                      [("atom_dit_decoder", i) for i in range(d.no_blocks_atom)]
        P["wz_atom_T"], P["bz_atom"] = self._fold_pair_bias(sd, atom_stacks, d.c_ap)
        P["wz_tok_T"], P["bz_tok"] = self._fold_pair_bias(sd, [("token_dit", i) for i in range(d.no_blocks_dit)], d.c_z)
        blocks = (_lib.BlockWeights * len(self._blocks_in_order()))()
        for bi, (stack, i, c) in enumerate(self._blocks_in_order()):
            p = f"{stack}.blocks.{i}."
            a, f = p + "attention.", p + "transition.feed_forward."
            t = {}
            t["wqkv_h"], t["wqkv_l"] = _split_planes(torch.cat([sd[a + "linear_q.weight"], sd[a + "linear_k.weight"],
                                                                sd[a + "linear_v.weight"]], 0))
            t["wo_h"], t["wo_l"] = _split_planes(sd[a + "linear_o.weight"])
            t["w13_h"], t["w13_l"] = _split_planes(_interleave16(sd[f + "w1.weight"], sd[f + "w3.weight"]))
            t["w2_h"], t["w2_l"] = _split_planes(sd[f + "w2.weight"])
            t["bo"] = sd[a + "linear_o.bias"].contiguous()
            t["norm_q"], t["norm_k"] = sd[a + "norm_q.weight"].contiguous(), sd[a + "norm_k.weight"].contiguous()
            for n, v in t.items():
                P[f"b{bi}.{n}"] = v
                setattr(blocks[bi], n, v.data_ptr())
            blocks[bi].mod_attn_off = offs[(stack, i, "attention.norm_s")]
            blocks[bi].mod_ffn_off = offs[(stack, i, "transition.ffn_norm")]
        for k, v in P.items():
            assert v.is_cuda and v.is_contiguous(), k
        if self._handle is None:
            dims = _lib.DitDims(c_a=d.c_a, c_ap=d.c_ap, c_s=d.c_s, c_z=d.c_z, n_atom_blocks=d.no_blocks_atom,
                                n_token_blocks=d.no_blocks_dit, hidden_a=d.ffn_hidden(d.c_a),
                                hidden_s=d.ffn_hidden(d.c_s), n_mod=n_mod, sigma_data=d.sigma_data, eps=d.eps,
                                inf=d.inf)
            h = C.c_void_p()
            _lib.check(lib.pdk_dit_create(C.byref(dims), C.byref(h)), "pdk_dit_create")
            self._handle = h
        W = _lib.DitWeights()
        for name, _ in _lib.DitWeights._fields_:
            if name in P:
                setattr(W, name, P[name].data_ptr())
        W.blocks = C.cast(blocks, C.POINTER(_lib.BlockWeights))
        W.n_blocks = len(blocks)
        _lib.check(lib.pdk_dit_set_weights(self._handle, C.byref(W)), "pdk_dit_set_weights")
        self._packed, self._pack_sig, self._block_array = P, sig, blocks
        self._pack_version = self._weights_version
        self._n_mod = n_mod
        self._complex_sig = None      # bias caches depend on the weights
        self._complex_keep = None
        self._graphs.clear()          # captured graphs hold pointers into the old packed weights

    # ------------------------------------------------------------------ per-complex cache
    def prepare_complex(self, batch: Dict[str, torch.Tensor], a: torch.Tensor, ap: torch.Tensor, s: torch.Tensor,
                        z: torch.Tensor) -> None:
        """Caches the per-complex pair biases and index maps (runs the pair-bias prepass kernels)."""
        self._pack()
        lib = _lib.load()
        dev = ap.device
        Na, Nt = a.shape[0], s.shape[0]
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()   # noqa: E731
        a_, ap_, s_, z_ = f32(a), f32(ap), f32(s), f32(z)
        apm, zm = f32(batch["ap_mask"]), f32(batch["z_mask"])
        chunk = batch["token_id_to_chunk_sizes"].to(dev).long()
        tok_start = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(chunk, 0)]).int().contiguous()
        atom2tok = batch["atom_id_to_token_id"].to(dev).int().contiguous()
        if int(tok_start[-1]) != Na or atom2tok.numel() != Na or chunk.numel() != Nt:
            raise _lib.PdkError("token_id_to_chunk_sizes / atom_id_to_token_id inconsistent with a, s")
        ab, tb = C.c_size_t(), C.c_size_t()
        _lib.check(lib.pdk_dit_bias_bytes(self._handle, Na, Nt, C.byref(ab), C.byref(tb)), "pdk_dit_bias_bytes")
        # the two bias caches (403 MB + 50 MB at 2048 / 256) are persistent buffers of this module, reused for the next complex
        # when they are large enough (a fresh 453 MB allocation per screening ligand cost milliseconds of cudaMalloc / free)
        self._graphs.clear()                  # captured graphs read the buffers about to be overwritten
        bufs = self._bias_bufs
        if bufs is None or bufs[0].numel() < ab.value // 4 or bufs[1].numel() < tb.value // 4 or bufs[0].device != dev:
            self._bias_bufs = bufs = None
            self._complex_keep = None         # release the previous caches before allocating the new ones
            bufs = self._bias_bufs = (torch.empty(ab.value // 4, dtype=torch.float32, device=dev),
                                      torch.empty(tb.value // 4, dtype=torch.float32, device=dev))
        bias_a, bias_t = bufs[0][: ab.value // 4], bufs[1][: tb.value // 4]
        _lib.check(lib.pdk_dit_prepare_complex(self._handle, _lib.ptr(a_), _lib.ptr(ap_), _lib.ptr(s_), _lib.ptr(z_),
                                               _lib.ptr(apm), _lib.ptr(zm), _lib.ptr(tok_start), _lib.ptr(atom2tok),
                                               Na, Nt, _lib.ptr(bias_a), _lib.ptr(bias_t), _lib.stream_ptr(dev)),
                   "pdk_dit_prepare_complex")
        # the handle keeps raw pointers into a_, s_, tok_start, atom2tok and the bias caches; `sig_tensors` keeps every
        # tensor the complex signature was taken from alive, so that a later complex cannot be handed the same addresses
        # (which would make its signature collide with this one's)
        self._complex_token += 1
        self._complex_keep = dict(a=a_, s=s_, tok_start=tok_start, atom2tok=atom2tok, bias_a=bias_a, bias_t=bias_t,
                                  Na=Na, Nt=Nt, token=self._complex_token,
                                  sig_tensors=(a, ap, s, z, batch["ap_mask"], batch["z_mask"],
                                               batch["token_id_to_chunk_sizes"], batch["atom_id_to_token_id"]))
        self._graphs.clear()

    @staticmethod
    def _complex_signature(batch, a, ap, s, z):
        ts = (a, ap, s, z, batch["ap_mask"], batch["z_mask"], batch["token_id_to_chunk_sizes"],
              batch["atom_id_to_token_id"])
        # inference-mode tensors (a trunk run under torch.inference_mode) carry no version counter: they are immutable
        return tuple((t.data_ptr(), -1 if t.is_inference() else t._version, tuple(t.shape)) for t in ts)

    def _ensure_workspace(self, B: int, dev) -> torch.Tensor:
        lib = _lib.load()
        need = C.c_size_t()
        k = self._complex_keep
        _lib.check(lib.pdk_dit_workspace_bytes(self._handle, B, k["Na"], k["Nt"], C.byref(need)), "pdk_dit_workspace_bytes")
        if self._workspace is None or self._workspace.numel() < need.value or self._workspace.device != dev:
            # captured graphs hold raw pointers into the old buffer: drop them BEFORE it goes back to the allocator
            self._graphs.clear()
            self._workspace = None
            self._workspace = torch.empty(need.value, dtype=torch.uint8, device=dev)
        return self._workspace

    def _check_device(self, t: torch.Tensor) -> None:
        if not t.is_cuda:
            raise _lib.PdkError("physdock_b200 kernels need CUDA tensors (no CPU fallback)")
        if t.device.index != torch.cuda.current_device():
            raise _lib.PdkError(f"tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                                "wrap the call in torch.cuda.device(...) (one process per GPU is the intended use)")

    # ------------------------------------------------------------------ AF3DiT.forward
    def denoise(self, x_hat: torch.Tensor, t_hat: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_denoised for the prepared complex.  x_hat [B,Na,3], t_hat [B] (fp32, CUDA)."""
        if self._complex_keep is None:
            raise _lib.PdkError("prepare_complex() has not been called")
        lib = _lib.load()
        self._check_device(x_hat)
        B = x_hat.shape[0]
        dev = x_hat.device
        x_hat = x_hat.float().contiguous()
        t_hat = t_hat.to(device=dev, dtype=torch.float32).contiguous()
        if x_hat.shape[1] != self._complex_keep["Na"] or t_hat.numel() != B:
            raise _lib.PdkError("x_hat / t_hat shape does not match the prepared complex")
        ws = self._ensure_workspace(B, dev)
        if out is None:
            out = torch.empty_like(x_hat)
        _lib.check(lib.pdk_dit_denoise(self._handle, _lib.ptr(x_hat), _lib.ptr(t_hat), B, _lib.ptr(ws), ws.numel(),
                                       _lib.ptr(out), _lib.stream_ptr(dev)), "pdk_dit_denoise")
        return out

    def _graph_key(self, *tensors):
        return tuple(0 if t is None else t.data_ptr() for t in tensors) + (tuple(tensors[0].shape), self._complex_keep["token"],
                                                                            self._pack_version, self._workspace.data_ptr())

    def _replay(self, key, enqueue):
        """Runs `enqueue()` (which only enqueues library calls with persistent pointer arguments) from a CUDA graph."""
        g = self._graphs.get(key)
        if g is None:
            enqueue()                                 # warm-up: tensor maps, function attributes
            torch.cuda.synchronize()
            if len(self._graphs) > 16:
                self._graphs.clear()
            g = torch.cuda.CUDAGraph()
            # thread_local: another host thread may be driving its own streams meanwhile (pipeline.prefetch_complexes runs the
            # next complex's trunk on a side stream); only this thread's calls are part of / restricted by the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                enqueue()
            self._graphs[key] = g
        g.replay()

    def denoise_graphed(self, x_hat: torch.Tensor, t_hat: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """`denoise` replayed from a CUDA graph (one graph per set of argument buffers, workspace and prepared complex).

        The call enqueues ~124 kernels whose arguments are all pointers into persistent buffers, so the whole
        denoiser is captured once and replayed with a single launch; x_hat / t_hat / out must be persistent
        tensors that the caller updates in place."""
        if not (x_hat.is_contiguous() and t_hat.is_contiguous() and out.is_contiguous()
                and x_hat.dtype == t_hat.dtype == out.dtype == torch.float32) or not self.use_cuda_graph:
            return self.denoise(x_hat, t_hat, out)
        self._ensure_workspace(x_hat.shape[0], x_hat.device)
        self._replay(self._graph_key(x_hat, t_hat, out), lambda: self.denoise(x_hat, t_hat, out))
        return out

    # ------------------------------------------------------------------ conditioning hoisted out of the step
    def cond_width(self) -> int:
        self._pack_fast()
        return int(_lib.load().pdk_dit_cond_width(self._handle))

    def conditioning_table(self, t_hat: torch.Tensor) -> torch.Tensor:
        """Rows [modulations (n_mod) | c_in, c_skip, c_out, t_hat | t_next, eta, 0, 0] for every noise level of `t_hat` [n]
        (fp32, CUDA): everything AF3DiT derives from t_hat alone, computed once per schedule instead of once per step.
        Returns a [n, cond_width] view of the table; entries 4..7 of the coefficient block are left for the caller."""
        self._pack_fast()
        lib = _lib.load()
        t_hat = t_hat.float().contiguous()
        self._check_device(t_hat)
        n, dev, width = t_hat.numel(), t_hat.device, self.cond_width()
        need = C.c_size_t()
        _lib.check(lib.pdk_dit_conditioning_workspace_bytes(self._handle, n, C.byref(need)), "pdk_dit_conditioning_workspace_bytes")
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        table = torch.zeros(int(lib.pdk_pad_len(n)), width, dtype=torch.float32, device=dev)
        _lib.check(lib.pdk_dit_conditioning(self._handle, _lib.ptr(t_hat), n, _lib.ptr(ws), ws.numel(), _lib.ptr(table), width,
                                            _lib.stream_ptr(dev)), "pdk_dit_conditioning")
        ws.record_stream(torch.cuda.current_stream(dev))
        return table[:n]

    def denoise_cond(self, x_hat: torch.Tensor, cond: torch.Tensor, out: torch.Tensor,
                     x_next: Optional[torch.Tensor] = None) -> torch.Tensor:
        """AF3DiT for the prepared complex with precomputed conditioning: `cond` is one row [cond_width] shared by all
        samples, or [B, cond_width].  With `x_next` the physics-free Euler update is written by the last kernel."""
        if self._complex_keep is None:
            raise _lib.PdkError("prepare_complex() has not been called")
        lib = _lib.load()
        self._check_device(x_hat)
        B, dev = x_hat.shape[0], x_hat.device
        if cond.dim() == 1:
            stride = 0
        else:
            if cond.shape[0] != B:
                raise _lib.PdkError("cond must be one row or one row per sample")
            stride = cond.stride(0)
        if cond.stride(-1) != 1 or cond.shape[-1] != self.cond_width() or cond.dtype != torch.float32:
            raise _lib.PdkError("cond rows must be contiguous fp32 of width cond_width()")
        if x_hat.shape[1] != self._complex_keep["Na"] or not x_hat.is_contiguous() or x_hat.dtype != torch.float32:
            raise _lib.PdkError("x_hat must be contiguous fp32 [B, Na, 3] of the prepared complex")
        ws = self._ensure_workspace(B, dev)
        _lib.check(lib.pdk_dit_denoise_cond(self._handle, x_hat.data_ptr(), cond.data_ptr(), stride, B, _lib.ptr(ws), ws.numel(),
                                            _lib.ptr(out), _lib.ptr(x_next), _lib.stream_ptr(dev)), "pdk_dit_denoise_cond")
        return out

    def denoise_cond_graphed(self, x_hat, cond, out, x_next=None):
        """`denoise_cond` replayed from a CUDA graph; all four tensors must be persistent buffers updated in place."""
        if not self.use_cuda_graph:
            return self.denoise_cond(x_hat, cond, out, x_next)
        self._ensure_workspace(x_hat.shape[0], x_hat.device)
        self._replay(self._graph_key(x_hat, cond, out, x_next), lambda: self.denoise_cond(x_hat, cond, out, x_next))
        return out

    def forward(self, batch: Dict[str, torch.Tensor], x_hat: torch.Tensor, t_hat: torch.Tensor, a: torch.Tensor,
                ap: torch.Tensor, s: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        """Same contract as AF3DiT.forward (transformers.py:235-262); leading batch dims other than the
        sample dimension are not supported (the reference sampler never uses them).

        This is the call the reference sampler makes every step through `partial(self.dit, batch=..., a=..., ap=..., s=...,
        z=...)(x_hat=..., t_hat=...)` (model.py:153,221): inputs are copied into persistent buffers, the denoiser is
        replayed from a CUDA graph, and a fresh tensor is returned."""
        self._pack_fast()
        sig = self._complex_signature(batch, a, ap, s, z)
        if sig != self._complex_sig:
            self._pack()                      # full parameter scan, once per complex
            self.prepare_complex(batch, a, ap, s, z)
            self._complex_sig = sig
        self._check_device(x_hat)
        B = x_hat.shape[0]
        t_hat = t_hat.to(device=x_hat.device, dtype=torch.float32).reshape(-1)
        if t_hat.numel() == 1 and B > 1:
            t_hat = t_hat.expand(B)
        if not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
            return self.denoise(x_hat, t_hat)
        key = (B, x_hat.shape[1], x_hat.device.index)
        st = self._fwd_static.get(key)
        if st is None:
            if len(self._fwd_static) > 4:
                self._fwd_static.clear()
            st = self._fwd_static[key] = (torch.empty(B, x_hat.shape[1], 3, dtype=torch.float32, device=x_hat.device),
                                          torch.empty(B, dtype=torch.float32, device=x_hat.device),
                                          torch.empty(B, x_hat.shape[1], 3, dtype=torch.float32, device=x_hat.device))
        xs, ts, outs = st
        xs.copy_(x_hat)
        ts.copy_(t_hat)
        self.denoise_graphed(xs, ts, outs)
        return outs.clone()

    def launches_per_denoise(self, cond: bool = False) -> int:
        """Kernels per denoiser call (`cond`: with the conditioning precomputed, as the sampler runs it)."""
        self._pack_fast()
        lib = _lib.load()
        return int((lib.pdk_dit_launches_per_denoise_cond if cond else lib.pdk_dit_launches_per_denoise)(self._handle))
