"""B200DiT -- drop-in replacement for the reference denoiser `AF3DiT`
(PhysDock/models/layers/transformers.py:178-262).

Same constructor arguments, same `forward(batch, x_hat, t_hat, a, ap, s, z) -> x_denoised` contract, same
`state_dict()` keys/shapes (so PhysDock/utils/import_weights.py:31-41 keeps working), but the arithmetic
runs in the hand-written sm_100a kernels behind include/physdock_b200.h:

    model.dit = B200DiT.from_reference(model.dit)        # the whole drop-in

What is cached and when:
  * weights  -> re-laid-out once per parameter version (`_pack`): q|k|v concatenated, w1/w3 interleaved in blocks of 16,
    every matrix split into fp16 hi/lo planes, all 36 AdaLN-Zero linears concatenated, LayerNorm(z) affine
    folded into linear_z.
  * complex  -> `prepare_complex` runs once per (a, ap, s, z, masks): the pair-bias of every block
    ([6,4,Sa,Sa] + [12,16,St,St] fp32) that the reference recomputes in every block of every step.
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
        self._complex_sig = None
        self._complex_keep = None
        self._workspace: Optional[torch.Tensor] = None
        self._graphs: Dict[tuple, object] = {}
        self._block_array = None

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
    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

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
        self._n_mod = n_mod
        self._complex_sig = None      # bias caches depend on the weights

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
        bias_a = torch.empty(ab.value // 4, dtype=torch.float32, device=dev)
        bias_t = torch.empty(tb.value // 4, dtype=torch.float32, device=dev)
        _lib.check(lib.pdk_dit_prepare_complex(self._handle, _lib.ptr(a_), _lib.ptr(ap_), _lib.ptr(s_), _lib.ptr(z_),
                                               _lib.ptr(apm), _lib.ptr(zm), _lib.ptr(tok_start), _lib.ptr(atom2tok),
                                               Na, Nt, _lib.ptr(bias_a), _lib.ptr(bias_t), _lib.stream_ptr(dev)),
                   "pdk_dit_prepare_complex")
        # the handle keeps raw pointers into these
        self._complex_keep = dict(a=a_, s=s_, tok_start=tok_start, atom2tok=atom2tok, bias_a=bias_a, bias_t=bias_t,
                                  Na=Na, Nt=Nt)
        self._graphs.clear()

    @staticmethod
    def _complex_signature(batch, a, ap, s, z):
        ts = (a, ap, s, z, batch["ap_mask"], batch["z_mask"], batch["token_id_to_chunk_sizes"],
              batch["atom_id_to_token_id"])
        return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in ts)

    def _ensure_workspace(self, B: int, dev) -> torch.Tensor:
        lib = _lib.load()
        need = C.c_size_t()
        k = self._complex_keep
        _lib.check(lib.pdk_dit_workspace_bytes(self._handle, B, k["Na"], k["Nt"], C.byref(need)), "pdk_dit_workspace_bytes")
        if self._workspace is None or self._workspace.numel() < need.value or self._workspace.device != dev:
            self._workspace = torch.empty(need.value, dtype=torch.uint8, device=dev)
        return self._workspace

    # ------------------------------------------------------------------ AF3DiT.forward
    def denoise(self, x_hat: torch.Tensor, t_hat: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_denoised for the prepared complex.  x_hat [B,Na,3], t_hat [B] (fp32, CUDA)."""
        if self._complex_keep is None:
            raise _lib.PdkError("prepare_complex() has not been called")
        lib = _lib.load()
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

    def denoise_graphed(self, x_hat: torch.Tensor, t_hat: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """`denoise` replayed from a CUDA graph (one graph per set of argument buffers and prepared complex).

        The call enqueues ~136 kernels whose arguments are all pointers into persistent buffers, so the whole
        denoiser is captured once and replayed with a single launch; x_hat / t_hat / out must be persistent
        tensors that the caller updates in place (DiffusionSampler does)."""
        if not (x_hat.is_contiguous() and t_hat.is_contiguous() and out.is_contiguous()
                and x_hat.dtype == t_hat.dtype == out.dtype == torch.float32):
            return self.denoise(x_hat, t_hat, out)
        key = (x_hat.data_ptr(), t_hat.data_ptr(), out.data_ptr(), tuple(x_hat.shape), id(self._complex_keep),
               self._pack_sig is not None and hash(self._pack_sig))
        g = self._graphs.get(key)
        if g is None:
            self.denoise(x_hat, t_hat, out)          # warm-up: workspace, tensor maps, function attributes
            torch.cuda.synchronize(x_hat.device)
            if len(self._graphs) > 8:
                self._graphs.clear()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.denoise(x_hat, t_hat, out)
            self._graphs[key] = g
        g.replay()
        return out

    def forward(self, batch: Dict[str, torch.Tensor], x_hat: torch.Tensor, t_hat: torch.Tensor, a: torch.Tensor,
                ap: torch.Tensor, s: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        """Same contract as AF3DiT.forward (transformers.py:235-262); leading batch dims other than the
        sample dimension are not supported (the reference sampler never uses them)."""
        self._pack()
        sig = self._complex_signature(batch, a, ap, s, z)
        if sig != self._complex_sig:
            self.prepare_complex(batch, a, ap, s, z)
            self._complex_sig = sig
        return self.denoise(x_hat, t_hat)

    def launches_per_denoise(self) -> int:
        self._pack()
        return int(_lib.load().pdk_dit_launches_per_denoise(self._handle))
