"""TEST HARNESS (not part of the product package): thin Python wrappers over the op-level C entry points (pdk_op_*), one
kernel each, so that the parity tests (and the timing tools under tools/) can drive every kernel in isolation through the
C ABI.  The product path (`B200DiT.denoise*`) drives the same launchers from C++ in a single call.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from physdock_b200 import _lib

LOG2E = 1.4426950408889634


def pad_len(n: int) -> int:
    return (n + 127) // 128 * 128


def split_planes(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device-side fp32 -> (hi, lo) fp16 planes."""
    lib = _lib.load()
    x = x.float().contiguous()
    hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    _lib.check(lib.pdk_op_split(_lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), x.numel(), _lib.stream_ptr(x.device)), "split")
    return hi, lo


def planes_to_float(hi: torch.Tensor, lo: torch.Tensor) -> torch.Tensor:
    return hi.float() + lo.float()


def pair_bias(pair, mask, wfoldT, bfold, S_pad: int, ln_eps: float = 1e-5, inf: float = 1e9) -> torch.Tensor:
    lib = _lib.load()
    S, C = pair.shape[0], pair.shape[2]
    LH = bfold.numel()
    out = torch.empty(LH, S_pad, S_pad, dtype=torch.float32, device=pair.device)
    _lib.check(lib.pdk_op_pair_bias(_lib.ptr(pair.contiguous()), _lib.ptr(mask.contiguous()), _lib.ptr(wfoldT.contiguous()),
                                    _lib.ptr(bfold.contiguous()), _lib.ptr(out), S, S_pad, C, LH, ln_eps, inf,
                                    _lib.stream_ptr(pair.device)), "pair_bias")
    return out


def time_embed(t_hat, freq, w1, b1, w2, b2, sigma_data: float = 16.0):
    lib = _lib.load()
    B = t_hat.numel()
    tsilu = torch.empty(B, 256, dtype=torch.float32, device=t_hat.device)
    coef = torch.empty(B, 4, dtype=torch.float32, device=t_hat.device)
    _lib.check(lib.pdk_op_time_embed(_lib.ptr(t_hat.contiguous()), _lib.ptr(freq), _lib.ptr(w1.contiguous()),
                                     _lib.ptr(b1.contiguous()), _lib.ptr(w2.contiguous()), _lib.ptr(b2.contiguous()),
                                     sigma_data, _lib.ptr(tsilu), _lib.ptr(coef), B, _lib.stream_ptr(t_hat.device)),
               "time_embed")
    return tsilu, coef


def mod_gemv(tsilu, wmod, bmod):
    lib = _lib.load()
    B, n_mod = tsilu.shape[0], wmod.shape[0]
    mod = torch.empty(B, n_mod, dtype=torch.float32, device=tsilu.device)
    _lib.check(lib.pdk_op_mod_gemv(_lib.ptr(tsilu.contiguous()), _lib.ptr(wmod.contiguous()), _lib.ptr(bmod.contiguous()),
                                   _lib.ptr(mod), B, n_mod, _lib.stream_ptr(tsilu.device)), "mod_gemv")
    return mod


def adaln(x, mod, mod_off: int, eps: float):
    """x [B,S_pad,c] fp32, mod [B,n_mod] -> planes of LN(x)*(1+scale)+shift."""
    lib = _lib.load()
    B, S_pad, c = x.shape
    hi = torch.empty(B, S_pad, c, dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    _lib.check(lib.pdk_op_adaln(_lib.ptr(x.contiguous()), _lib.ptr(mod.contiguous()), mod.shape[1], mod_off, _lib.ptr(hi),
                                _lib.ptr(lo), B, S_pad, c, eps, _lib.stream_ptr(x.device)), "adaln")
    return hi, lo


def gemm_store(ah, al, wh, wl, bias: Optional[torch.Tensor] = None, silu: bool = False) -> torch.Tensor:
    lib = _lib.load()
    M, K = ah.shape
    N = wh.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device=ah.device)
    _lib.check(lib.pdk_op_gemm_store(_lib.ptr(ah), _lib.ptr(al), K, _lib.ptr(wh), _lib.ptr(wl), K, M, N, K,
                                     _lib.ptr(bias), int(silu), _lib.ptr(out), N, _lib.stream_ptr(ah.device)), "gemm_store")
    return out


def gemm_gate_resid(ah, al, wh, wl, bias, gate, gate_stride: int, rows_per_sample: int, x: torch.Tensor):
    """x (fp32 [M,N], in place) += (A W^T + bias) * gate[sample(row)]; `gate` may be a view into the mod rows."""
    lib = _lib.load()
    M, K = ah.shape
    N = wh.shape[0]
    _lib.check(lib.pdk_op_gemm_gate_resid(_lib.ptr(ah), _lib.ptr(al), K, _lib.ptr(wh), _lib.ptr(wl), K, M, N, K,
                                          _lib.ptr(bias), gate.data_ptr(), gate_stride, rows_per_sample, _lib.ptr(x),
                                          N, _lib.stream_ptr(ah.device)), "gemm_gate_resid")
    return x


def gemm_swiglu(ah, al, w13h, w13l):
    lib = _lib.load()
    M, K = ah.shape
    N = w13h.shape[0]
    ph = torch.empty(M, N // 2, dtype=torch.float16, device=ah.device)
    pl = torch.empty_like(ph)
    _lib.check(lib.pdk_op_gemm_swiglu(_lib.ptr(ah), _lib.ptr(al), K, _lib.ptr(w13h), _lib.ptr(w13l), K, M, N, K,
                                      _lib.ptr(ph), _lib.ptr(pl), N // 2, _lib.stream_ptr(ah.device)), "gemm_swiglu")
    return ph, pl


def transition_fused(x: torch.Tensor, mod: torch.Tensor, mod_off: int, w13h, w13l, w2h, w2l, rows_per_sample: int,
                     eps: float) -> torch.Tensor:
    """In place on x [M,128] fp32: the whole atom DiTTransition (AdaLN + SwiGLU + w2 + gate + residual) in one kernel."""
    lib = _lib.load()
    M, c = x.shape
    hidden = w2h.shape[1]
    _lib.check(lib.pdk_op_transition_fused(_lib.ptr(x), _lib.ptr(mod), mod.shape[1], mod_off, _lib.ptr(w13h), _lib.ptr(w13l),
                                           _lib.ptr(w2h), _lib.ptr(w2l), M, hidden, rows_per_sample, eps,
                                           _lib.stream_ptr(x.device)), "transition_fused")
    return x


def interleave_planes(x: torch.Tensor) -> torch.Tensor:
    """fp32 [..., 32] -> fp16 [..., 64] with rows [hi 32 | lo 32] (the q/k/v operand layout of the attention kernel)."""
    hi, lo = split_planes(x)
    return torch.cat([hi, lo], dim=-1).contiguous()


def deinterleave_to_float(t: torch.Tensor) -> torch.Tensor:
    return t[..., :32].float() + t[..., 32:].float()


def gemm_qkv(ah, al, wh, wl, norm_q, norm_k, rms_eps: float, B: int, S_pad: int):
    """-> q, k, v fp16 [B,H,S_pad,64] (rows [hi|lo]); q is RMS-normed and pre-scaled by log2e/sqrt(32), k RMS-normed."""
    lib = _lib.load()
    M, c = ah.shape
    H = c // 32
    outs = [torch.empty(B, H, S_pad, 64, dtype=torch.float16, device=ah.device) for _ in range(3)]
    _lib.check(lib.pdk_op_gemm_qkv(_lib.ptr(ah), _lib.ptr(al), c, _lib.ptr(wh), _lib.ptr(wl), c, M, c,
                                   _lib.ptr(norm_q.contiguous()), _lib.ptr(norm_k.contiguous()), rms_eps,
                                   LOG2E / math.sqrt(32.0), S_pad, *[_lib.ptr(o) for o in outs],
                                   _lib.stream_ptr(ah.device)), "gemm_qkv")
    return outs


def attention(q, k, v, bias):
    """q,k,v fp16 [B,H,S_pad,64] (rows [hi|lo]) + bias [H,S_pad,S_pad] (log2 domain) -> o planes [B*S_pad, H*32]."""
    lib = _lib.load()
    B, H, S_pad, _ = q.shape
    oh = torch.empty(B * S_pad, H * 32, dtype=torch.float16, device=q.device)
    ol = torch.empty_like(oh)
    _lib.check(lib.pdk_op_attention(_lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(bias.contiguous()),
                                    _lib.ptr(oh), _lib.ptr(ol), B, H, S_pad, _lib.stream_ptr(q.device)), "attention")
    return oh, ol


def precond(x_hat, coef, a, wx, bx, S_pad: int):
    lib = _lib.load()
    B, Na, _ = x_hat.shape
    c_a = a.shape[1]
    ba = torch.empty(B, S_pad, c_a, dtype=torch.float32, device=x_hat.device)
    _lib.check(lib.pdk_op_precond(_lib.ptr(x_hat.contiguous()), _lib.ptr(coef), _lib.ptr(a.contiguous()),
                                  _lib.ptr(wx.contiguous()), _lib.ptr(bx.contiguous()), _lib.ptr(ba), B, Na, S_pad, c_a,
                                  _lib.stream_ptr(x_hat.device)), "precond")
    return ba


def segment_mean(h, tok_start, s, St_pad: int):
    lib = _lib.load()
    B, Sa_pad, c_s = h.shape
    Nt = s.shape[0]
    bs = torch.empty(B, St_pad, c_s, dtype=torch.float32, device=h.device)
    _lib.check(lib.pdk_op_segment_mean(_lib.ptr(h.contiguous()), _lib.ptr(tok_start), _lib.ptr(s.contiguous()), _lib.ptr(bs),
                                       B, Nt, Sa_pad, St_pad, c_s, _lib.stream_ptr(h.device)), "segment_mean")
    return bs


def gather_add(ba, up, atom2tok, Na: int):
    lib = _lib.load()
    B, Sa_pad, c_a = ba.shape
    St_pad = up.shape[1]
    _lib.check(lib.pdk_op_gather_add(_lib.ptr(ba), _lib.ptr(up.contiguous()), _lib.ptr(atom2tok), B, Na, Sa_pad, St_pad,
                                     c_a, _lib.stream_ptr(ba.device)), "gather_add")
    return ba


def denoise_out(ba, x_hat, coef, ln_w, ln_b, wr, eps: float):
    lib = _lib.load()
    B, Na, _ = x_hat.shape
    S_pad, c_a = ba.shape[1], ba.shape[2]
    out = torch.empty_like(x_hat)
    _lib.check(lib.pdk_op_denoise_out(_lib.ptr(ba.contiguous()), _lib.ptr(x_hat.contiguous()), _lib.ptr(coef),
                                      _lib.ptr(ln_w.contiguous()), _lib.ptr(ln_b.contiguous()), _lib.ptr(wr.contiguous()),
                                      _lib.ptr(out), B, Na, S_pad, c_a, eps, _lib.stream_ptr(ba.device)), "denoise_out")
    return out


def precond_adaln(x_hat, coef, a, wx, bx, S_pad: int, mod, mod_off: int, eps: float):
    """The first AdaLN of the atom encoder with AF3DiT.precond fused in: returns (ba, hi, lo)."""
    lib = _lib.load()
    B, Na, _ = x_hat.shape
    c_a = a.shape[1]
    ba = torch.empty(B, S_pad, c_a, dtype=torch.float32, device=x_hat.device)
    hi = torch.empty(B, S_pad, c_a, dtype=torch.float16, device=x_hat.device)
    lo = torch.empty_like(hi)
    _lib.check(lib.pdk_op_precond_adaln(_lib.ptr(x_hat.contiguous()), _lib.ptr(coef), _lib.ptr(a.contiguous()),
                                        _lib.ptr(wx.contiguous()), _lib.ptr(bx.contiguous()), _lib.ptr(ba),
                                        _lib.ptr(mod.contiguous()), mod.shape[1], mod_off, _lib.ptr(hi), _lib.ptr(lo), B, Na,
                                        S_pad, c_a, eps, _lib.stream_ptr(x_hat.device)), "precond_adaln")
    return ba, hi, lo


def upscale_adaln(ba, up, atom2tok, Na: int, mod, mod_off: int, eps: float):
    """The first AdaLN of the atom decoder with the upscale gather-add fused in: ba is updated in place; returns (ba, hi, lo)."""
    lib = _lib.load()
    B, Sa_pad, c_a = ba.shape
    hi = torch.empty(B, Sa_pad, c_a, dtype=torch.float16, device=ba.device)
    lo = torch.empty_like(hi)
    _lib.check(lib.pdk_op_upscale_adaln(_lib.ptr(ba), _lib.ptr(up.contiguous()), _lib.ptr(atom2tok), _lib.ptr(mod.contiguous()),
                                        mod.shape[1], mod_off, _lib.ptr(hi), _lib.ptr(lo), B, Na, Sa_pad, up.shape[1], c_a, eps,
                                        _lib.stream_ptr(ba.device)), "upscale_adaln")
    return ba, hi, lo
