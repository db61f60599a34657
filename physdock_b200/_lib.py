"""ctypes binding of libphysdock_b200.so (include/physdock_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from .build import LIB_PATH

_c_float_p = C.c_void_p     # device pointers travel as integers
_vp = C.c_void_p
_i64 = C.c_int64
_f32 = C.c_float
_int = C.c_int


class PdkError(RuntimeError):
    pass


class DitDims(C.Structure):
    _fields_ = [(n, _i64) for n in ("c_a", "c_ap", "c_s", "c_z", "n_atom_blocks", "n_token_blocks",
                                    "hidden_a", "hidden_s", "n_mod")] + \
               [(n, C.c_double) for n in ("sigma_data", "eps", "inf")]


class BlockWeights(C.Structure):
    _fields_ = [(n, _vp) for n in ("wqkv_h", "wqkv_l", "wo_h", "wo_l", "w13_h", "w13_l", "w2_h", "w2_l",
                                   "bo", "norm_q", "norm_k")] + \
               [("mod_attn_off", _i64), ("mod_ffn_off", _i64)]


class DitWeights(C.Structure):
    _fields_ = [(n, _vp) for n in ("freq", "te_w1", "te_b1", "te_w2", "te_b2", "wmod_h", "wmod_l", "bmod", "wx", "bx",
                                   "wdown_h", "wdown_l", "bdown", "wup_h", "wup_l", "bup",
                                   "norm_r_w", "norm_r_b", "wr", "wz_atom_T", "bz_atom", "wz_tok_T", "bz_tok")] + \
               [("blocks", C.POINTER(BlockWeights)), ("n_blocks", _i64)]


# name -> (restype, argtypes); every symbol include/physdock_b200.h declares
PROTOTYPES = {
    "pdk_abi_version": (_int, []),
    "pdk_last_error": (C.c_char_p, []),
    "pdk_pad_len": (_i64, [_i64]),
    "pdk_dit_create": (_int, [C.POINTER(DitDims), C.POINTER(_vp)]),
    "pdk_dit_destroy": (_int, [_vp]),
    "pdk_dit_set_weights": (_int, [_vp, C.POINTER(DitWeights)]),
    "pdk_dit_bias_bytes": (_int, [_vp, _i64, _i64, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "pdk_dit_workspace_bytes": (_int, [_vp, _i64, _i64, _i64, C.POINTER(C.c_size_t)]),
    "pdk_dit_prepare_complex": (_int, [_vp] + [_vp] * 8 + [_i64, _i64, _vp, _vp, _vp]),
    "pdk_dit_denoise": (_int, [_vp, _vp, _vp, _i64, _vp, C.c_size_t, _vp, _vp]),
    "pdk_dit_launches_per_denoise": (_i64, [_vp]),
    "pdk_dit_cond_width": (_i64, [_vp]),
    "pdk_dit_conditioning_workspace_bytes": (_int, [_vp, _i64, C.POINTER(C.c_size_t)]),
    "pdk_dit_conditioning": (_int, [_vp, _vp, _i64, _vp, C.c_size_t, _vp, _i64, _vp]),
    "pdk_dit_denoise_cond": (_int, [_vp, _vp, _vp, _i64, _i64, _vp, C.c_size_t, _vp, _vp, _vp]),
    "pdk_dit_launches_per_denoise_cond": (_i64, [_vp]),
    "pdk_centre_augment": (_int, [_vp] * 5 + [_f32, _f32, _f32, _vp, _i64, _i64, _vp]),
    "pdk_euler_update": (_int, [_vp] * 5 + [_f32, _f32, _vp, _i64, _i64, _vp]),
    "pdk_template_select": (_int, [_vp] * 7 + [_i64] * 4 + [_vp]),
    "pdk_rigid_align": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _i64, _i64, _vp]),
    "pdk_attention_work_list": (_i64, [_i64, _i64, _i64, _i64, _vp, _i64]),
    "pdk_pairwise_rmsd": (_int, [_vp, _vp, _i64, _i64, _vp]),
    "pdk_pair_energy_grad": (_int, [_vp] * 7 + [_i64, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _i64, _i64,
                                    _vp]),
    "pdk_pair_descend": (_int, [_vp] * 7 + [_i64, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _i64, _f32, _f32, _vp, _i64, _i64, _vp]),
    "pdk_descent_update": (_int, [_vp, _vp, _vp, _f32, _f32, _vp, _i64, _i64, _vp]),
    "pdk_op_pair_bias": (_int, [_vp] * 5 + [_i64] * 4 + [_f32, _f32, _vp]),
    "pdk_op_time_embed": (_int, [_vp] * 6 + [_f32, _vp, _vp, _i64, _vp]),
    "pdk_op_mod_gemv": (_int, [_vp] * 4 + [_i64, _i64, _vp]),
    "pdk_op_adaln": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _f32, _vp]),
    "pdk_op_precond_adaln": (_int, [_vp] * 7 + [_i64, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp]),
    "pdk_op_upscale_adaln": (_int, [_vp] * 4 + [_i64, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _f32, _vp]),
    "pdk_op_split": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "pdk_op_gemm_store": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _int, _vp, _i64, _vp]),
    "pdk_op_gemm_gate_resid": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _i64,
                                      _vp, _i64, _vp]),
    "pdk_op_gemm_swiglu": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _vp]),
    "pdk_op_transition_fused": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _f32, _vp]),
    "pdk_op_gemm_qkv": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _f32, _f32, _i64] +
                        [_vp] * 3 + [_vp]),
    "pdk_op_attention": (_int, [_vp] * 6 + [_i64, _i64, _i64, _vp]),
    "pdk_op_precond": (_int, [_vp] * 6 + [_i64] * 4 + [_vp]),
    "pdk_op_segment_mean": (_int, [_vp] * 4 + [_i64] * 5 + [_vp]),
    "pdk_op_gather_add": (_int, [_vp] * 3 + [_i64] * 5 + [_vp]),
    "pdk_op_denoise_out": (_int, [_vp] * 7 + [_i64] * 4 + [_f32, _vp]),
}

ABI_VERSION = 2
_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """Loads the shared library (no CUDA call is made) and binds every prototype."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("PHYSDOCK_B200_LIB", LIB_PATH)
    if not os.path.exists(p) and p == LIB_PATH:
        try:                        # a fresh checkout: compile the CUDA library in-tree once (nvcc, ~30 s)
            from .build import build_library
            build_library()
        except Exception as e:      # noqa: BLE001 -- reported below; there is nothing to fall back to
            raise PdkError(f"{p} not found and building it failed ({e}); "
                           "there is no CPU or PyTorch fallback for the sampling step") from e
    if not os.path.exists(p):
        raise PdkError(f"{p} not found: build it with `python -m physdock_b200.build` "
                       "(there is no CPU or PyTorch fallback for the sampling step)")
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    if lib.pdk_abi_version() != ABI_VERSION:
        raise PdkError("ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().pdk_last_error()
        raise PdkError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t: Optional[torch.Tensor]):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise PdkError("physdock_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise PdkError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream
