"""ctypes binding of libhma_b200.so (the C ABI in include/hma_b200.h).

There is no fallback: if the shared library is missing or an entry point fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libhma_b200.so"

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float
c_fp = ctypes.c_void_p  # float* passed as raw address
c_u64 = ctypes.c_ulonglong


class HmaError(RuntimeError):
    pass


# name -> argtypes; every function returns int unless listed in _RESTYPES
_SIGNATURES = {
    "hma_abi_version": [],
    "hma_device_check": [],
    "hma_gemm_nt": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll,
                    c_fp, c_fp, c_ll, c_void_p, c_ll, c_float, c_fp, c_fp, c_void_p],
    "hma_gemm_nt_ln": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_ll, c_fp, c_fp, c_ll, c_float, c_int,
                       c_fp, c_fp, c_fp, c_int, c_float, c_void_p, c_ll, c_fp, c_void_p],
    "hma_mar_sampler": [c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_fp, c_fp, c_fp, c_void_p, c_ll, c_int,
                        c_void_p, c_fp, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_fp, c_fp,
                        c_void_p, c_void_p, c_void_p, c_fp, c_void_p, c_void_p],
    "hma_attn_spatial_fwd": [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_ll, c_fp,
                             c_void_p],
    "hma_attn_spatial_bwd": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_fp, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_float, c_void_p, c_ll, c_fp, c_void_p],
    "hma_attn_temporal_fwd": [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p,
                              c_ll, c_fp, c_void_p],
    "hma_attn_temporal_bwd": [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_fp, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_float, c_void_p, c_ll, c_void_p],
    "hma_kv_cache_append": [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_void_p],
    "hma_attn_temporal_cached": [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_float,
                                 c_void_p, c_ll, c_int, c_int, c_void_p],
    "hma_qk_norm_fwd": [c_void_p, c_ll, c_int, c_fp, c_fp, c_float, c_void_p, c_ll, c_void_p],
    "hma_qk_norm_bwd": [c_void_p, c_ll, c_int, c_fp, c_float, c_void_p, c_ll, c_fp, c_fp, c_void_p],
    "hma_ln_fwd": [c_fp, c_ll, c_int, c_int, c_fp, c_fp, c_fp, c_int, c_float, c_void_p, c_ll, c_fp, c_int, c_int,
                   c_void_p],
    "hma_group_add": [c_fp, c_fp, c_fp, c_int, c_int, c_void_p],
    "hma_group_colsum": [c_fp, c_fp, c_int, c_int, c_void_p],
    "hma_rows_scatter": [c_fp, c_fp, c_int, c_int, c_int, c_void_p],
    "hma_ln_bwd": [c_void_p, c_ll, c_fp, c_ll, c_fp, c_int, c_int, c_fp, c_fp, c_int, c_fp, c_ll, c_fp, c_fp, c_fp,
                   c_void_p, c_fp, c_void_p],
    "hma_colsum_bf16": [c_void_p, c_ll, c_int, c_int, c_fp, c_void_p],
    "hma_colsum_f32": [c_fp, c_ll, c_int, c_int, c_fp, c_void_p],
    "hma_cast_transpose": [c_fp, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p],
    "hma_cast_bf16": [c_fp, c_void_p, c_ll, c_void_p],
    "hma_cast_colsum": [c_fp, c_void_p, c_int, c_fp, c_void_p],
    "hma_cast_transpose_batched": [c_void_p, c_int, c_int, c_int, c_void_p],
    "hma_action_prep": [c_fp, c_int, c_int, c_fp, c_fp, c_int, c_void_p, c_int, c_void_p],
    "hma_ln_relu_fwd": [c_fp, c_int, c_fp, c_fp, c_float, c_void_p, c_fp, c_void_p],
    "hma_ln_relu_bwd": [c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_void_p],
    "hma_embed_fwd": [c_void_p, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_ll, c_fp,
                      c_void_p],
    "hma_embed_bwd": [c_void_p, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_ll, c_fp, c_fp, c_fp, c_fp, c_fp,
                      c_void_p],
    "hma_collate_maskgit": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_fp, c_float, c_void_p,
                            c_int, c_fp, c_fp, c_fp, c_fp, c_void_p],
    "hma_ce_fwd": [c_fp, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_float, c_fp, c_fp, c_fp,
                   c_void_p],
    "hma_ce_bwd": [c_fp, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_float, c_fp, c_fp, c_fp,
                   c_void_p, c_ll, c_void_p],
    "hma_sample_tokens": [c_fp, c_ll, c_ll, c_int, c_int, c_int, c_int, c_fp, c_float, c_void_p, c_fp, c_void_p],
    "hma_rank_remask": [c_fp, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_ll, c_void_p, c_void_p],
    "hma_sumsq": [c_fp, c_ll, c_fp, c_void_p],
    "hma_adamw_step": [c_fp, c_fp, c_fp, c_fp, c_ll, c_ll, c_float, c_float, c_float, c_float, c_float, c_int, c_float, c_fp,
                       c_float, c_void_p],
    "hma_gemm_wgrad": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_fp, c_ll, c_void_p],
    "hma_gemm_wgrad_grouped": [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p],
    "hma_mar_embed_fwd": [c_fp, c_void_p, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                          c_int, c_fp, c_fp, c_fp, c_void_p],
    "hma_mar_embed_bwd": [c_fp, c_fp, c_void_p, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_fp, c_fp, c_fp,
                          c_fp, c_void_p],
    "hma_mar_ln_fwd": [c_fp, c_int, c_int, c_fp, c_fp, c_float, c_void_p, c_ll, c_int, c_int, c_fp, c_int, c_fp, c_void_p, c_fp,
                       c_void_p, c_ll, c_int, c_void_p, c_fp, c_void_p],
    "hma_mar_ln_bwd": [c_void_p, c_fp, c_fp, c_fp, c_int, c_int, c_fp, c_fp, c_void_p, c_ll, c_int, c_int, c_fp, c_int, c_void_p,
                       c_fp, c_fp, c_void_p, c_ll, c_fp, c_int, c_void_p],
    "hma_mar_gate_fwd": [c_fp, c_void_p, c_ll, c_int, c_void_p, c_int, c_int, c_fp, c_void_p],
    "hma_mar_gate_bwd": [c_fp, c_void_p, c_ll, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_ll, c_void_p],
    "hma_mar_silu_fwd": [c_fp, c_fp, c_ll, c_int, c_void_p, c_void_p],
    "hma_mar_silu_steps": [c_fp, c_fp, c_ll, c_int, c_int, c_void_p, c_void_p],
    "hma_mar_silu_bwd": [c_fp, c_fp, c_ll, c_void_p, c_void_p],
    "hma_mar_q_sample": [c_fp, c_fp, c_void_p, c_fp, c_ll, c_int, c_int, c_void_p, c_void_p],
    "hma_mar_timestep_embed": [c_void_p, c_ll, c_void_p, c_void_p],
    "hma_mar_diff_loss_fwd": [c_fp, c_ll, c_fp, c_fp, c_void_p, c_fp, c_fp, c_ll, c_int, c_fp, c_fp, c_fp, c_void_p],
    "hma_mar_diff_loss_bwd": [c_fp, c_ll, c_fp, c_fp, c_void_p, c_fp, c_fp, c_ll, c_int, c_fp, c_fp, c_void_p, c_ll, c_void_p],
    "hma_mar_p_sample": [c_fp, c_ll, c_fp, c_fp, c_fp, c_int, c_ll, c_int, c_float, c_int, c_fp, c_void_p, c_int, c_void_p],
    "hma_mar_gather_rows": [c_fp, c_void_p, c_ll, c_int, c_fp, c_void_p, c_void_p],
    "hma_mar_scatter_rows": [c_fp, c_void_p, c_ll, c_int, c_fp, c_void_p],
    "hma_gather_token_windows": [c_void_p, c_int, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "hma_gather_rows_f32": [c_fp, c_ll, c_ll, c_void_p, c_int, c_ll, c_fp, c_void_p],
    "hma_conv3x3_nhwc": [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_fp, c_fp, c_ll, c_void_p],
    "hma_lfq_entry": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "hma_gn_stats": [c_fp, c_int, c_int, c_int, c_int, c_fp, c_fp, c_void_p],
    "hma_gn_swish": [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p],
    "hma_depth_to_space": [c_fp, c_int, c_int, c_int, c_int, c_fp, c_void_p],
    "hma_to_uint8": [c_fp, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "hma_dropout_bf16": [c_void_p, c_ll, c_float, c_u64, c_void_p, c_void_p],
    "hma_dropout_add_f32": [c_fp, c_fp, c_fp, c_ll, c_float, c_u64, c_void_p, c_void_p],
    "hma_dropout_cast_bf16": [c_fp, c_void_p, c_ll, c_float, c_u64, c_void_p, c_void_p],
}
_RESTYPES = {"hma_last_error": ctypes.c_char_p}


def register(name: str, argtypes: list) -> None:
    _SIGNATURES[name] = argtypes


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            if os.environ.get("HMA_B200_AUTOBUILD", "1") == "1":
                from . import build as _build

                _build.build()
            if not LIB_PATH.exists():
                raise HmaError(
                    f"{LIB_PATH} not found: build it with `python -m hma_b200.build` (no CPU fallback exists)"
                )
        L = ctypes.CDLL(str(LIB_PATH))
        L.hma_last_error.restype = ctypes.c_char_p
        L.hma_last_error.argtypes = []
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing: loud by design
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = L
    return _lib


def call(name: str, *args) -> None:
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        msg = L.hma_last_error()
        raise HmaError(f"{name} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t) -> int | None:
    """Device address of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
