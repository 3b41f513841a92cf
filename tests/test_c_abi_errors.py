"""Error behaviour of the C ABI (include/hma_b200.h): invalid arguments are rejected before anything is launched, with a
negative return code and a message from hma_last_error(); no exception crosses the boundary. Runs without a GPU — every
call below fails its argument validation first."""
import ctypes

import pytest

from hma_b200 import _lib


def _call(name, *args):
    L = _lib.lib()
    rc = getattr(L, name)(*args)
    msg = L.hma_last_error()
    return rc, (msg.decode() if msg else "")


P = 0x1000  # a non-null dummy address: validation must fail before it is ever dereferenced


def test_gemm_rejects_bad_shapes():
    rc, msg = _call("hma_gemm_nt", P, 256, P, 256, 128, 256, 100, 0, P, 256, None, 0, None, None, 0, None, 0, 1.0, None, None, None)
    assert rc < 0 and "multiple of 64" in msg
    rc, msg = _call("hma_gemm_nt", P, 256, P, 256, 128, 200, 256, 0, P, 200, None, 0, None, None, 0, None, 0, 1.0, None, None, None)
    assert rc < 0 and "multiple of 128" in msg
    rc, msg = _call("hma_gemm_nt", P, 256, P, 256, 128, 256, 256, 0, None, 256, None, 0, None, None, 0, None, 0, 1.0, None, None, None)
    assert rc < 0 and "out is null" in msg
    rc, msg = _call("hma_gemm_nt", P, 256, P, 256, 0, 256, 256, 0, P, 256, None, 0, None, None, 0, None, 0, 1.0, None, None, None)
    assert rc == 0  # an empty problem is a no-op
    rc, msg = _call("hma_gemm_wgrad", P, 256, P, 256, 1000, 100, 256, P, 256, None)
    assert rc < 0 and "multiple of 128" in msg


def test_mar_entry_points_reject_unsupported_configurations():
    rc, msg = _call("hma_mar_ln_fwd", P, 8, 512, None, None, 1e-6, None, 0, 0, 0, None, 0, P, None, None, None, 0, 0, None, None, None)
    assert rc < 0 and "256 or 1024" in msg
    rc, msg = _call("hma_mar_ln_fwd", P, 8, 256, P, None, 1e-6, None, 0, 0, 0, None, 0, P, None, None, None, 0, 0, None, None, None)
    assert rc < 0 and "gamma and beta" in msg
    rc, msg = _call("hma_mar_ln_bwd", None, None, P, P, 8, 256, None, None, None, 0, 0, 0, P, 0, None, None, None, None, 0, None, 0, None)
    assert rc < 0 and "exactly one of" in msg
    rc, msg = _call("hma_mar_embed_fwd", P, None, None, None, P, None, P, 320, 2, 4, 16, 16, 4, 3, 0, 0, P, None, None, None)
    assert rc < 0 and "multiples of the patch size" in msg
    rc, msg = _call("hma_mar_embed_fwd", P, None, None, None, P, None, P, 320, 2, 4, 16, 16, 32, 2, 0, 0, P, None, None, None)
    assert rc < 0 and "not supported" in msg
    rc, msg = _call("hma_mar_embed_fwd", P, None, None, None, P, None, P, 320, 2, 4, 16, 16, 4, 2, 64, 0, P, None, None, None)
    assert rc < 0 and "action" in msg
    rc, msg = _call("hma_mar_p_sample", P, 8, P, None, P, 5, 16, 16, 1.0, 1, P, None, 16, None)
    assert rc < 0  # ldo < 2 * D
    rc, msg = _call("hma_mar_p_sample", P, 128, P, None, P, 5, 16, 16, 1.0, 1, P, None, 16, None)
    assert rc < 0 and "noise is required" in msg
    rc, msg = _call("hma_dropout_bf16", P, 1024, 1.0, 1, None, None)
    assert rc < 0 and "out of range" in msg
    rc, msg = _call("hma_dropout_bf16", P + 2, 1024, 0.1, 1, None, None)
    assert rc < 0 and "aligned" in msg
    rc, msg = _call("hma_gather_token_windows", P, 8, 100, P, 4, 4, 1, 256, P, None)
    assert rc < 0 and "uint16 or uint32" in msg


def test_row_and_attention_entry_points_reject_bad_arguments():
    rc, msg = _call("hma_colsum_bf16", P, 4096, 128, 4096, P, None)
    assert rc < 0 and "<= 2048" in msg
    rc, msg = _call("hma_adamw_step", P, P, P, P, 1024, 1024, 1e-4, 0.9, 0.999, 1e-8, 0.0, 0, 1.0, None, 1.0, None)
    assert rc < 0 and "step counts from 1" in msg
    rc, msg = _call("hma_sumsq", P + 4, 1024, P, None)
    assert rc < 0 and "aligned" in msg
    rc, msg = _call("hma_action_prep", P, 4, 100, None, None, 0, P, 64, None)
    assert rc < 0 and "kpad" in msg


def test_python_binding_raises_with_the_library_message():
    with pytest.raises(_lib.HmaError, match="multiple of 64"):
        _lib.call("hma_gemm_nt", P, 256, P, 256, 128, 256, 100, 0, P, 256, None, 0, None, None, 0, None, 0, 1.0, None, None, None)
    assert isinstance(_lib.lib(), ctypes.CDLL)


def test_round2_entry_points_reject_bad_arguments():
    # residual GEMM + LayerNorm epilogue: whole rows only (N == 256), a valid mode, its operands
    ln = lambda N, mode, gamma, out: _call("hma_gemm_nt_ln", P, 256, P, 256, 128, N, 256, P, N, None, None, 0, 1.0, mode, gamma, gamma,  # noqa: E731
                                           None, 0, 1e-5, out, 256, None, None)
    rc, msg = ln(512, 1, P, P)
    assert rc < 0 and "N == 256" in msg
    rc, msg = ln(256, 3, P, P)
    assert rc < 0 and "mode" in msg
    rc, msg = ln(256, 1, None, P)
    assert rc < 0 and "gamma/beta" in msg
    rc, msg = ln(256, 2, None, P)
    assert rc < 0 and "shift/scale" in msg
    # cached temporal attention with several frames per sample: rows must be whole samples
    rc, msg = _call("hma_attn_temporal_cached", P, 768, 0, 256, 512, P, 4096, 100, 3, 8, 0.17, P, 256, 2, 32, None)
    assert rc < 0 and "whole number" in msg
    # persistent sampler: token dimension, depth, modulation range
    vp = (ctypes.c_void_p * 4)(P, P, P, P)
    arr = ctypes.cast(vp, ctypes.c_void_p)
    smp = lambda D, depth, lo, m0: _call("hma_mar_sampler", 64, D, depth, 10, lo, 1.0, 1, P, P, P, P, 14336, m0, P, P, arr, arr, arr, arr,  # noqa: E731
                                         arr, arr, P, P, P, P, P, P, None, P, None)
    rc, msg = smp(40, 4, 0, 0)
    assert rc < 0 and "token dimension" in msg
    rc, msg = smp(16, 9, 0, 0)
    assert rc < 0 and "depth" in msg
    rc, msg = smp(16, 4, 2, 5)
    assert rc < 0 and "modulations start" in msg
    rc, msg = _call("hma_mar_sampler", 0, 16, 4, 10, 0, 1.0, 1, P, P, P, P, 14336, 0, P, P, arr, arr, arr, arr, arr, arr, P, P, P, P, P, P,
                    None, P, None)
    assert rc == 0  # no rows: a no-op
