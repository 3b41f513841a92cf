"""Pins the shared-memory descriptor variants used by the attention kernels on the real GPU
(K-major / MN-major, 128B / 64B swizzle, sub-atom start offsets) with the single-CTA probe."""
import ctypes

import pytest
import torch

from hma_b200 import _lib, build as _build

pytestmark = pytest.mark.gpu

_PROBE = None


def _probe_lib():
    """The probe is test infrastructure with its own shared library (tests/libhma_b200_probe.so, built by
    hma_b200.build.build_probe); it is not an entry point of the product ABI."""
    global _PROBE
    if _PROBE is None:
        _PROBE = ctypes.CDLL(str(_build.build_probe()))
        _PROBE.hma_umma_probe.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]
        _PROBE.hma_umma_probe.restype = ctypes.c_int
    return _PROBE


def probe(A, B, *, box_inner, layout_type, a_rows, a_boxes, b_rows, b_boxes, a_major, b_major, a_off=0, b_off=0,
          a_lbo=16, a_sbo=1024, b_lbo=16, b_sbo=1024, ksteps=4, a_kstep=32, b_kstep=32, N=64):
    params = (ctypes.c_int * 18)(a_rows, a_boxes, b_rows, b_boxes, box_inner, layout_type, a_major, b_major, a_off,
                                 b_off, a_lbo, a_sbo, b_lbo, b_sbo, ksteps, a_kstep, b_kstep, N)
    out = torch.zeros(128, N, device="cuda")
    rc = _probe_lib().hma_umma_probe(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
                                     ctypes.cast(params, ctypes.c_void_p), out.data_ptr(), _lib.current_stream())
    assert rc == 0, rc
    torch.cuda.synchronize()
    return out


def rnd(r, c, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(r, c, device="cuda", generator=g).bfloat16()


def check(out, ref, name):
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"[probe] {name}: max_err={err:.4g} ref_max={scale:.4g}")
    assert err <= 2e-2 * scale + 1e-3, name


def test_sw128_kmajor_baseline():
    A, B = rnd(128, 64, 1), rnd(64, 64, 2)
    out = probe(A, B, box_inner=64, layout_type=2, a_rows=128, a_boxes=1, b_rows=64, b_boxes=1, a_major=0, b_major=0)
    check(out, A.float() @ B.float().t(), "sw128 K-major")


def test_sw128_kmajor_subatom_offset():
    A, B = rnd(128, 64, 3), rnd(64, 64, 4)
    out = probe(A, B, box_inner=64, layout_type=2, a_rows=128, a_boxes=1, b_rows=64, b_boxes=1, a_major=0, b_major=0,
                a_off=64, b_off=64, ksteps=2)
    check(out, A[:, 32:].float() @ B[:, 32:].float().t(), "sw128 K-major +64B")


def test_sw128_mnmajor_b():
    A, B = rnd(128, 64, 5), rnd(64, 64, 6)  # B: [k=64, n=64]
    out = probe(A, B, box_inner=64, layout_type=2, a_rows=128, a_boxes=1, b_rows=64, b_boxes=1, a_major=0, b_major=1,
                b_lbo=8192, b_sbo=1024, b_kstep=2048, N=64)
    check(out, A.float() @ B.float(), "sw128 MN-major B")


@pytest.mark.parametrize("off", [0, 64])
def test_sw128_mnmajor_b_n32_offset(off):
    A, B = rnd(128, 64, 7), rnd(64, 64, 8)
    out = probe(A, B, box_inner=64, layout_type=2, a_rows=128, a_boxes=1, b_rows=64, b_boxes=1, a_major=0, b_major=1,
                b_off=off, b_lbo=8192, b_sbo=1024, b_kstep=2048, N=32)
    c0 = off // 2
    check(out, A.float() @ B[:, c0:c0 + 32].float(), f"sw128 MN-major B N=32 off={off}")


def test_sw128_mnmajor_a():
    A, B = rnd(64, 128, 9), rnd(64, 64, 10)  # A: [k=64, m=128]
    out = probe(A, B, box_inner=64, layout_type=2, a_rows=64, a_boxes=2, b_rows=64, b_boxes=1, a_major=1, b_major=0,
                a_lbo=8192, a_sbo=1024, a_kstep=2048)
    check(out, A.float().t() @ B.float().t(), "sw128 MN-major A")


def test_sw64_kmajor():
    A, B = rnd(128, 32, 11), rnd(64, 32, 12)
    out = probe(A, B, box_inner=32, layout_type=4, a_rows=128, a_boxes=1, b_rows=64, b_boxes=1, a_major=0, b_major=0,
                a_sbo=512, b_sbo=512, ksteps=2)
    check(out, A.float() @ B.float().t(), "sw64 K-major")


def test_sw64_kmajor_b256():
    A, B = rnd(128, 32, 13), rnd(256, 32, 14)
    out = probe(A, B, box_inner=32, layout_type=4, a_rows=128, a_boxes=1, b_rows=256, b_boxes=1, a_major=0, b_major=0,
                a_sbo=512, b_sbo=512, ksteps=2, N=256)
    check(out, A.float() @ B.float().t(), "sw64 K-major N=256")


def test_sw64_mnmajor_b():
    A, B = rnd(128, 32, 15), rnd(32, 32, 16)  # B: [k=32, n=32]
    out = probe(A, B, box_inner=32, layout_type=4, a_rows=128, a_boxes=1, b_rows=32, b_boxes=1, a_major=0, b_major=1,
                a_sbo=512, b_lbo=2048, b_sbo=512, ksteps=2, b_kstep=1024, N=32)
    check(out, A.float() @ B.float(), "sw64 MN-major B")


def test_sw64_mnmajor_b_longk():
    """P.V shape: A = P [128 x 256 keys] K-major SW128 is a different swizzle, so here A is SW64 with
    K = 32 only; the long-K walk of the MN-major operand is checked with 8 k-steps over 128 key rows
    against a K-major SW64 A that is re-used (a_kstep = 0) for every step."""
    A, B = rnd(128, 32, 17), rnd(128, 32, 18)  # B: [k=128, n=32]
    out = probe(A, B, box_inner=32, layout_type=4, a_rows=128, a_boxes=1, b_rows=128, b_boxes=1, a_major=0, b_major=1,
                a_sbo=512, b_lbo=2048, b_sbo=512, ksteps=8, a_kstep=0, b_kstep=1024, N=32)
    Af = A.float()[:, :16]
    ref = sum(Af @ B.float()[16 * k:16 * k + 16] for k in range(8))
    check(out, ref, "sw64 MN-major B long K")


def test_sw64_mnmajor_a():
    A, B = rnd(32, 128, 19), rnd(64, 32, 20)  # A: [k=32, m=128] -> 4 boxes of [32 x 32]
    out = probe(A, B, box_inner=32, layout_type=4, a_rows=32, a_boxes=4, b_rows=64, b_boxes=1, a_major=1, b_major=0,
                a_lbo=2048, a_sbo=512, a_kstep=1024, b_sbo=512, ksteps=2)
    check(out, A.float().t() @ B.float().t(), "sw64 MN-major A")
