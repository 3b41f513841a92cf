"""tcgen05 contractions vs fp32 torch matmul on the same bf16-rounded inputs (GPU only)."""
import pytest
import torch

from hma_b200 import _lib

pytestmark = pytest.mark.gpu

EPI_BF16, EPI_GELU, EPI_DGELU, EPI_RESID = 0, 1, 2, 3


def _report(name, got, ref):
    err = (got.float() - ref.float()).abs()
    denom = ref.float().abs().max().clamp_min(1e-6)
    rel = (err.max() / denom).item()
    if rel > 2e-2:
        # error structure by 32-row / 32-col blocks helps to localise descriptor / swizzle bugs
        M, N = err.shape
        rb = err[: M // 32 * 32].reshape(M // 32, 32, N).amax(dim=(1, 2))
        cb = err[:, : N // 32 * 32].reshape(M, N // 32, 32).amax(dim=(0, 2))
        print(f"[{name}] rel={rel:.3e} rowblocks={rb[:16].tolist()} colblocks={cb[:16].tolist()}")
    return rel


def gemm_nt(A, B, epi, bias=None, resid=None, aux=None, alpha=1.0, want_z=False, colsum=None, rowdot=None):
    M, K = A.shape
    N = B.shape[0]
    out_dtype = torch.float32 if epi == EPI_RESID else torch.bfloat16
    out = torch.empty(M, N, device=A.device, dtype=out_dtype)
    out2 = torch.empty(M, N, device=A.device, dtype=torch.bfloat16) if want_z else None
    _lib.call(
        "hma_gemm_nt", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K, epi, out.data_ptr(),
        out.stride(0), _lib.ptr(out2), N, _lib.ptr(bias), _lib.ptr(resid), N, _lib.ptr(aux), N, alpha,
        _lib.ptr(colsum), _lib.ptr(rowdot), _lib.current_stream(),
    )
    return out, out2


def test_gemm_nt_rowdot_is_the_attention_delta():
    """EPI_BF16 with rowdot: per (row, 32-column chunk) sum of bf16(out) * aux — delta = rowsum(dO * O) per (token, head)."""
    torch.manual_seed(3)
    M, N, K = 1300, 256, 256
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    O = torch.randn(M, N, device="cuda").bfloat16()
    rowdot = torch.full((M, N // 32), float("nan"), device="cuda")
    out, _ = gemm_nt(A, B, EPI_BF16, aux=O, rowdot=rowdot)
    plain, _ = gemm_nt(A, B, EPI_BF16)
    assert torch.equal(out, plain)
    want = (out.float() * O.float()).view(M, N // 32, 32).sum(-1)
    assert torch.allclose(rowdot, want, rtol=1e-4, atol=1e-4), (rowdot - want).abs().max().item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 256), (1000, 768, 256), (4096, 1024, 256),
                                   (640, 256, 1024), (333, 256, 768), (40960, 256, 256)])
def test_gemm_nt_bf16(M, N, K):
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda")
    out, _ = gemm_nt(A, B, EPI_BF16, bias=bias)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t() + bias
    assert _report(f"nt {M}x{N}x{K}", out, ref) < 1e-2


def test_gemm_nt_epilogues():
    torch.manual_seed(1)
    M, N, K = 777, 1024, 256
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.08).bfloat16()
    bias = torch.randn(N, device="cuda") * 0.5
    z_ref = A.float() @ B.float().t() + bias
    h, z = gemm_nt(A, B, EPI_GELU, bias=bias, want_z=True)
    torch.cuda.synchronize()
    assert _report("gelu.z", z, z_ref) < 1e-2
    assert _report("gelu.h", h, torch.nn.functional.gelu(z_ref)) < 1e-2

    # dGELU: out = acc * gelu'(aux)
    aux = torch.randn(M, N, device="cuda").bfloat16()
    g, _ = gemm_nt(A, B, EPI_DGELU, aux=aux)
    torch.cuda.synchronize()
    zz = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    assert _report("dgelu", g, (A.float() @ B.float().t()) * zz.grad) < 1e-2
    # fused bias gradient: colsum += column sums of the d-activation output (accumulates into the caller's buffer)
    for Mc in (777, 5000):
        Ac = torch.randn(Mc, K, device="cuda").bfloat16()
        auxc = torch.randn(Mc, N, device="cuda").bfloat16()
        cs = torch.full((N,), 3.0, device="cuda")
        gc, _ = gemm_nt(Ac, B, EPI_DGELU, aux=auxc, colsum=cs)
        torch.cuda.synchronize()
        ref_cs = 3.0 + gc.float().sum(0)
        assert ((cs - ref_cs).abs().max() / ref_cs.abs().max()).item() < 5e-3

    # fp32 residual
    resid = torch.randn(M, N, device="cuda")
    o, _ = gemm_nt(A, B, EPI_RESID, bias=bias, resid=resid, alpha=0.5)
    torch.cuda.synchronize()
    ref = resid + 0.5 * (A.float() @ B.float().t()) + bias
    assert _report("resid", o, ref) < 2e-3
    o2, _ = gemm_nt(A, B, EPI_RESID)
    torch.cuda.synchronize()
    assert _report("f32", o2, A.float() @ B.float().t()) < 2e-3


@pytest.mark.parametrize("tokens,Mw,Nw", [(64, 128, 128), (4096, 256, 256), (5000, 768, 256), (3000, 256, 1024),
                                          (40960, 1024, 256)])
def test_gemm_wgrad(tokens, Mw, Nw):
    torch.manual_seed(2)
    G = (torch.randn(tokens, Mw, device="cuda") * 0.1).bfloat16()
    X = torch.randn(tokens, Nw, device="cuda").bfloat16()
    dW = torch.zeros(Mw, Nw, device="cuda")
    for _ in range(2):  # accumulates
        _lib.call("hma_gemm_wgrad", G.data_ptr(), Mw, X.data_ptr(), Nw, tokens, Mw, Nw, dW.data_ptr(), Nw,
                  _lib.current_stream())
    torch.cuda.synchronize()
    ref = 2 * (G.float().t() @ X.float())
    assert _report(f"wgrad {tokens}x{Mw}x{Nw}", dW, ref) < 2e-3


def test_gemm_wgrad_strided_views():
    """Operands that are column slices of wider matrices (as qkv / grad buffers are)."""
    torch.manual_seed(3)
    tokens = 2000
    Gfull = (torch.randn(tokens, 768, device="cuda") * 0.1).bfloat16()
    X = torch.randn(tokens, 256, device="cuda").bfloat16()
    G = Gfull[:, 256:512]
    dW = torch.zeros(256, 256, device="cuda")
    _lib.call("hma_gemm_wgrad", G.data_ptr(), 768, X.data_ptr(), 256, tokens, 256, 256, dW.data_ptr(), 256,
              _lib.current_stream())
    torch.cuda.synchronize()
    assert _report("wgrad-view", dW, G.float().t() @ X.float()) < 2e-3


@pytest.mark.parametrize("tokens", [2560, 40960, 1000])
def test_gemm_wgrad_grouped_matches_fp32_reference(tokens):
    """The seven weight gradients of an ST block in one persistent launch (+ one shape the grouped kernel hands to the
    single-GEMM kernel): dW_j += G_j^T X_j against an fp32 matmul of the same bf16 operands; accumulation into a non-zero
    dW; ragged token count (zero-filled TMA tail)."""
    from hma_b200 import ops
    torch.manual_seed(tokens)
    shapes = [(256, 1024), (1024, 256), (256, 256), (768, 256), (256, 256), (256, 256), (768, 256), (256, 128)]
    group, refs = [], []
    for j, (mw, nw) in enumerate(shapes):
        G = (torch.randn(tokens, mw, device="cuda") * 0.5).bfloat16()
        X = torch.randn(tokens, nw, device="cuda").bfloat16()
        dW = torch.full((mw, nw), float(j), device="cuda")
        group.append((G, X, dW))
        refs.append(float(j) + G.float().t() @ X.float())
    ops.gemm_wgrad_grouped(group)
    torch.cuda.synchronize()
    for (G, X, dW), ref, shp in zip(group, refs, shapes):
        err = (dW - ref).abs().max().item()
        assert err <= 2e-3 * ref.abs().max().item() + 1e-3, (shp, err, ref.abs().max().item())


@pytest.mark.parametrize("M,K,mode", [(1300, 256, 1), (1300, 256, 2), (41, 1024, 1), (640, 1024, 2), (40960, 256, 1)])
def test_gemm_nt_ln_epilogue_emits_the_next_prenorm(M, K, mode):
    """hma_gemm_nt_ln: the fp32 residual output is bit-identical to the plain residual epilogue, and the bf16 LayerNorm it
    emits (affine / modulated per row group) plus the (mean, rstd) rows match fp32 torch on that output."""
    torch.manual_seed(5)
    N, rpg = 256, 320
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda") * 3 + 0.5
    gamma, beta = torch.randn(N, device="cuda"), torch.randn(N, device="cuda")
    groups = (M + rpg - 1) // rpg
    mod = torch.randn(groups, 2 * N, device="cuda") * 0.3
    eps = 1e-5 if mode == 1 else 1e-6
    x = torch.empty(M, N, device="cuda")
    y = torch.full((M, N), float("nan"), device="cuda").bfloat16()
    stats = torch.full((M, 2), float("nan"), device="cuda")
    _lib.call("hma_gemm_nt_ln", A.data_ptr(), K, B.data_ptr(), K, M, N, K, x.data_ptr(), N, _lib.ptr(bias), _lib.ptr(resid), N, 1.0,
              mode, _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(mod), rpg, eps, y.data_ptr(), N, _lib.ptr(stats), _lib.current_stream())
    plain, _ = gemm_nt(A, B, EPI_RESID, bias=bias, resid=resid)
    assert torch.equal(x, plain)
    mean = x.mean(-1, keepdim=True)
    rstd = (x.var(-1, unbiased=False, keepdim=True) + eps).rsqrt()
    assert torch.allclose(stats[:, :1], mean, rtol=1e-5, atol=1e-5)
    assert torch.allclose(stats[:, 1:], rstd, rtol=1e-4, atol=1e-6)
    xh = (x - mean) * rstd
    if mode == 1:
        want = xh * gamma + beta
    else:
        g = torch.arange(M, device="cuda") // rpg
        want = xh * (1 + mod[g, N:]) + mod[g, :N]
    err = (y.float() - want).abs().max().item()
    assert err <= 2e-2 * want.abs().max().item() and torch.isfinite(y.float()).all(), err
    # and the separate kernel it replaces rounds to the same bf16 values up to the last place
    from hma_b200 import ops
    ref = ops.ln_fwd(x, mode, gamma=gamma, beta=beta, mod=mod, rows_per_group=rpg, eps=eps)
    assert (y.float() - ref.float()).abs().max().item() <= 4e-2 * want.abs().max().item() / 4


def test_gemm_nt_ln_in_place_residual():
    """Inference writes the stream in place (out == resid)."""
    torch.manual_seed(6)
    M, K, N = 700, 256, 256
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    x0 = torch.randn(M, N, device="cuda")
    gamma, beta = torch.randn(N, device="cuda"), torch.randn(N, device="cuda")
    from hma_b200 import ops
    x_ref, y_ref, _ = ops.gemm_nt_ln(A, B, resid=x0, ln_mode=1, gamma=gamma, beta=beta)
    x1 = x0.clone()
    x_out, y, _ = ops.gemm_nt_ln(A, B, resid=x1, out=x1, ln_mode=1, gamma=gamma, beta=beta)
    assert x_out.data_ptr() == x1.data_ptr() and torch.equal(x1, x_ref) and torch.equal(y, y_ref)
