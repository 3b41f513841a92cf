"""Thin typed wrappers over the C ABI (include/hma_b200.h). Torch is used only to own device memory
and to supply the current stream; every computation below happens inside libhma_b200.so. There is
no fallback path: a missing library or a failing entry point raises.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

EPI_BF16, EPI_GELU, EPI_DGELU, EPI_RESID, EPI_SILU, EPI_DSILU = 0, 1, 2, 3, 4, 5
BF16 = torch.bfloat16
F32 = torch.float32


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


class Profiler:
    """CUDA-event timing of individual stage launches on the launching stream (bench.py / tools)."""

    def __init__(self, kinds=None):
        self.kinds = kinds  # None = every stage; else a set of kind prefixes
        self.records = {}

    def want(self, kind: str) -> bool:
        return self.kinds is None or any(kind.startswith(k) for k in self.kinds)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for kind, recs in self.records.items():
            ms = [a.elapsed_time(b) for a, b, _ in recs]
            srt = sorted(ms)
            # robust total: median x launches (an event pair also times whatever else delays its kernel when the host, not
            # the GPU, is the bottleneck of an eagerly launched step; a few such outliers must not define a stage's share)
            out[kind] = {"launches": len(ms), "total_ms": sum(ms), "robust_ms": srt[len(srt) // 2] * len(ms),
                         "max_ms": srt[-1], "work": sum(w for _, _, w in recs)}
        return out


PROFILER: Optional[Profiler] = None
LAUNCHES = 0


def _call(kind: str, work: float, name: str, *args) -> None:
    """One C-ABI launch; `work` = algorithmic FLOPs (tensor kernels) or bytes (HBM-bound kernels)."""
    global LAUNCHES
    LAUNCHES += 1
    prof = PROFILER
    if prof is not None and prof.want(kind):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call(name, *args)
        e1.record()
        prof.records.setdefault(kind, []).append((e0, e1, work))
    else:
        _lib.call(name, *args)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def gemm_nt(A: torch.Tensor, B: torch.Tensor, epi: int, *, out: Optional[torch.Tensor] = None,
            bias: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
            aux: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None, alpha: float = 1.0,
            colsum: Optional[torch.Tensor] = None, rowdot: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = epi(A[M,K] @ B[N,K]^T); A, B bf16 (row stride arbitrary, unit column stride)."""
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K and A.dtype == BF16 and B.dtype == BF16 and A.stride(1) == 1 and B.stride(1) == 1
    if out is None:
        out = torch.empty(M, N, device=A.device, dtype=F32 if epi == EPI_RESID else BF16)
    _call(f"gemm_nt[N={N},K={K},epi={epi}]", 2.0 * M * N * K, "hma_gemm_nt", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, K, epi, out.data_ptr(),
              out.stride(0), _p(out2), out2.stride(0) if out2 is not None else 0, _p(bias), _p(resid),
              resid.stride(0) if resid is not None else 0, _p(aux), aux.stride(0) if aux is not None else 0,
              float(alpha), _p(colsum), _p(rowdot), _s())
    return out


def gemm_nt_ln(A: torch.Tensor, B: torch.Tensor, *, resid: Optional[torch.Tensor], ln_mode: int, out: Optional[torch.Tensor] = None,
               bias: Optional[torch.Tensor] = None, alpha: float = 1.0, gamma=None, beta=None, mod=None, rows_per_group: int = 0,
               eps: float = 1e-5, want_stats: bool = False):
    """x = resid + A[M,K] @ B[256,K]^T + bias (fp32) and, from the same epilogue, y = bf16(LayerNorm(x)) for the next stage
    (ln_mode 1: affine; 2: (1 + scale) * LN(x) + shift per group of rows_per_group rows). Returns (x, y, stats or None)."""
    M, K = A.shape
    N = B.shape[0]
    assert N == 256 and B.shape[1] == K and A.dtype == BF16 and B.dtype == BF16 and A.stride(1) == 1 and B.stride(1) == 1
    if out is None:
        out = torch.empty(M, N, device=A.device, dtype=F32)
    y = torch.empty(M, N, device=A.device, dtype=BF16)
    stats = torch.empty(M, 2, device=A.device, dtype=F32) if want_stats else None
    _call(f"gemm_nt_ln[K={K},mode={ln_mode}]", 2.0 * M * N * K, "hma_gemm_nt_ln", A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0),
          M, N, K, out.data_ptr(), out.stride(0), _p(bias), _p(resid), resid.stride(0) if resid is not None else 0, float(alpha),
          ln_mode, _p(gamma), _p(beta), _p(mod), rows_per_group, float(eps), y.data_ptr(), y.stride(0), _p(stats), _s())
    return out, y, stats


def gemm_wgrad(G: torch.Tensor, X: torch.Tensor, dW: torch.Tensor) -> None:
    """dW[Mw,Nw] (fp32) += G[tokens,Mw]^T @ X[tokens,Nw]."""
    tokens, Mw = G.shape
    Nw = X.shape[1]
    assert X.shape[0] == tokens and dW.shape == (Mw, Nw) and dW.dtype == F32 and dW.stride(1) == 1
    _call(f"gemm_wgrad[{Mw}x{Nw}]", 2.0 * tokens * Mw * Nw, "hma_gemm_wgrad", G.data_ptr(), G.stride(0), X.data_ptr(), X.stride(0), tokens, Mw, Nw, dW.data_ptr(),
              dW.stride(0), _s())


def gemm_wgrad_grouped(group) -> None:
    """group: list of (G [tokens, Mw] bf16, X [tokens, Nw] bf16, dW [Mw, Nw] fp32), all over the same tokens:
    dW += G^T @ X for every entry in ONE persistent launch (the weight gradients of an ST block). Entries whose shape the
    grouped kernel does not take (Nw % 256 != 0) go through gemm_wgrad."""
    import ctypes
    take = []
    for G, X, dW in group:
        tokens, Mw = G.shape
        Nw = X.shape[1]
        assert X.shape[0] == tokens and dW.shape == (Mw, Nw) and dW.dtype == F32 and dW.stride(1) == 1
        if Mw % 128 == 0 and Nw % 256 == 0 and (not take or take[0][0].shape[0] == tokens):
            take.append((G, X, dW))
        else:
            gemm_wgrad(G, X, dW)
    for lo in range(0, len(take), 8):
        part = take[lo:lo + 8]
        n = len(part)
        vp, ll, ci = ctypes.c_void_p * n, ctypes.c_longlong * n, ctypes.c_int * n
        args = (vp(*[g.data_ptr() for g, _, _ in part]), ll(*[g.stride(0) for g, _, _ in part]),
                vp(*[x.data_ptr() for _, x, _ in part]), ll(*[x.stride(0) for _, x, _ in part]),
                ci(*[g.shape[1] for g, _, _ in part]), ci(*[x.shape[1] for _, x, _ in part]),
                vp(*[w.data_ptr() for _, _, w in part]), ll(*[w.stride(0) for _, _, w in part]))
        tokens = part[0][0].shape[0]
        flops = sum(2.0 * tokens * g.shape[1] * x.shape[1] for g, x, _ in part)
        cast = lambda a: ctypes.cast(a, ctypes.c_void_p)  # noqa: E731
        _call("gemm_wgrad_grouped", flops, "hma_gemm_wgrad_grouped", n, cast(args[0]), cast(args[1]), cast(args[2]), cast(args[3]),
              tokens, cast(args[4]), cast(args[5]), cast(args[6]), cast(args[7]), _s())


def ln_fwd(x: torch.Tensor, mode: int, *, gamma=None, beta=None, mod=None, rows_per_group: int = 0, eps: float = 1e-5,
           want_stats: bool = False, out: Optional[torch.Tensor] = None, rows: Optional[int] = None,
           src_group: int = 0, dst_group: int = 0):
    """fp32 [rows,256] -> bf16 [rows,256]; mode 0 cast, 1 affine LN, 2 LN + (1+scale)*x + shift."""
    assert x.dtype == F32 and x.shape[-1] == 256
    n_rows = x.shape[0] if rows is None else rows
    if out is None:
        out = torch.empty(n_rows, 256, device=x.device, dtype=BF16)
    stats = torch.empty(n_rows, 2, device=x.device, dtype=F32) if want_stats else None
    _call(f"ln_fwd[mode={mode}]", n_rows * 256 * 6.0, "hma_ln_fwd", x.data_ptr(), x.stride(0), n_rows, mode, _p(gamma), _p(beta), _p(mod), rows_per_group,
              float(eps), out.data_ptr(), out.stride(0), _p(stats), src_group, dst_group, _s())
    return (out, stats) if want_stats else out


def ln_bwd(dy: torch.Tensor, x: torch.Tensor, stats: torch.Tensor, mode: int, dx: torch.Tensor, *, gamma=None,
           mod=None, rows_per_group: int = 0, dgamma=None, dbeta=None, dmod=None, want_next: bool = False,
           colsum_next=None):
    """dx += LN backward; optionally returns bf16(dx) for the next stage and accumulates its column sums."""
    rows = dy.shape[0]
    dy_next = torch.empty(rows, 256, device=dy.device, dtype=BF16) if want_next else None
    _call(f"ln_bwd[mode={mode}]", rows * 256 * (16.0 if want_next else 14.0), "hma_ln_bwd", dy.data_ptr(),
          dy.stride(0), _p(x), x.stride(0) if x is not None else 0, _p(stats), rows, mode, _p(gamma), _p(mod),
          rows_per_group, dx.data_ptr(), dx.stride(0), _p(dgamma), _p(dbeta), _p(dmod), _p(dy_next),
          _p(colsum_next) if want_next else None, _s())
    return dy_next


def qk_norm_fwd(qkv: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """[LN32(q) | LN32(k) | v] of the bf16 projection output qkv [rows, 768] (qk_norm=True)."""
    out = torch.empty_like(qkv)
    _call("qk_norm_fwd", qkv.numel() * 4.0, "hma_qk_norm_fwd", qkv.data_ptr(), qkv.stride(0), qkv.shape[0], gamma.data_ptr(),
          beta.data_ptr(), float(eps), out.data_ptr(), out.stride(0), _s())
    return out


def qk_norm_bwd(qkv: torch.Tensor, gamma: torch.Tensor, dqkv: torch.Tensor, dgamma: torch.Tensor, dbeta: torch.Tensor,
                eps: float = 1e-5) -> None:
    """dqkv (w.r.t. the normalised matrix) -> gradient w.r.t. the raw qkv, in place; dgamma / dbeta accumulated."""
    _call("qk_norm_bwd", qkv.numel() * 6.0, "hma_qk_norm_bwd", qkv.data_ptr(), qkv.stride(0), qkv.shape[0], gamma.data_ptr(),
          float(eps), dqkv.data_ptr(), dqkv.stride(0), dgamma.data_ptr(), dbeta.data_ptr(), _s())


def colsum_bf16(G: torch.Tensor, out: torch.Tensor) -> None:
    _call("colsum_bf16", G.numel() * 2.0, "hma_colsum_bf16", G.data_ptr(), G.stride(0), G.shape[0], G.shape[1], out.data_ptr(), _s())


def colsum_f32(G: torch.Tensor, out: torch.Tensor) -> None:
    _call("small", 0.0, "hma_colsum_f32", G.data_ptr(), G.stride(0), G.shape[0], G.shape[1], out.data_ptr(), _s())


def cast_transpose(W: torch.Tensor, want_plain: bool = True, want_t: bool = True, alpha: float = 1.0):
    """fp32 [R,C] -> (bf16 [R,C], bf16 [C,R])."""
    assert W.dtype == F32 and W.dim() == 2 and W.is_contiguous()
    R, C = W.shape
    Wb = torch.empty(R, C, device=W.device, dtype=BF16) if want_plain else None
    Wt = torch.empty(C, R, device=W.device, dtype=BF16) if want_t else None
    _call("cast_transpose", W.numel() * 8.0, "hma_cast_transpose", W.data_ptr(), R, C, _p(Wb), _p(Wt), float(alpha), _s())
    return Wb, Wt


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    assert x.dtype == F32 and x.is_contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    _call("cast_bf16", x.numel() * 6.0, "hma_cast_bf16", x.data_ptr(), y.data_ptr(), x.numel(), _s())
    return y


def cast_colsum(x: torch.Tensor, colsum: Optional[torch.Tensor]) -> torch.Tensor:
    """bf16 copy of fp32 [rows,256] and colsum[256] += its column sums (bias gradient of the consumer)."""
    assert x.dtype == F32 and x.is_contiguous() and x.shape[1] == 256
    y = torch.empty(x.shape, device=x.device, dtype=BF16)
    _call("cast_colsum", x.numel() * 6.0, "hma_cast_colsum", x.data_ptr(), y.data_ptr(), x.shape[0], _p(colsum), _s())
    return y


def cast_transpose_batched(desc: torch.Tensor, count: int, max_rows: int, max_cols: int) -> None:
    _call("cast_transpose_batched", 0.0, "hma_cast_transpose_batched", desc.data_ptr(), count, max_rows, max_cols, _s())


def action_prep(a: torch.Tensor, kpad: int, mean=None, std=None) -> torch.Tensor:
    """(a - mean)/(std + 1e-10) per action-dim chunk, zero-padded to kpad columns, bf16."""
    assert a.dtype == F32 and a.dim() == 2 and a.is_contiguous()
    rows, da = a.shape
    y = torch.empty(rows, kpad, device=a.device, dtype=BF16)
    _call("small", 0.0, "hma_action_prep", a.data_ptr(), rows, da, _p(mean), _p(std), mean.numel() if mean is not None else 0,
              y.data_ptr(), kpad, _s())
    return y


def ln_relu_fwd(x: torch.Tensor, gamma, beta, eps: float = 1e-5):
    rows = x.shape[0]
    y = torch.empty(rows, 256, device=x.device, dtype=BF16)
    stats = torch.empty(rows, 2, device=x.device, dtype=F32)
    _call("small", 0.0, "hma_ln_relu_fwd", x.data_ptr(), rows, gamma.data_ptr(), beta.data_ptr(), float(eps), y.data_ptr(),
              stats.data_ptr(), _s())
    return y, stats


def ln_relu_bwd(dy, x, stats, gamma, beta, dgamma, dbeta) -> torch.Tensor:
    dx = torch.empty_like(x)
    _call("small", 0.0, "hma_ln_relu_bwd", dy.data_ptr(), x.data_ptr(), stats.data_ptr(), x.shape[0], gamma.data_ptr(),
              beta.data_ptr(), dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), _s())
    return dx


def group_add(x: torch.Tensor, v: torch.Tensor, rows_per_group: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x[r] + v[r // rows_per_group] for fp32 rows of 256 (additive action conditioning)."""
    if out is None:
        out = torch.empty_like(x)
    _call("group_add", x.numel() * 8.0, "hma_group_add", x.data_ptr(), v.data_ptr(), out.data_ptr(), x.shape[0], rows_per_group, _s())
    return out


def group_colsum(dx: torch.Tensor, dv: torch.Tensor, rows_per_group: int) -> None:
    """dv[g] += column sums of the rows of group g of dx (fp32 rows of 256)."""
    _call("group_colsum", dx.numel() * 4.0, "hma_group_colsum", dx.data_ptr(), dv.data_ptr(), dx.shape[0] // rows_per_group,
          rows_per_group, _s())


def rows_scatter(src: torch.Tensor, frames: int, S: int, n: int) -> torch.Tensor:
    dst = torch.empty(frames * n, 256, device=src.device, dtype=F32)
    _call("rows_scatter", frames * n * 256 * 8.0, "hma_rows_scatter", src.data_ptr(), dst.data_ptr(), frames, S, n, _s())
    return dst


def attn_spatial_fwd(qkv: torch.Tensor, frames: int, n: int, heads: int, scale: float, want_lse: bool):
    C = heads * 32
    out = torch.empty(frames * n, C, device=qkv.device, dtype=BF16)
    lse = torch.empty(frames, heads, n, device=qkv.device, dtype=F32) if want_lse else None
    _call("attn_spatial_fwd", 4.0 * frames * heads * n * n * 32, "hma_attn_spatial_fwd", qkv.data_ptr(), qkv.stride(0), frames, n, heads, 0, C, 2 * C, float(scale),
              out.data_ptr(), C, _p(lse), _s())
    return out, lse


def attn_spatial_bwd(qkv, out, dout, lse, frames: int, n: int, heads: int, scale: float, delta: Optional[torch.Tensor] = None) -> torch.Tensor:
    """delta: fp32 [frames*n, heads] = rowsum(dout * out) per (token, head) (gemm_nt's rowdot), or None to let the kernel
    compute it from `out`."""
    C = heads * 32
    dqkv = torch.empty_like(qkv)
    _call("attn_spatial_bwd", 10.0 * frames * heads * n * n * 32, "hma_attn_spatial_bwd", qkv.data_ptr(), qkv.stride(0),
          _p(out) if delta is None else None, out.stride(0) if (out is not None and delta is None) else 0, dout.data_ptr(),
          dout.stride(0), lse.data_ptr(), frames, n, heads, 0, C, 2 * C, float(scale), dqkv.data_ptr(), dqkv.stride(0),
          _p(delta), _s())
    return dqkv


def attn_temporal_fwd(qkv: torch.Tensor, B: int, T: int, n: int, heads: int, scale: float, want_lse: bool = False):
    C = heads * 32
    out = torch.empty(B * T * n, C, device=qkv.device, dtype=BF16)
    lse = torch.empty(B * T * n, heads, device=qkv.device, dtype=F32) if want_lse else None
    _call("attn_temporal_fwd", B * T * n * C * 8.0, "hma_attn_temporal_fwd", qkv.data_ptr(), qkv.stride(0), B, T, n,
          heads, 0, C, 2 * C, float(scale), out.data_ptr(), C, _p(lse), _s())
    return out, lse


def attn_temporal_bwd(qkv, out, dout, lse, B: int, T: int, n: int, heads: int, scale: float) -> torch.Tensor:
    C = heads * 32
    dqkv = torch.empty_like(qkv)
    _call("attn_temporal_bwd", B * T * n * C * 18.0, "hma_attn_temporal_bwd", qkv.data_ptr(), qkv.stride(0),
          out.data_ptr(), out.stride(0), dout.data_ptr(), dout.stride(0), _p(lse), B, T, n, heads, 0, C, 2 * C,
          float(scale), dqkv.data_ptr(), dqkv.stride(0), _s())
    return dqkv


def kv_cache_append(qkv: torch.Tensor, B: int, frames: int, n: int, kv: torch.Tensor, t0: int) -> None:
    """K/V columns of `frames` frames of qkv ((b, t, s) order, [B*frames*n, 768]) -> cache frames [t0, t0+frames).
    kv: bf16 [Tmax, B*n, 512]."""
    C = 256
    assert kv.dtype == BF16 and kv.shape[1] == B * n and kv.shape[2] == 2 * C and kv.is_contiguous()
    assert t0 + frames <= kv.shape[0]
    _call("kv_cache_append", B * frames * n * 2 * C * 4.0, "hma_kv_cache_append", qkv.data_ptr(), qkv.stride(0), C, 2 * C,
          B, frames, n, kv.data_ptr(), kv.stride(0), t0, _s())


def attn_temporal_cached(qkv: torch.Tensor, kv: Optional[torch.Tensor], n_prev: int, heads: int, scale: float,
                         frames: int = 1, n: int = 0) -> torch.Tensor:
    """Temporal attention of a pass of `frames` window frames per sample ([B*frames*n, 768] q|k|v, (b, f, s) order): frame f
    attends to cache frames [0, n_prev + f) + itself (the pass has appended its own frames to the cache already)."""
    C = heads * 32
    rows = qkv.shape[0]
    out = torch.empty(rows, C, device=qkv.device, dtype=BF16)
    reads = n_prev + (frames - 1) / 2.0
    _call("attn_temporal_cached", rows * C * 2.0 * (2 * reads + 4), "hma_attn_temporal_cached", qkv.data_ptr(), qkv.stride(0),
          0, C, 2 * C, _p(kv) if (n_prev or frames > 1) else None, kv.stride(0) if kv is not None else 0, rows, n_prev, heads,
          float(scale), out.data_ptr(), C, frames, n, _s())
    return out


def embed_fwd(ids, E0, E1, mask_embed, act, pos, pos_n: int, B: int, T: int, S: int, A: int, vs: int, mask_id: int):
    x = torch.empty(B * T * (S + A), 256, device=ids.device, dtype=F32)
    _call("embed_fwd", B * T * (S + A) * 256 * 8.0, "hma_embed_fwd", ids.data_ptr(), E0.data_ptr(), _p(E1), mask_embed.data_ptr(), _p(act), pos.data_ptr(),
              pos_n, B, T, S, A, vs, mask_id, x.data_ptr(), _s())
    return x


def embed_bwd(ids, dx, pos_n, B, T, S, A, vs, mask_id, dE0, dE1, dmask, dact, dpos) -> None:
    _call("embed_bwd", B * T * (S + A) * 256 * 8.0, "hma_embed_bwd", ids.data_ptr(), dx.data_ptr(), pos_n, B, T, S, A, vs, mask_id, dE0.data_ptr(), _p(dE1),
              dmask.data_ptr(), _p(dact), dpos.data_ptr(), _s())


def collate_maskgit(tokens, B, T, S, nv, vs, mask_id, corrupt_r, corrupt_thresh, rand_vals, first_masked_frame, frame_rates,
                    frame_r, mask_prob, mask_r):
    """One pass of the training collator (hma/data.py:28-98) given its random draws. Returns (input_ids, labels)."""
    input_ids = torch.empty_like(tokens)
    labels = torch.empty_like(tokens)
    _call("collate_maskgit", tokens.numel() * 24.0, "hma_collate_maskgit", tokens.data_ptr(), input_ids.data_ptr(), labels.data_ptr(),
          B, T, S, nv, vs, mask_id, _p(corrupt_r), float(corrupt_thresh), _p(rand_vals), first_masked_frame, _p(frame_rates),
          _p(frame_r), _p(mask_prob), _p(mask_r), _s())
    return input_ids, labels


def ce_fwd(logits, labels, input_ids, B, T, S, nv, vs, mask_id, smoothing):
    rows = B * T * S
    lse = torch.empty(rows, nv, device=logits.device, dtype=F32)
    sums = torch.empty(3, device=logits.device, dtype=F32)
    loss_acc = torch.empty(2, device=logits.device, dtype=F32)
    _call("ce_fwd", logits.numel() * 4.0, "hma_ce_fwd", logits.data_ptr(), logits.stride(0), labels.data_ptr(), input_ids.data_ptr(), B, T, S, nv,
              vs, mask_id, float(smoothing), lse.data_ptr(), sums.data_ptr(), loss_acc.data_ptr(), _s())
    return loss_acc, lse, sums


def ce_bwd(logits, labels, input_ids, B, T, S, nv, vs, mask_id, smoothing, lse, sums, dloss) -> torch.Tensor:
    dlogits = torch.empty(B * T * S, nv * vs, device=logits.device, dtype=BF16)
    _call("ce_bwd", logits.numel() * 6.0, "hma_ce_bwd", logits.data_ptr(), logits.stride(0), labels.data_ptr(), input_ids.data_ptr(), B, T, S, nv,
              vs, mask_id, float(smoothing), lse.data_ptr(), sums.data_ptr(), dloss.data_ptr(), dlogits.data_ptr(),
              dlogits.stride(0), _s())
    return dlogits


def sample_tokens(logits_frame: torch.Tensor, nv: int, vs: int, exp_noise: Optional[torch.Tensor], temperature: float = 1.0):
    """logits_frame: fp32 view [B, S, nv*vs] (arbitrary batch stride). Returns (samples i64 [B,S], conf f32 [B,S])."""
    B, S, _ = logits_frame.shape
    samples = torch.empty(B, S, device=logits_frame.device, dtype=torch.long)
    conf = torch.empty(B, S, device=logits_frame.device, dtype=F32)
    _call("sample_tokens", B * S * nv * vs * 4.0, "hma_sample_tokens", logits_frame.data_ptr(), logits_frame.stride(0), logits_frame.stride(1), B, S, nv,
              vs, _p(exp_noise), float(temperature), samples.data_ptr(), conf.data_ptr(), _s())
    return samples, conf


def rank_remask(keys, unmasked_u8, samples, frame_view, n_mask: int, mask_id: int) -> torch.Tensor:
    """frame_view: i64 view [B, S] of prompt[:, out_t] (unit stride over S). Updates unmasked and frame in place."""
    B, S = samples.shape
    assert frame_view.stride(1) == 1
    out = torch.empty_like(samples)
    _call("rank_remask", B * S * 32.0, "hma_rank_remask", _p(keys), unmasked_u8.data_ptr(), samples.data_ptr(), frame_view.data_ptr(),
              frame_view.stride(0), B, S, n_mask, mask_id, out.data_ptr(), _s())
    return out


# ----------------------------------------------------------------------------------------------
# STMAR stages (csrc/mar.cu)
# ----------------------------------------------------------------------------------------------
def mar_embed_fwd(lat, mask_u8, mask_token, xp_in, We, act, pos, pos_n: int, B: int, T: int, H: int, W: int, Cv: int, p: int,
                  A: int, fill_inplace: bool, want_xp: bool, want_rowmask: bool = False):
    Sp, D = (H // p) * (W // p), Cv * p * p
    dev = (lat if lat is not None else xp_in).device
    u = torch.empty(B * T * (Sp + A), 256, device=dev, dtype=F32)
    xp = torch.empty(B * T * Sp, D, device=dev, dtype=F32) if want_xp else None
    rowmask = torch.empty(B * T * Sp, device=dev, dtype=F32) if want_rowmask else None
    _call("mar_embed_fwd", u.numel() * 4.0, "hma_mar_embed_fwd", _p(lat), _p(mask_u8), _p(mask_token), _p(xp_in), We.data_ptr(),
          _p(act), pos.data_ptr(), pos_n, B, T, H, W, Cv, p, A, int(fill_inplace), u.data_ptr(), _p(xp), _p(rowmask), _s())
    return u, xp, rowmask


def mar_embed_bwd(du, xp, mask_u8, We, pos_n: int, B: int, T: int, H: int, W: int, Cv: int, p: int, A: int, dWe, dmask_token,
                  dact, dpos) -> None:
    _call("mar_embed_bwd", du.numel() * 4.0, "hma_mar_embed_bwd", du.data_ptr(), xp.data_ptr(), _p(mask_u8), We.data_ptr(), pos_n,
          B, T, H, W, Cv, p, A, dWe.data_ptr(), _p(dmask_token), _p(dact), dpos.data_ptr(), _s())


def mar_ln_fwd(x: torch.Tensor, *, gamma=None, beta=None, eps: float = 1e-6, mod=None, shift_off: int = 0, scale_off: int = 0,
               add=None, want32: bool = False, want16: bool = True, want_stats: bool = False, gate=None):
    """LayerNorm over the last dim (256 or 1024) of fp32 x, optional affine / bf16 modulation / additive row table.
    gate = (gmod, gate_off, h2, xsum): normalise x + gmod[:, gate_off:] * h2 instead and write that sum to xsum."""
    assert x.dtype == F32 and x.is_contiguous() and x.dim() == 2
    rows, C = x.shape
    y32 = torch.empty(rows, C, device=x.device, dtype=F32) if want32 else None
    y16 = torch.empty(rows, C, device=x.device, dtype=BF16) if want16 else None
    stats = torch.empty(rows, 2, device=x.device, dtype=F32) if want_stats else None
    if add is not None:
        assert add.dtype == F32 and add.is_contiguous() and add.shape[-1] == C
    _call(f"mar_ln_fwd[{C}]", rows * C * 6.0, "hma_mar_ln_fwd", x.data_ptr(), rows, C, _p(gamma), _p(beta), float(eps), _p(mod),
          mod.stride(0) if mod is not None else 0, shift_off, scale_off, _p(add),
          add.numel() // C if add is not None else 0, _p(y32), _p(y16), _p(stats),
          gate[0].data_ptr() if gate else None, gate[0].stride(0) if gate else 0, gate[1] if gate else 0,
          gate[2].data_ptr() if gate else None, gate[3].data_ptr() if gate else None, _s())
    return y32, y16, stats


def mar_ln_bwd(x: torch.Tensor, stats: torch.Tensor, *, dy16=None, dy32=None, gamma=None, beta=None, mod=None, shift_off: int = 0,
               scale_off: int = 0, dx32=None, accumulate: bool = False, want16: bool = False, dgamma=None, dbeta=None, dmod=None,
               dadd=None):
    rows, C = x.shape
    dx16 = torch.empty(rows, C, device=x.device, dtype=BF16) if want16 else None
    _call(f"mar_ln_bwd[{C}]", rows * C * 10.0, "hma_mar_ln_bwd", _p(dy16), _p(dy32), x.data_ptr(), stats.data_ptr(), rows, C,
          _p(gamma), _p(beta), _p(mod), mod.stride(0) if mod is not None else 0, shift_off, scale_off, _p(dx32),
          int(accumulate), _p(dx16), _p(dgamma), _p(dbeta), _p(dmod), dmod.stride(0) if dmod is not None else 0, _p(dadd),
          dadd.numel() // C if dadd is not None else 0, _s())
    return dx16


def mar_gate_fwd(x: torch.Tensor, mod: torch.Tensor, gate_off: int, h2: torch.Tensor) -> torch.Tensor:
    rows, C = x.shape
    out = torch.empty_like(x)
    _call("mar_gate_fwd", rows * C * 12.0, "hma_mar_gate_fwd", x.data_ptr(), mod.data_ptr(), mod.stride(0), gate_off, h2.data_ptr(),
          rows, C, out.data_ptr(), _s())
    return out


def mar_gate_bwd(dx: torch.Tensor, mod: torch.Tensor, gate_off: int, h2: torch.Tensor, dmod: torch.Tensor) -> torch.Tensor:
    rows, C = dx.shape
    dh2 = torch.empty(rows, C, device=dx.device, dtype=BF16)
    _call("mar_gate_bwd", rows * C * 12.0, "hma_mar_gate_bwd", dx.data_ptr(), mod.data_ptr(), mod.stride(0), gate_off,
          h2.data_ptr(), rows, C, dh2.data_ptr(), dmod.data_ptr(), dmod.stride(0), _s())
    return dh2


def mar_silu_fwd(y: torch.Tensor, rowvec: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows, C = y.shape
    out = torch.empty(rows, C, device=y.device, dtype=BF16)
    _call("mar_silu_fwd", rows * C * 6.0, "hma_mar_silu_fwd", y.data_ptr(), _p(rowvec), rows, C, out.data_ptr(), _s())
    return out


def mar_silu_steps(c: torch.Tensor, te: torch.Tensor) -> torch.Tensor:
    """bf16 [steps*n, C] = SiLU(c[r] + te[i]) in step-major order."""
    n, C = c.shape
    steps = te.shape[0]
    out = torch.empty(steps * n, C, device=c.device, dtype=BF16)
    _call("mar_silu_fwd", steps * n * C * 2.0, "hma_mar_silu_steps", c.data_ptr(), te.data_ptr(), n, steps, C, out.data_ptr(), _s())
    return out


def mar_silu_bwd(dsy: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    dy = torch.empty(y.shape, device=y.device, dtype=BF16)
    _call("mar_silu_bwd", y.numel() * 10.0, "hma_mar_silu_bwd", dsy.data_ptr(), y.data_ptr(), y.numel(), dy.data_ptr(), _s())
    return dy


def mar_q_sample(x0: torch.Tensor, noise, t, tables, kpad: int) -> torch.Tensor:
    N, D = x0.shape
    out = torch.empty(N, kpad, device=x0.device, dtype=BF16)
    _call("small", 0.0, "hma_mar_q_sample", x0.data_ptr(), _p(noise), _p(t), _p(tables), N, D, kpad, out.data_ptr(), _s())
    return out


def mar_timestep_embed(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.int64 and t.is_contiguous()
    out = torch.empty(t.numel(), 256, device=t.device, dtype=BF16)
    _call("small", 0.0, "hma_mar_timestep_embed", t.data_ptr(), t.numel(), out.data_ptr(), _s())
    return out


def mar_diff_loss_fwd(out: torch.Tensor, x0, noise, t, mask, tables, D: int, want_rows: bool = False):
    N = x0.shape[0]
    sums = torch.zeros(2, device=x0.device, dtype=F32)
    loss = torch.empty((), device=x0.device, dtype=F32)
    rows = torch.empty(N, device=x0.device, dtype=F32) if want_rows else None
    _call("mar_diff_loss_fwd", N * D * 16.0, "hma_mar_diff_loss_fwd", out.data_ptr(), out.stride(0), x0.data_ptr(), noise.data_ptr(),
          t.data_ptr(), _p(mask), tables.data_ptr(), N, D, _p(rows), sums.data_ptr(), loss.data_ptr(), _s())
    return loss, sums, rows


def mar_diff_loss_bwd(out: torch.Tensor, x0, noise, t, mask, tables, D: int, sums, dloss, ldd: int) -> torch.Tensor:
    N = x0.shape[0]
    dout = torch.empty(N, ldd, device=x0.device, dtype=BF16)
    _call("mar_diff_loss_bwd", N * D * 16.0, "hma_mar_diff_loss_bwd", out.data_ptr(), out.stride(0), x0.data_ptr(), noise.data_ptr(),
          t.data_ptr(), _p(mask), tables.data_ptr(), N, D, sums.data_ptr(), _p(dloss), dout.data_ptr(), ldd, _s())
    return dout


def mar_p_sample(out: torch.Tensor, x: torch.Tensor, noise, tables, step: int, temperature: float, clip: bool, x_next: torch.Tensor,
                 x16: Optional[torch.Tensor]) -> None:
    N, D = x.shape
    _call("small", 0.0, "hma_mar_p_sample", out.data_ptr(), out.stride(0), x.data_ptr(), _p(noise), tables.data_ptr(), step, N, D,
          float(temperature), int(clip), x_next.data_ptr(), _p(x16), x16.shape[1] if x16 is not None else D, _s())


def mar_sampler(xt: torch.Tensor, noise: torch.Tensor, tables: torch.Tensor, mods: torch.Tensor, mods_step0: int, step_hi: int,
                step_lo: int, temperature: float, clip: bool, w_in_t: torch.Tensor, b_in: torch.Tensor, w1, w2, ln_g, ln_b, b1, b2,
                w_f: torch.Tensor, b_f: torch.Tensor, work: dict, dbg_out: Optional[torch.Tensor] = None) -> None:
    """Ancestral steps step_hi-1 ... step_lo of the diffusion head's sampler in ONE persistent launch (csrc/mar_sampler.cu);
    xt fp32 [R, D] is updated in place. w1 / w2 / ln_g / ln_b / b1 / b2: lists of `depth` tensors. work: dict of workspaces
    (x fp32, u16 / a16 / h2 bf16 [R, 1024], barrier int32 [1]) owned by the caller (static under graph capture)."""
    import ctypes
    R, D = xt.shape
    depth = len(w1)
    assert xt.dtype == F32 and xt.is_contiguous() and noise.dtype == F32 and noise.is_contiguous() and noise.shape[1:] == (R, D)
    assert mods.dtype == BF16 and mods.stride(1) == 1 and tables.dtype == F32 and tables.is_contiguous()
    for t in list(w1) + list(w2):
        assert t.dtype == BF16 and t.shape == (1024, 1024) and t.is_contiguous()
    for t in list(ln_g) + list(ln_b) + list(b1) + list(b2):
        assert t.dtype == F32 and t.numel() == 1024 and t.is_contiguous()
    assert w_in_t.dtype == BF16 and w_in_t.shape == (D, 1024) and w_in_t.is_contiguous() and w_f.dtype == BF16 and w_f.is_contiguous()
    assert w_f.shape[0] >= 2 * D and w_f.shape[1] == 1024 and b_f.numel() >= 2 * D
    vp = ctypes.c_void_p * depth
    arr = lambda ts: ctypes.cast(vp(*[t.data_ptr() for t in ts]), ctypes.c_void_p)  # noqa: E731
    import os
    _call("mar_sampler", 0.0, "hma_mar_sampler", R, D, depth, step_hi, step_lo, float(temperature), int(clip) | int(os.environ.get("HMA_SAMPLER_DBG", "0")), xt.data_ptr(),
          noise.data_ptr(), tables.data_ptr(), mods.data_ptr(), mods.stride(0), mods_step0, w_in_t.data_ptr(),
          b_in.data_ptr(), arr(w1), arr(w2), arr(ln_g), arr(ln_b), arr(b1), arr(b2), w_f.data_ptr(), b_f.data_ptr(),
          work["x"].data_ptr(), work["u16"].data_ptr(), work["a16"].data_ptr(), work["h2"].data_ptr(), _p(dbg_out),
          work["barrier"].data_ptr(), _s())


def mar_gather_rows(src: torch.Tensor, idx: torch.Tensor, want32: bool, want16: bool):
    assert src.dtype == F32 and src.is_contiguous() and idx.dtype == torch.int32
    n, C = idx.numel(), src.shape[-1]
    d32 = torch.empty(n, C, device=src.device, dtype=F32) if want32 else None
    d16 = torch.empty(n, C, device=src.device, dtype=BF16) if want16 else None
    _call("small", 0.0, "hma_mar_gather_rows", src.data_ptr(), idx.data_ptr(), n, C, _p(d32), _p(d16), _s())
    return d32, d16


def mar_scatter_rows(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> None:
    assert src.dtype == F32 and dst.dtype == F32 and src.is_contiguous() and dst.is_contiguous() and idx.dtype == torch.int32
    _call("small", 0.0, "hma_mar_scatter_rows", src.data_ptr(), idx.data_ptr(), idx.numel(), src.shape[-1], dst.data_ptr(), _s())


def dropout_bf16_(x: torch.Tensor, p: float, seed: int, seed_dev: Optional[torch.Tensor] = None) -> None:
    assert x.dtype == BF16 and x.is_contiguous()
    _call("dropout", x.numel() * 4.0, "hma_dropout_bf16", x.data_ptr(), x.numel(), float(p), seed, _p(seed_dev), _s())


def dropout_add_f32(a: torch.Tensor, resid: torch.Tensor, p: float, seed: int, out: Optional[torch.Tensor] = None,
                    seed_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert a.dtype == F32 and resid.dtype == F32 and a.is_contiguous() and resid.is_contiguous()
    if out is None:
        out = torch.empty_like(resid)
    _call("dropout", a.numel() * 12.0, "hma_dropout_add_f32", a.data_ptr(), resid.data_ptr(), out.data_ptr(), a.numel(), float(p), seed, _p(seed_dev), _s())
    return out


def dropout_cast_bf16(a: torch.Tensor, p: float, seed: int, seed_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert a.dtype == F32 and a.is_contiguous()
    out = torch.empty(a.shape, device=a.device, dtype=BF16)
    _call("dropout", a.numel() * 6.0, "hma_dropout_cast_bf16", a.data_ptr(), out.data_ptr(), a.numel(), float(p), seed, _p(seed_dev), _s())
    return out
