"""Isolated micro-benchmarks of the hot stages at the benchmark shapes, for quick kernel iteration on a B200:

    gpurun --timeout 120 -- 'python tools/kbench.py [name ...] [--iters 30]'

Each stage is launched through the same hma_b200.ops wrapper the engine uses, timed with CUDA events on the launching
stream after 3 warm-up launches, with a 512 MB write between iterations so that every launch starts with a cold L2 (as it
does inside a training step, whose per-layer working set is ~0.6 GB). Prints the mean launch time, the achieved rate on the
stage's own bound and the fraction of the measured peak (MEASURED_PEAKS.json). A 20-second call instead of a 2-minute bench.
Not part of tests / bench.py."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hma_b200 import ops  # noqa: E402
from hma_b200.ops import EPI_BF16, EPI_DGELU, EPI_GELU, EPI_RESID  # noqa: E402

DEV = "cuda"
B, T, n, H, C = 8, 16, 320, 8, 256          # config 2
N, M = B * T * n, B * T
SCALE = 32 ** -0.5
BF = torch.bfloat16


def rnd(*shape, dtype=BF, s=1.0):
    return (torch.randn(*shape, device=DEV) * s).to(dtype)


def stages():
    """name -> (callable, algorithmic work, 'flop' | 'byte')"""
    st = {}
    rnd_a, rnd_w = rnd(N, C), rnd(C, C, s=0.05)
    qkv = rnd(N, 3 * C)
    out_s, lse = ops.attn_spatial_fwd(qkv, M, n, H, SCALE, want_lse=True)
    dout = rnd(N, C)
    st["attn_spatial_fwd"] = (lambda: ops.attn_spatial_fwd(qkv, M, n, H, SCALE, want_lse=True), 4.0 * M * H * n * n * 32, "flop")
    delta = (dout.float() * out_s.float()).view(N, H, 32).sum(-1).contiguous()
    st["attn_spatial_bwd"] = (lambda: ops.attn_spatial_bwd(qkv, None, dout, lse, M, n, H, SCALE, delta=delta), 10.0 * M * H * n * n * 32, "flop")
    st["attn_spatial_bwd_nodelta"] = (lambda: ops.attn_spatial_bwd(qkv, out_s, dout, lse, M, n, H, SCALE), 10.0 * M * H * n * n * 32, "flop")
    st["gemm_datt_rowdot"] = (lambda: ops.gemm_nt(rnd_a, rnd_w, EPI_BF16, aux=out_s, rowdot=delta), N * C * (2.0 + 2 + 2), "byte")
    st["gemm_datt_plain"] = (lambda: ops.gemm_nt(rnd_a, rnd_w, EPI_BF16), N * C * (2.0 + 2), "byte")
    out_t, _ = ops.attn_temporal_fwd(qkv, B, T, n, H, SCALE)
    st["attn_temporal_fwd"] = (lambda: ops.attn_temporal_fwd(qkv, B, T, n, H, SCALE), N * C * 2.0 * 4, "byte")
    st["attn_temporal_bwd"] = (lambda: ops.attn_temporal_bwd(qkv, out_t, dout, None, B, T, n, H, SCALE), N * C * 2.0 * 7, "byte")
    rows_dec = 64 * n  # one frame of a batch-64 decode pass against 11 cached frames (the mean of config 3)
    kv = rnd(16, rows_dec, 2 * C)
    qkv_dec = rnd(rows_dec, 3 * C)
    st["attn_temporal_cached"] = (lambda: ops.attn_temporal_cached(qkv_dec, kv, 11, H, SCALE), rows_dec * C * 2.0 * (2 * 11 + 4), "byte")
    a256, a1024 = rnd(N, C), rnd(N, 4 * C)
    w_qkv, w_fc1, w_fc2, w_proj = rnd(3 * C, C, s=0.05), rnd(4 * C, C, s=0.05), rnd(C, 4 * C, s=0.05), rnd(C, C, s=0.05)
    x32 = rnd(N, C, dtype=torch.float32)
    z = rnd(N, 4 * C)
    st["gemm_qkv"] = (lambda: ops.gemm_nt(a256, w_qkv, EPI_BF16), 2.0 * N * 3 * C * C, "flop")
    st["gemm_fc1_gelu"] = (lambda: ops.gemm_nt(a256, w_fc1, EPI_GELU, out2=z), 2.0 * N * 4 * C * C, "flop")
    st["gemm_fc2_resid"] = (lambda: ops.gemm_nt(a1024, w_fc2, EPI_RESID, resid=x32), 2.0 * N * 4 * C * C, "flop")
    st["gemm_proj_resid"] = (lambda: ops.gemm_nt(a256, w_proj, EPI_RESID, resid=x32), N * C * (2.0 + 4 + 4), "byte")
    gam, bet = rnd(C, dtype=torch.float32), rnd(C, dtype=torch.float32)
    mod = rnd(M, 2 * C, dtype=torch.float32, s=0.2)
    st["gemm_proj_resid_ln1"] = (lambda: ops.gemm_nt_ln(a256, w_proj, resid=x32, ln_mode=1, gamma=gam, beta=bet, want_stats=True),
                                 N * C * (2.0 + 4 + 4 + 2), "byte")
    st["gemm_proj_resid_ln2"] = (lambda: ops.gemm_nt_ln(a256, w_proj, resid=x32, ln_mode=2, mod=mod, rows_per_group=n, eps=1e-6,
                                                        want_stats=True), N * C * (2.0 + 4 + 4 + 2), "byte")
    st["gemm_fc2_resid_ln1"] = (lambda: ops.gemm_nt_ln(a1024, w_fc2, resid=x32, ln_mode=1, gamma=gam, beta=bet, want_stats=True),
                                2.0 * N * 4 * C * C, "flop")
    st["ln_fwd_mode1"] = (lambda: ops.ln_fwd(x32, 1, gamma=gam, beta=bet, want_stats=True), N * C * 6.0, "byte")
    st["gemm_dgelu"] = (lambda: ops.gemm_nt(a256, w_fc2.t().contiguous(), EPI_DGELU, aux=z), 2.0 * N * 4 * C * C, "flop")
    dw = torch.zeros(4 * C, C, device=DEV)
    dw2 = torch.zeros(C, C, device=DEV)
    st["wgrad_1024x256"] = (lambda: ops.gemm_wgrad(a1024, a256, dw), 2.0 * N * 4 * C * C, "flop")
    st["wgrad_256x256"] = (lambda: ops.gemm_wgrad(a256, dout, dw2), 2.0 * N * C * C, "flop")
    shapes = [(C, 4 * C), (4 * C, C), (C, C), (3 * C, C), (C, C), (C, C), (3 * C, C)]
    grp = [(rnd(N, mw), rnd(N, nw), torch.zeros(mw, nw, device=DEV)) for mw, nw in shapes]
    st["wgrad_layer_grouped"] = (lambda: ops.gemm_wgrad_grouped(grp), sum(2.0 * N * mw * nw for mw, nw in shapes), "flop")

    def seven():
        for G_, X_, W_ in grp:
            ops.gemm_wgrad(G_, X_, W_)
    st["wgrad_layer_seven_launches"] = (seven, sum(2.0 * N * mw * nw for mw, nw in shapes), "flop")
    gam, bet = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    st["ln_fwd"] = (lambda: ops.ln_fwd(x32, 1, gamma=gam, beta=bet, want_stats=True), N * C * 6.0, "byte")
    _, stats = ops.ln_fwd(x32, 1, gamma=gam, beta=bet, want_stats=True)
    dx = torch.zeros(N, C, device=DEV)
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    st["ln_bwd"] = (lambda: ops.ln_bwd(a256, x32, stats, 1, dx, gamma=gam, dgamma=dg, dbeta=db, want_next=True), N * C * 16.0, "byte")
    # HMA-MAR diffusion head (config 4: 6144 rows of width 1024)
    R, Wd = 6144, 1024
    xm = rnd(R, Wd, dtype=torch.float32)
    mod = rnd(R, 3 * Wd, s=0.3)
    g1, b1 = torch.ones(Wd, device=DEV), torch.zeros(Wd, device=DEV)
    st["mar_ln_fwd_1024"] = (lambda: ops.mar_ln_fwd(xm, gamma=g1, beta=b1, mod=mod, shift_off=0, scale_off=Wd, want_stats=True),
                             R * Wd * (4.0 + 4 + 2), "byte")
    _, _, stm = ops.mar_ln_fwd(xm, gamma=g1, beta=b1, mod=mod, shift_off=0, scale_off=Wd, want_stats=True)
    dym, dxm, dmod = rnd(R, Wd), torch.zeros(R, Wd, device=DEV), torch.empty(R, 3 * Wd, device=DEV, dtype=BF)
    dg1, db1 = torch.zeros(Wd, device=DEV), torch.zeros(Wd, device=DEV)
    st["mar_ln_bwd_1024"] = (lambda: ops.mar_ln_bwd(xm, stm, dy16=dym, gamma=g1, beta=b1, mod=mod, shift_off=0, scale_off=Wd, dx32=dxm,
                                                   accumulate=True, dgamma=dg1, dbeta=db1, dmod=dmod), R * Wd * (2.0 + 4 + 2 + 8 + 4), "byte")
    sy = rnd(512, Wd)
    w_ada = rnd(3 * Wd, Wd, s=0.03)
    st["gemm_sampler_512x3072"] = (lambda: ops.gemm_nt(sy, w_ada, EPI_BF16), 2.0 * 512 * 3 * Wd * Wd, "flop")
    return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf, gbs = peaks.get("bf16_tflops_sustained", 1400.0), peaks.get("hbm_gbs", 6500.0)
    st = stages()
    flush = torch.empty(512 << 20, device=DEV, dtype=torch.uint8)
    for name, (fn, work, kind) in st.items():
        if args.names and not any(k in name for k in args.names):
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(args.iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        us = tot / args.iters * 1e3
        rate = work / (us * 1e-6)
        if kind == "flop":
            print(f"{name:26s} {us:8.1f} us  {rate / 1e12:8.1f} TFLOP/s  {rate / 1e12 / tf:6.1%} of {tf:.0f}")
        else:
            print(f"{name:26s} {us:8.1f} us  {rate / 1e9:8.0f} GB/s     {rate / 1e9 / gbs:6.1%} of {gbs:.0f}")


if __name__ == "__main__":
    main()
