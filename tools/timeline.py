#!/usr/bin/env python
"""Per-CTA event timeline of one kernel (development aid). The HMA_TL(ev, i) stamps are added to a kernel by the
patch scripts in tools/instr/ (apply, build with HMA_B200_TIMELINE=1, run this on the GPU, `git checkout` the .cu):
    python tools/instr/gemm_nt.py && HMA_B200_TIMELINE=1 python -m hma_b200.build --force && python tools/timeline.py gemm_nt
Prints clock64() stamps of CTA 0 relative to its first event, one row per loop iteration."""
import ctypes, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hma_b200 import ops, _lib

KERNELS = {
    "attn_spatial_fwd": dict(names={0: "start", 11: "loaded", 1: "top", 2: "Sa_ready", 3: "Pa_arrived", 4: "Sb_ready", 5: "Pb_arrived",
                                    6: "O_ready", 7: "stored", 10: "end", 12: "i:Sa_issued", 13: "i:Pa_seen", 14: "i:PVa+Sb_issued",
                                    15: "i:Pb_seen"}, single=(0, 10, 11)),
    "gemm_nt": dict(single=(0, 10), names={0: "start", 10: "end", 1: "mma:top", 2: "mma:tmem_free", 3: "mma:committed", 4: "epi:top",
                                           5: "epi:acc_ready", 6: "epi:done", 7: "ln:sent", 8: "ln:wait", 9: "ln:got", 11: "ln:done"}),
    "mar_sampler": dict(single=(), names={0: "sync_exit(e)", 7: "sync_enter(e)", 1: "P:first", 2: "P:last", 3: "M:full0", 8: "M:full1",
                                           4: "M:full15", 5: "E:tfull", 6: "E:done"}),
    "gemm_wgrad": dict(single=(0, 3, 4, 5), names={0: "start", 3: "mma_done", 4: "red_issued", 5: "end", 1: "tma_issued", 2: "full"}),
    "attn_spatial_bwd": dict(names={0: "start", 11: "loaded", 1: "SdP_issued", 5: "A0:P", 6: "A3:P", 7: "A4:P", 8: "A7:P",
                                    12: "B8:dS", 13: "B9:dS", 14: "B12:dS", 15: "B15:dS",
                                    2: "dV:go", 3: "dV:issued", 4: "dK:issued", 9: "final", 10: "end"}),
}

def run_attn_spatial_bwd():
    frames, n, H = 128, 320, 8
    qkvs = [torch.randn(frames * n, 768, device="cuda").bfloat16() for _ in range(4)]  # rotate: operands come from HBM
    outs = [ops.attn_spatial_fwd(q, frames, n, H, 0.17, True) for q in qkvs]
    douts = [torch.randn_like(o[0]) for o in outs]
    deltas = [(d.float() * o[0].float()).view(frames * n, H, 32).sum(-1).contiguous() for d, o in zip(douts, outs)]
    for i in range(4):
        ops.attn_spatial_bwd(qkvs[i], None, douts[i], outs[i][1], frames, n, H, 0.17, delta=deltas[i])

def run_mar_sampler():  # tools/instr/mar_sampler.py first; stamps of the last persistent launch (64 rows)
    import runpy
    sys.argv = [sys.argv[0], "64"]
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ubench", "sampler_call.py"), run_name="__main__")
    sys.argv = [sys.argv[0], "mar_sampler"]


def run_gemm_wgrad():
    N = 40960
    Mw, Nw = int(os.environ.get("MW", 256)), int(os.environ.get("NW", 256))
    Gs = [torch.randn(N, Mw, device="cuda").bfloat16() for _ in range(6)]   # rotate buffers: operands come from HBM
    Xs = [torch.randn(N, Nw, device="cuda").bfloat16() for _ in range(6)]
    dW = torch.zeros(Mw, Nw, device="cuda")
    for i in range(6):
        ops.gemm_wgrad(Gs[i], Xs[i], dW)

def run_gemm_nt():
    M, N, K, epi = 40960, int(os.environ.get("N", 1024)), int(os.environ.get("K", 256)), int(os.environ.get("EPI", 1))
    As = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(4)]
    W = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    outs = [torch.empty(M, N, device="cuda", dtype=torch.float32 if epi == 3 else torch.bfloat16) for _ in range(2)]
    out2 = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(2)]
    if os.environ.get("LN"):
        gam, bet = torch.randn(N, device="cuda"), torch.randn(N, device="cuda")
        for i in range(4):
            ops.gemm_nt_ln(As[i], W, resid=outs[(i + 1) % 2], out=outs[i % 2], bias=bias, ln_mode=1, gamma=gam, beta=bet, want_stats=True)
        return
    for i in range(4):
        ops.gemm_nt(As[i], W, epi, out=outs[i % 2], bias=bias, out2=out2[i % 2] if epi == 1 else None,
                    aux=out2[i % 2] if epi == 2 else None, resid=outs[(i + 1) % 2] if epi == 3 else None)

def run_attn_spatial_fwd():
    frames, n, H = 128, 320, 8
    qkvs = [torch.randn(frames * n, 768, device="cuda").bfloat16() for _ in range(4)]  # rotate: operands come from HBM
    for i in range(4):
        ops.attn_spatial_fwd(qkvs[i], frames, n, H, 0.17, True)

def main():
    name = sys.argv[1]
    globals()["run_" + name]()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (16 * 64))()
    fn = getattr(_lib.lib(), "hma_timeline_" + name)
    fn.argtypes = [ctypes.c_void_p]
    assert fn(buf) == 0
    a = np.array(buf).reshape(16, 64)
    names = KERNELS[name]["names"]
    t0 = a[0, 0] or a[0, 1]
    single = [e for e in names if e in KERNELS[name].get('single', (0, 9, 10, 11))]
    print(" ".join(f"{names[e]}={a[e, 0] - t0}" for e in single))
    print("event-0 sub-stamps:", " ".join(f"[{i}]={a[0, i] - t0}" for i in range(1, 8) if a[0, i]))
    per_it = [e for e in sorted(names) if e not in single]
    for i in range(40):
        if all(a[e, i] == 0 for e in per_it):
            break
        print(i, " ".join(f"{names[e]}={a[e, i] - t0:6d}" for e in per_it))

if __name__ == "__main__":
    main()
