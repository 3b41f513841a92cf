import torch, sys, os
sys.path.insert(0, "/root/repo")
from hma_b200 import ops
from hma_b200.ops import EPI_BF16, EPI_SILU, EPI_RESID
dev = "cuda"
R, w = int(os.environ.get("R", 512)), 1024
a = (torch.randn(R, w, device=dev) * 0.1).bfloat16()
W1 = (torch.randn(w, w, device=dev) * 0.03).bfloat16()
W2 = (torch.randn(w, w, device=dev) * 0.03).bfloat16()
b = torch.zeros(w, device=dev)
x32 = torch.randn(R, w, device=dev)
def chain(n, fn):
    g = torch.cuda.CUDAGraph()
    fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 5 / n
h = torch.empty(R, w, device=dev, dtype=torch.bfloat16)
h2 = torch.empty(R, w, device=dev, dtype=torch.bfloat16)
def two():
    ops.gemm_nt(a, W1, EPI_SILU, bias=b, out=h)
    ops.gemm_nt(h, W2, EPI_BF16, bias=b, out=h2)
print("gemm pair (SiLU, BF16) per GEMM us:", chain(50, two) / 2)
mod = (torch.randn(R, 3 * w, device=dev) * 0.1).bfloat16()
gam, bet = torch.ones(w, device=dev), torch.zeros(w, device=dev)
def ln():
    ops.mar_ln_fwd(x32, gamma=gam, beta=bet, eps=1e-6, mod=mod, shift_off=0, scale_off=w, want_stats=False)
print("mar_ln_fwd per launch us:", chain(100, ln))
def three():
    _, u16, _ = ops.mar_ln_fwd(x32, gamma=gam, beta=bet, eps=1e-6, mod=mod, shift_off=0, scale_off=w, want_stats=False)
    ops.gemm_nt(u16, W1, EPI_SILU, bias=b, out=h)
    ops.gemm_nt(h, W2, EPI_BF16, bias=b, out=h2)
print("block (ln, g1, g2) us:", chain(30, three))
