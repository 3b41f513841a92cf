"""What plain streaming kernels reach on this box for the byte mixes of the HBM-bound stages (development aid):
    gpurun --timeout 120 -- 'python tools/ubench/stream.py'
A 512 MB write between launches keeps L2 cold, as in tools/kbench.py."""
import torch

N, C = 40960, 256
dev = "cuda"
x32 = torch.randn(N, C, device=dev)
y32 = torch.randn(N, C, device=dev)
z32 = torch.empty(N, C, device=dev)
a16 = torch.randn(N, C, device=dev).bfloat16()
b16 = torch.empty(N, 3 * C, device=dev, dtype=torch.bfloat16)
w16 = torch.randn(N, 3 * C, device=dev).bfloat16()
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)

cases = {
    "copy fp32 42+42 MB": (lambda: z32.copy_(x32), 2 * N * C * 4),
    "add fp32 84 in + 42 out MB (residual GEMM mix without A)": (lambda: torch.add(x32, y32, out=z32), 3 * N * C * 4),
    "fill 63 MB bf16 (pure write)": (lambda: b16.fill_(1.0), N * 3 * C * 2),
    "copy bf16 63+63 MB": (lambda: b16.copy_(w16), 2 * N * 3 * C * 2),
    "cast fp32->bf16 42 in + 21 out": (lambda: a16.copy_(x32), N * C * 6),
}
for name, (fn, nbytes) in cases.items():
    for _ in range(3):
        fn()
    ts = []
    for _ in range(20):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    t = ts[len(ts) // 2]
    print(f"{name:60s} {t:7.1f} us  {nbytes / t / 1e6:7.2f} TB/s")
