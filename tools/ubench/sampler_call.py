"""Device time of one diffusion-sampler call (100 ancestral steps) for a given row count: the persistent kernel
(csrc/mar_sampler.cu) against the kernel-by-kernel loop, both replayed from a CUDA graph (development aid):
    gpurun --timeout 200 -- 'python tools/ubench/sampler_call.py [rows ...]'"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hma_b200.mar import STMAR, DiffusionGenieConfig  # noqa: E402

dev = torch.device("cuda")
cfg = DiffusionGenieConfig(num_layers=1, num_heads=8, d_model=256, T=4, S=256, num_factored_vocabs=2, use_mup=False, qkv_bias=True,
                           proj_bias=True, qk_norm=False, mlp_bias=False, mlp_drop=0.0, attn_drop=0.1, patch_size=2,
                           action_network="concat+modulate")
torch.manual_seed(0)
with torch.device(dev):
    m = STMAR(cfg)
    m.init_action_projectors(["a"], [7], [[[0.0] * 7, [1.0] * 7]], "concat+modulate")
with torch.no_grad():
    for p_ in m.parameters():
        if p_.dim() >= 2:
            p_.normal_(0.0, 0.03)
m.eval()
eng, p = m._engine, m._inference_params()
eng.prepare_diffloss(p, False)
te_tab = eng.time_table(p, cfg.num_sampling_steps, dev)
for n in [int(a) for a in sys.argv[1:]] or [64, 512, 4096]:
    z16 = torch.randn(n, 256, device=dev).bfloat16()
    x0 = torch.randn(n, 16, device=dev)
    noise = torch.randn(100, n, 16, device=dev)
    res = {}
    for persistent in (False, True):
        eng.persistent_sampler = persistent
        eng.PERSISTENT_MAX_ROWS = 1 << 30
        eng.sample(p, z16, x0, noise, te_tab, cfg.num_sampling_steps, 1.0, True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = eng.sample(p, z16, x0, noise, te_tab, cfg.num_sampling_steps, 1.0, True)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[persistent] = e0.elapsed_time(e1) / 3
    print(f"rows {n:5d}: kernel-by-kernel {res[False]:7.2f} ms  persistent {res[True]:7.2f} ms  per step {res[False] * 10:6.1f} / {res[True] * 10:6.1f} us")
