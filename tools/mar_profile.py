"""One HMA-MAR training step (eager launches) inside a cudaProfilerStart/Stop range, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/mar_profile.py [--layers L]
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hma_b200.mar import STMAR, DiffusionGenieConfig, MarTrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--sample-steps", type=int, default=0, help="also profile this many ancestral sampler steps on 512 rows")
args = ap.parse_args()
dev = torch.device("cuda", 0)
nd, T, B, Hh = 2, 12, 8, 16
domains = [f"dom{i:02d}" for i in range(nd)]
cfg = DiffusionGenieConfig(num_layers=args.layers, num_heads=8, d_model=256, T=T, S=256, num_factored_vocabs=2, use_mup=False,
                           qkv_bias=True, proj_bias=True, qk_norm=False, mlp_bias=False, mlp_drop=0.05, patch_size=2,
                           action_network="concat+modulate")
torch.manual_seed(0)
with torch.device(dev):
    model = STMAR(cfg)
    model.init_action_projectors(domains, [14, 10], [[[0.0] * 7, [1.0] * 7], [[0.0] * 10, [1.0] * 10]], "concat+modulate")
step = MarTrainStep(model, lr=2e-4, weight_decay=0.01, max_grad_norm=10.0, cuda_graphs=False)
g = torch.Generator().manual_seed(0)
lat = (torch.randn(B, T * Hh * Hh, 4, generator=g) * 0.9).to(dev)
rate = torch.cos(math.pi / 2 * torch.rand(B, T, 1, 1, generator=g))
mask = (torch.rand(B, T, Hh, Hh, generator=g) < rate).to(dev)
act = torch.randn(B, T, 14, generator=g).to(dev)
for _ in range(2):
    step(lat.clone(), lat.clone(), act, [domains[0]] * B, mask)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step(lat.clone(), lat.clone(), act, [domains[0]] * B, mask)
if args.sample_steps:
    model.eval()
    eng, p = model._engine, model._inference_params()
    eng.prepare_diffloss(p, False)
    z16 = torch.randn(B * 64, 256, device=dev).bfloat16()
    te = eng.time_table(p, "100", dev)
    tb, _, steps = eng.tables("100", dev)
    c = eng.sample_cond(p, z16)
    x = torch.randn(B * 64, 16, device=dev)
    from hma_b200 import ops
    x16 = ops.mar_q_sample(x, None, None, None, 128)
    nx, nx16 = torch.empty_like(x), torch.empty_like(x16)
    for i in range(steps - 1, steps - 1 - args.sample_steps, -1):
        eng.sample_step(p, c, te, tb, i, x, x16, torch.randn_like(x), 1.0, True, nx, nx16)
        x, nx, x16, nx16 = nx, x, nx16, x16
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
