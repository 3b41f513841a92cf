"""Per-stage CUDA-event breakdown of one generate() call (config 3: batch 64, 8 -> 8 frames, K = 2), launched eagerly
(no graph replay) so that every stage launch carries its own events. Development aid:
    gpurun --timeout 300 -- 'python tools/gen_profile.py [batch]'"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hma_b200 import GenieConfig, STMaskGIT, ops  # noqa: E402

Bg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T, S, Tp, K = 16, 256, 8, 2
dev = torch.device("cuda")
cfg = GenieConfig(num_layers=32, num_heads=8, d_model=256, T=T, S=S, num_factored_vocabs=2, qk_norm=False, qkv_bias=False,
                  use_mup=False, action_network="concat+modulate")
torch.manual_seed(0)
with torch.device(dev):
    model = STMaskGIT(cfg)
    model.init_action_projectors(["a", "b"], [7, 14], [[[0.0] * 7, [1.0] * 7], [[0.0] * 14, [1.0] * 14]], "concat+modulate")
with torch.no_grad():
    for p in model.parameters():
        if p.dim() >= 2:
            p.normal_(0.0, 0.02)
model.eval()
model.decode_cuda_graphs = False
g = torch.Generator().manual_seed(1)
prompt = torch.randint(0, 262144, (Bg, Tp * S), generator=g).to(dev)
actions = torch.randn(Bg, T, 14, generator=g).to(dev)


def run():
    return model.generate(prompt, None, (T - Tp) * S, maskgit_steps=K, temperature=1.0, action_ids=actions, domain=["b"] * Bg,
                          h=[16], w=[16])


run()
torch.cuda.synchronize()
ops.PROFILER = ops.Profiler()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run()
e1.record()
summ = ops.PROFILER.summary()
ops.PROFILER = None
tot = sum(v["robust_ms"] for v in summ.values())
print(f"batch {Bg}: eager generate call {e0.elapsed_time(e1):.1f} ms; sum of stage medians x launches {tot:.1f} ms")
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["robust_ms"])[:24]:
    print(f"{k:38s} n={v['launches']:5d} median {v['robust_ms'] / v['launches'] * 1e3:7.1f} us  total {v['robust_ms']:7.2f} ms  {100 * v['robust_ms'] / tot:5.1f}%")
