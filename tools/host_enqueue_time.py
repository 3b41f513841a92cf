"""Host time to enqueue one training step without CUDA graphs vs the GPU time of the step (why TrainStep replays graphs)."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from hma_b200 import GenieConfig, STMaskGIT
from hma_b200.train import TrainStep
dev = torch.device('cuda', 0)
domains = [f"dom{i:02d}" for i in range(4)]
d_actions = [bench.D_ACTION_CYCLE[i % 10] for i in range(4)]
adims = [bench.ACTION_DIM_CYCLE[i % 10] for i in range(4)]
stats = [[[0.0] * a, [1.0] * a] for a in adims]
cfg = GenieConfig(num_layers=32, num_heads=8, d_model=256, T=16, S=256, num_factored_vocabs=2, qk_norm=False, qkv_bias=False, use_mup=False, action_network="concat+modulate")
with torch.device(dev):
    model = STMaskGIT(cfg); model.init_action_projectors(domains, d_actions, stats, "concat+modulate")
with torch.no_grad():
    for p in model.parameters():
        if p.dim() >= 2: p.normal_(0, 0.02)
step = TrainStep(model)
gen = torch.Generator().manual_seed(0)
b = [t.to(dev) for t in bench.synthetic_batch(gen, 8, d_actions[1])]
for _ in range(3): step(b[0], b[1], b[2], [domains[1]] * 8)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8): step(b[0], b[1], b[2], [domains[1]] * 8)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/8:.2f} ms/step, total {1e3*(t2-t0)/8:.2f} ms/step")
