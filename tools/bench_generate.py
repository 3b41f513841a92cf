#!/usr/bin/env python
"""MaskGIT generation throughput (BASELINE.json configs[2]): HMA-MagVit 32L, 8 prompt frames -> 8
generated frames, 16x16 tokens, batch 64 on one GPU, maskgit_steps K, through the public
`STMaskGIT.generate` API.   python tools/bench_generate.py [--batch 64] [--steps-k 2] [--reps 3]"""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import D_ACTION_CYCLE, ACTION_DIM_CYCLE  # noqa: E402

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps-k", type=int, default=2)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--domains", type=int, default=4)
    ap.add_argument("--temperature", type=float, default=1.0)
    ap.add_argument("--algorithm", default="incremental", choices=["incremental", "full"])
    ap.add_argument("--no-graphs", action="store_true")
    args = ap.parse_args()
    from hma_b200 import GenieConfig, STMaskGIT
    dev = torch.device("cuda", 0)
    T, S, Tp = 16, 256, 8
    domains = [f"dom{i:02d}" for i in range(args.domains)]
    d_actions = [D_ACTION_CYCLE[i % 10] for i in range(args.domains)]
    stats = [[[0.0] * a, [1.0] * a] for a in [ACTION_DIM_CYCLE[i % 10] for i in range(args.domains)]]
    cfg = GenieConfig(num_layers=args.layers, num_heads=8, d_model=256, T=T, S=S, num_factored_vocabs=2, qk_norm=False,
                      action_network="concat+modulate")
    torch.manual_seed(0)
    with torch.device(dev):
        model = STMaskGIT(cfg)
        model.init_action_projectors(domains, d_actions, stats, "concat+modulate")
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.02)
    model.eval()
    model.decode_algorithm = args.algorithm
    model.decode_cuda_graphs = not args.no_graphs
    g = torch.Generator().manual_seed(1234)
    prompt = torch.randint(0, 262144, (args.batch, Tp * S), generator=g).to(dev)
    actions = torch.randn(args.batch, T, d_actions[1], generator=g).to(dev)
    dom = [domains[1]] * args.batch
    def run():
        return model.generate(prompt, None, (T - Tp) * S, maskgit_steps=args.steps_k, temperature=args.temperature,
                              action_ids=actions, domain=dom, h=[16], w=[16])
    run(); run(); torch.cuda.synchronize()  # eager warm-up pass, then graph capture
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        toks = run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    frames = args.batch * (T - Tp)
    print(json.dumps({"metric": "maskgit_generated_frames_per_s", "value": frames / (ms / 1e3), "unit": "frames/s",
                      "ms_per_generate_call": ms, "config": {"workload": "HMA-MagVit 32L generate: 8 prompt -> 8 new frames, "
                      "16x16 tokens", "batch": args.batch, "maskgit_steps": args.steps_k, "temperature": args.temperature,
                      "layers": args.layers, "algorithm": ("frame-incremental, temporal K/V cache" + ("" if args.no_graphs else " + CUDA graphs"))
                      if args.algorithm == "incremental" else "full-window recompute (reference algorithm)"},
                      "all_unmasked": bool((toks != 262144).all().item())}))

if __name__ == "__main__":
    main()
