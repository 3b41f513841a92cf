#!/usr/bin/env python
"""Interactive single-frame latency (the loop of sim/simulator.py:233-372): B=1, a sliding window of
`--horizon` past frames, one maskgit_generate() per step with `--steps-k` MaskGIT iterations, through the public API.
    python tools/bench_interactive.py [--horizon 8] [--steps-k 2] [--iters 30] [--no-graphs]"""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import D_ACTION_CYCLE, ACTION_DIM_CYCLE  # noqa: E402

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--horizon", type=int, default=8)
    ap.add_argument("--steps-k", type=int, default=2)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--algorithm", default="incremental", choices=["incremental", "full"])
    args = ap.parse_args()
    from hma_b200 import GenieConfig, STMaskGIT
    dev = torch.device("cuda", 0)
    T, S, P = 16, 256, args.horizon
    domains = ["dom00", "dom01"]
    d_actions = D_ACTION_CYCLE[:2]
    stats = [[[0.0] * a, [1.0] * a] for a in ACTION_DIM_CYCLE[:2]]
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
    g = torch.Generator().manual_seed(1)
    frames = torch.randint(0, 262144, (P, 16, 16), generator=g).to(dev)     # cached_latent_frames
    actions = torch.randn(P, d_actions[1], generator=g).to(dev)             # cached_actions
    lat = []
    for it in range(args.iters + 5):
        a = torch.randn(1, d_actions[1], generator=g).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        window = torch.cat([frames, torch.zeros_like(frames[:1])]).unsqueeze(0)[:, : P + 1].long().contiguous()
        window[:, -1] = model.mask_token_id
        acts = torch.cat([actions, a, a]).view(1, -1, a.shape[-1])[:, : P + 1].float().contiguous()
        nxt = model.maskgit_generate(window, out_t=P, maskgit_steps=args.steps_k, temperature=1.0, action_ids=acts,
                                     domain=[domains[1]])[0].squeeze(0)
        nxt_host = nxt.cpu()  # the simulator decodes the tokens to pixels next: the step ends with the tokens on the host
        t1 = time.perf_counter()
        frames = torch.cat([frames[1:], nxt.unsqueeze(0)])
        actions = torch.cat([actions[1:], a])
        if it >= 5:
            lat.append((t1 - t0) * 1e3)
    lat.sort()
    print(json.dumps({"metric": "interactive_step_latency_ms", "median": lat[len(lat) // 2], "p90": lat[int(len(lat) * 0.9)],
                      "frames_per_s": 1e3 / lat[len(lat) // 2],
                      "config": {"batch": 1, "prompt_horizon": P, "maskgit_steps": args.steps_k, "layers": args.layers,
                                 "algorithm": args.algorithm, "cuda_graphs": not args.no_graphs}}))

if __name__ == "__main__":
    main()
