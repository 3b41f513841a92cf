#!/usr/bin/env python
"""Headline benchmark: HMA-MagVit (32 layers, d=256, 8 heads, 40 action domains) training step,
16 frames x 16x16 tokens (+64 action tokens per frame), batch 8 per GPU — BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = H2D of the batch (e2e leg only), bf16 weight cast, forward, fused factorised CE, backward,
gradient exchange (N > 1), global-norm clip and AdamW. `value` is video tokens/s over all ranks with
the batch already resident in HBM; `e2e` is the same step fed from pinned host buffers with the loss
read back every step. One JSON line is printed by rank 0.

`--impl reference` times the reference's own algorithm on the host cores: the CPU oracle
(oracle/stmaskgit_oracle.py, a restatement pinned against the real reference) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_DOMAINS = 40
D_ACTION_CYCLE = [2, 4, 7, 14, 35, 24, 14, 30, 70, 10]
ACTION_DIM_CYCLE = [2, 4, 7, 7, 7, 8, 14, 2, 7, 10]
L, HEADS, D_MODEL, T, S, B_PER_GPU = 32, 8, 256, 16, 256, 8
N_TOK_FRAME = S + 64
WORKLOAD = ("HMA-MagVit (magvit_n32_h8_d256_action, 40 domains) training step: 16 frames x 16x16 tokens + 64 action "
            "tokens/frame, batch 8/GPU; fwd + fused CE + bwd + grad exchange + clip + AdamW")


def train_flops_per_sample() -> float:
    """fwd = T*n*L*(34 d^2 + 4 n d + 4 d (T+1)/2) + 6 d^2 T L + T*S*2*d*1024; train = 3x."""
    n, d = N_TOK_FRAME, D_MODEL
    fwd = T * n * L * (34 * d * d + 4 * n * d + 4 * d * (T + 1) / 2) + 6 * d * d * T * L + T * S * 2 * d * 1024
    return 3.0 * fwd


def synthetic_batch(gen: torch.Generator, batch: int, d_action: int):
    """Collator distribution (data.py:42-83): per (sample, frame >= 1) mask rate cos(pi/2 * U)."""
    labels = torch.randint(0, 262144, (batch, T * S), generator=gen)
    rate = torch.cos(math.pi / 2 * torch.rand(batch, T, 1, generator=gen))
    rate[:, 0] = 0.0
    mask = torch.rand(batch, T, S, generator=gen) < rate
    ids = torch.where(mask, torch.full_like(labels.view(batch, T, S), 262144), labels.view(batch, T, S)).view(batch, T * S)
    actions = torch.randn(batch, T, d_action, generator=gen)
    return ids, labels, actions


def mar_leg(dev, world: int, rank: int, sync_all, layers: int, steps: int = 6, warmup: int = 3):
    """BASELINE.json configs[3]: HMA-MAR (hma/configs/mar_n32_h8_d256_action.json, 30 action domains, ~1 B parameters):
    12 frames x 16x16 latents of 4 channels (64 patch tokens + 64 action tokens per frame), batch 8 per GPU.
    (i) training step (forward, diffusion loss, backward, gradient exchange, clip, AdamW) fed from pinned host memory,
    (ii) generate(): 6 prompt frames -> 2 new frames, maskgit_steps 16, 100-step ancestral sampler (the reference algorithm:
    a full-window trunk pass per MaskGIT step). Device-event timed, max over ranks."""
    import torch.distributed as dist
    from hma_b200 import ops
    from hma_b200.mar import STMAR, DiffusionGenieConfig, MarTrainStep

    nd, Tm, Bm, Hh = 30, 12, 8, 16
    domains = [f"dom{i:02d}" for i in range(nd)]
    d_actions = [D_ACTION_CYCLE[i % 10] for i in range(nd)]
    stats = [[[0.0] * a, [1.0] * a] for a in (ACTION_DIM_CYCLE[i % 10] for i in range(nd))]
    cfg = DiffusionGenieConfig(num_layers=layers, num_heads=HEADS, d_model=D_MODEL, T=Tm, S=256, num_factored_vocabs=2,
                               use_mup=False, qkv_bias=True, proj_bias=True, qk_norm=False, mlp_bias=False, mlp_drop=0.05,
                               attn_drop=0.1, patch_size=2, action_network="concat+modulate")
    torch.manual_seed(0)
    with torch.device(dev):
        model = STMAR(cfg)
        model.init_action_projectors(domains, d_actions, stats, "concat+modulate")
    n_params = sum(p.numel() for p in model.parameters())
    step_fn = MarTrainStep(model, lr=2e-4, weight_decay=0.01, max_grad_norm=10.0, cuda_graphs=True)  # run_30datasets_mar_waction.sh
    gen = torch.Generator().manual_seed(4321 + rank)
    total = warmup + steps
    host, sched = [], []
    for i in range(total):
        di = (rank + i) % 2  # two domains alternate (each needs one graph capture; training revisits them for ever)
        lat = (torch.randn(Bm, Tm * Hh * Hh, 4, generator=gen) * 0.18215 * 5).pin_memory()
        rate = torch.cos(math.pi / 2 * torch.rand(Bm, Tm, 1, 1, generator=gen))
        rate[:, 0] = 0.0
        mask = (torch.rand(Bm, Tm, Hh, Hh, generator=gen) < rate).pin_memory()
        host.append((lat, mask, torch.randn(Bm, Tm, d_actions[di], generator=gen).pin_memory(), domains[di]))
        sched.append([domains[(r + i) % 2] for r in range(world)])
    for i in range(2):
        lat, mask, act, dname = host[i]
        step_fn.precapture(lat.to(dev), lat.to(dev), act.to(dev), [dname] * Bm, mask.to(dev))
    sync_all()

    def one(i):
        lat, mask, act, dname = host[i]
        x = lat.to(dev, non_blocking=True)
        return step_fn(x, x.clone(), act.to(dev, non_blocking=True), [dname] * Bm, mask.to(dev, non_blocking=True),
                       rank_domains=sched[i])

    for i in range(warmup):
        one(i)
    sync_all()
    n0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warmup, total):
        loss = one(i).cpu()
    e1.record()
    sync_all()
    ms_train = e0.elapsed_time(e1)
    launches = ops.LAUNCHES - n0
    # generation: 6 prompt frames -> 2 new ones
    model.eval()
    Tp, Tn = 6, 2
    prompt = (torch.randn(Bm, Tp * Hh * Hh, 4, generator=gen) * 0.9).pin_memory()
    acts = torch.randn(Bm, Tm, d_actions[0], generator=gen).pin_memory()

    def gen_once(pr=prompt, ac=acts):
        nb = pr.shape[0]
        return model.generate(pr.to(dev, non_blocking=True), None, Tn * Hh * Hh, temperature=1.0,
                              action_ids=ac.to(dev, non_blocking=True), domain=[domains[0]] * nb, h=[Hh], w=[Hh]).cpu()

    gen_once()  # captures the sampler graphs (one per MaskGIT-step row count)
    sync_all()
    e0.record()
    out = gen_once()
    e1.record()
    sync_all()
    ms_gen = e0.elapsed_time(e1)
    # the same call at batch 64 per GPU (the batch of the MaskGIT generation leg): the sampler's 2 000 launches per
    # MaskGIT step are latency-bound at 512 rows, so throughput grows almost linearly with the batch
    Bg = 64
    prompt64 = (torch.randn(Bg, Tp * Hh * Hh, 4, generator=gen) * 0.9).pin_memory()
    acts64 = torch.randn(Bg, Tm, d_actions[0], generator=gen).pin_memory()
    gen_once(prompt64, acts64)
    sync_all()
    e0.record()
    out64 = gen_once(prompt64, acts64)
    e1.record()
    sync_all()
    ms_gen64 = e0.elapsed_time(e1)
    times = torch.tensor([ms_train, ms_gen, ms_gen64], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_train, ms_gen, ms_gen64 = times.tolist()
    del step_fn, model
    torch.cuda.empty_cache()
    return {
        "workload": "HMA-MAR (mar_n32_h8_d256_action, 30 domains): 12 frames x 16x16x4 latents, patch 2 -> 64 + 64 action tokens "
                    "per frame, batch 8/GPU, diffusion-MLP head (depth 4, width 1024), mlp_drop 0.05",
        "layers": layers, "params": n_params,
        "train": {"metric": "train_latent_tokens_per_s", "value": world * Bm * Tm * 256 * steps / (ms_train / 1e3), "unit": "tokens/s",
                  "ms_per_step": ms_train / steps, "steps": steps, "warmup": warmup, "gpu_launches": launches,
                  "loss": float(loss), "io": "latents/mask/actions from pinned host memory, loss read back, every step",
                  "cuda_graph": "forward+loss+backward replayed per action domain; dropout seed on the device"},
        "generate": {"metric": "mar_generated_frames_per_s", "value": world * Bm * Tn / (ms_gen / 1e3), "unit": "frames/s",
                     "ms_per_generate_call": ms_gen, "batch_per_gpu": Bm, "prompt_frames": Tp, "new_frames": Tn,
                     "maskgit_steps": 16, "num_sampling_steps": 100, "finite": bool(torch.isfinite(out).all()),
                     "algorithm": "frame-incremental decode (context frames prefilled once per frame); adaLN modulations of all "
                                  "100 sampler steps in one GEMM; the 100-step ancestral sampler of each MaskGIT step is one "
                                  "CUDA-graph replay",
                     "batch_64": {"value": world * Bg * Tn / (ms_gen64 / 1e3), "unit": "frames/s", "ms_per_generate_call": ms_gen64,
                                  "batch_per_gpu": Bg, "finite": bool(torch.isfinite(out64).all())}},
    }


class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 7 and r[0].isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) >= 7 and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        smax = max([int(r[1]) for r in self.rows if len(r) >= 7 and r[1].isdigit()], default=0)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax or None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU legs (oracle): cpu_baseline of the default run, and the whole `--impl reference` arm
# --------------------------------------------------------------------------------------------------
def cpu_oracle_train_tokens_per_s(steps: int, warmup: int, layers: int = L, batch: int = 1):
    """fwd + bwd of the oracle (reference algorithm, fp32, math attention) on `batch` samples of the
    config-2 shape, all host threads. Returns (tokens/s, seconds per step, threads)."""
    from oracle import stmaskgit_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = O.OracleConfig(num_layers=layers, num_heads=HEADS, d_model=D_MODEL, T=T, S=S, num_factored_vocabs=2,
                         qk_norm=False, action_network="concat+modulate")
    sd = O.make_state_dict(cfg, ["dom00"], [D_ACTION_CYCLE[0]], seed=0, action_dims=[ACTION_DIM_CYCLE[0]])
    params = {k: v.requires_grad_(v.is_floating_point() and "action_preprocessor" not in k) for k, v in sd.items()}
    gen = torch.Generator().manual_seed(1234)
    ids, labels, actions = synthetic_batch(gen, batch, D_ACTION_CYCLE[0])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _, _ = O.forward(ids, labels, actions, ["dom00"] * batch, params, cfg)
        loss.backward()
        for v in params.values():
            v.grad = None
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch * T * S / sec, sec, threads


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    tps, sec, threads = cpu_oracle_train_tokens_per_s(steps, 1)
    line = {
        "impl": "reference", "metric": "train_video_tokens_per_s", "value": tps, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "layers": L, "parallelism": "host cores of rank 0",
                   "sample": "bounded sample of that workload: 1 sample (4096 video tokens) per step, full 32 layers, fwd + loss + "
                             "bwd of the reference algorithm (fp32 oracle port, math attention), 1 action domain"},
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port",
                         "sample": "1 sample (4096 video tokens) of the config-2 shape per step, full 32 layers, fwd+bwd"},
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def decode_flops_per_sample(Tp: int, Tn: int, K: int) -> float:
    """FLOPs the frame-incremental decode actually executes for one sample of generate(Tp prompt -> Tn new frames, K MaskGIT
    steps): one prefill of the Tp context frames (no head), then per new frame K one-frame passes with the head and (except
    for the last frame) one "commit" pass without it; a one-frame pass at window position t attends t+1 cached frames."""
    n, d = N_TOK_FRAME, D_MODEL
    per_tok = 34 * d * d + 4 * n * d
    prefill = Tp * n * L * (per_tok + 4 * d * (Tp + 1) / 2) + 6 * d * d * Tp * L
    head = S * 2 * d * 1024
    total = prefill
    for t in range(Tp, Tp + Tn):
        one = n * L * (per_tok + 4 * d * (t + 1)) + 6 * d * d * L
        total += K * (one + head) + (one if t != Tp + Tn - 1 else 0.0)
    return total


def interactive_leg(model, dev, domains, d_actions, horizon: int = 8, K: int = 2, iters: int = 20):
    """SURVEY.md §8(f) rank 3, the loop of sim/simulator.py:233-372: B=1, a sliding window of `horizon` past frames, one
    maskgit_generate() per simulator step with the new action, tokens read back to the host (the simulator decodes them to
    pixels next). Wall-clock latency per step through the public API (host-synchronous by construction)."""
    model.eval()
    model._sessions.clear()
    g = torch.Generator().manual_seed(1)
    P = horizon
    frames = torch.randint(0, 262144, (P, 16, 16), generator=g).to(dev)
    actions = torch.randn(P, d_actions[1], generator=g).to(dev)
    lat = []
    for it in range(iters + 5):
        a = torch.randn(1, d_actions[1], generator=g).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        window = torch.cat([frames, torch.zeros_like(frames[:1])]).unsqueeze(0)[:, : P + 1].long().contiguous()
        window[:, -1] = model.mask_token_id
        acts = torch.cat([actions, a, a]).view(1, -1, a.shape[-1])[:, : P + 1].float().contiguous()
        nxt = model.maskgit_generate(window, out_t=P, maskgit_steps=K, temperature=1.0, action_ids=acts,
                                     domain=[domains[1]])[0].squeeze(0)
        nxt.cpu()
        t1 = time.perf_counter()
        frames = torch.cat([frames[1:], nxt.unsqueeze(0)])
        actions = torch.cat([actions[1:], a])
        if it >= 5:
            lat.append((t1 - t0) * 1e3)
    lat.sort()
    model._sessions.clear()
    model.train()
    return {"metric": "interactive_step_latency_ms", "median_ms": lat[len(lat) // 2], "p90_ms": lat[int(len(lat) * 0.9)],
            "frames_per_s": 1e3 / lat[len(lat) // 2], "higher_is_better": False,
            "config": {"batch": 1, "prompt_horizon": P, "maskgit_steps": K, "layers": model.config.num_layers,
                       "loop": "sim/simulator.py:233-372 shape: slide the window, maskgit_generate(out_t=horizon), tokens to host",
                       "algorithm": "per-session temporal K/V cache, CUDA-graph replay of prefill and one-frame passes"}}


def pipeline_leg(dev, cfg_T: int = T, batch: int = B_PER_GPU, batches: int = 40):
    """feed rate of the on-device data path (MultiTaskBatchSampler -> index gather from the HBM-resident
    token / action tables -> on-device MaskGIT collator), in samples/s, next to the training step's consumption rate.
    Synthetic datasets in the reference's on-disk format (tests/_rawdata.py layout) written to a temp directory."""
    import tempfile

    import numpy as np

    from hma_b200 import GenieConfig
    from hma_b200.dataset import RawTokenDataset
    from hma_b200.sampler import DeviceBatchPipeline
    with tempfile.TemporaryDirectory() as tmp:
        dsets = []
        for j, (n_img, adim) in enumerate(((20000, 7), (12000, 14))):
            root = os.path.join(tmp, f"ds{j}")
            os.makedirs(os.path.join(root, "actions"))
            rng = np.random.default_rng(j)
            arrs = {"video.bin": rng.integers(0, 2 ** 18, size=(n_img, 16, 16)).astype(np.uint32),
                    "segment_ids.bin": (np.arange(n_img) // 200).astype(np.int32),
                    "actions/actions.bin": rng.normal(size=(n_img, adim)).astype(np.float32)}
            for name, arr in arrs.items():
                arr.tofile(os.path.join(root, name))
            with open(os.path.join(root, "metadata.json"), "w") as fh:
                json.dump({"token_dtype": "uint32", "action_dim": adim, "s": 16, "h": 16, "w": 16, "vocab_size": 2 ** 18, "hz": 2,
                           "num_images": n_img, "name": f"bench_robot_{j}"}, fh)
            dsets.append(RawTokenDataset(root, window_size=cfg_T, use_actions=True, freq_table={}).to_device(dev))
        cfg = GenieConfig(num_layers=1, num_heads=HEADS, d_model=D_MODEL, T=cfg_T, S=S, num_factored_vocabs=2,
                          dataloader_apply_corruption=True, non_mlm_ratio=0.5)
        pipe = DeviceBatchPipeline(dsets, cfg, batch_size=batch, seed=0)
        it = iter(pipe)
        for _ in range(5):
            next(it)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(batches):
            b = next(it)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        return {"metric": "device_pipeline_samples_per_s", "value": batches * batch / wall, "unit": "samples/s",
                "ms_per_batch_wall": wall / batches * 1e3, "ms_per_batch_device": e0.elapsed_time(e1) / batches,
                "config": {"batch": batch, "T": cfg_T, "datasets": 2, "corruption": True, "non_mlm_ratio": 0.5,
                           "path": "sampler (host index draw) -> hma_gather_token_windows / hma_gather_rows_f32 -> "
                                   "torch device RNG draws -> hma_collate_maskgit"}}


def config5_leg(dev, world: int, rank: int, sync_all, layers: int, steps: int = 4, warmup: int = 3):
    """BASELINE.json configs[4]: long context — 32 frames x 16x16 tokens (+64 action tokens per frame), batch 8 per GPU, batches
    drawn from a 40-domain mix with heterogeneous action widths (each rank trains one domain per step, as the reference's
    MultiTaskBatchSampler guarantees). Same TrainStep as the headline leg; device-timed, max over ranks."""
    import torch.distributed as dist
    from hma_b200 import GenieConfig, STMaskGIT
    from hma_b200.train import TrainStep
    T5 = 32
    domains = [f"dom{i:02d}" for i in range(NUM_DOMAINS)]
    d_actions = [D_ACTION_CYCLE[i % 10] for i in range(NUM_DOMAINS)]
    stats = [[[0.0] * a, [1.0] * a] for a in (ACTION_DIM_CYCLE[i % 10] for i in range(NUM_DOMAINS))]
    cfg = GenieConfig(num_layers=layers, num_heads=HEADS, d_model=D_MODEL, T=T5, S=S, num_factored_vocabs=2, qk_norm=False,
                      qkv_bias=False, use_mup=False, action_network="concat+modulate")
    torch.manual_seed(0)
    with torch.device(dev):
        model = STMaskGIT(cfg)
        model.init_action_projectors(domains, d_actions, stats, "concat+modulate")
    with torch.no_grad():
        for k, p in model.named_parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.02)
    step_fn = TrainStep(model, lr=1e-4, weight_decay=0.05, max_grad_norm=1.0, cuda_graphs=True)
    gen = torch.Generator().manual_seed(99 + rank)
    total = warmup + steps
    batches, sched = [], []
    for i in range(total):
        di = (rank + 3 * i) % NUM_DOMAINS
        labels = torch.randint(0, 262144, (B_PER_GPU, T5 * S), generator=gen)
        rate = torch.cos(math.pi / 2 * torch.rand(B_PER_GPU, T5, 1, generator=gen))
        rate[:, 0] = 0.0
        mask = torch.rand(B_PER_GPU, T5, S, generator=gen) < rate
        ids = torch.where(mask.view(B_PER_GPU, -1), torch.full_like(labels, 262144), labels)
        batches.append((ids.to(dev), labels.to(dev), torch.randn(B_PER_GPU, T5, d_actions[di], generator=gen).to(dev), domains[di]))
        sched.append([domains[(r + 3 * i) % NUM_DOMAINS] for r in range(world)])
    seen = set()
    for b in batches:
        if b[3] not in seen:
            seen.add(b[3])
            step_fn.precapture(b[0], b[1], b[2], [b[3]] * B_PER_GPU)
    sync_all()
    for i in range(warmup):
        b = batches[i]
        step_fn(b[0], b[1], b[2], [b[3]] * B_PER_GPU, rank_domains=sched[i])
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warmup, total):
        b = batches[i]
        out = step_fn(b[0], b[1], b[2], [b[3]] * B_PER_GPU, rank_domains=sched[i])
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    loss = float(out[0].item())
    n, d = N_TOK_FRAME, D_MODEL
    fwd = T5 * n * layers * (34 * d * d + 4 * n * d + 4 * d * (T5 + 1) / 2) + 6 * d * d * T5 * layers + T5 * S * 2 * d * 1024
    del step_fn, model
    torch.cuda.empty_cache()
    return {"workload": "HMA-MagVit 32 frames x 16x16 tokens + 64 action tokens/frame, batch 8/GPU, 40 action domains of widths "
                        "2..70 (a different domain every step), training step", "metric": "train_video_tokens_per_s",
            "value": world * B_PER_GPU * T5 * S / (ms / 1e3), "unit": "tokens/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "domains_visited": len(seen), "loss": loss, "model_tflops_per_gpu": 3.0 * fwd * B_PER_GPU / (ms / 1e3) / 1e12}


def decoder_leg(dev, batch: int = 16, reps: int = 5):
    """token -> pixel decode (visualize.py:124-169 / the simulator's display path): 16x16 token grids ->
    256x256 uint8 frames through MagVitDecoder.decode_tokens, host tokens in, host frames out inside the timed region."""
    from hma_b200.tokenizer import MagVitDecoder, VQConfig
    torch.manual_seed(0)
    with torch.device(dev):
        dec = MagVitDecoder(VQConfig())
    g = torch.Generator().manual_seed(5)
    tokens = torch.randint(0, 262144, (batch, 16, 16), generator=g).pin_memory()
    flops = 0.0
    for name, m in dec.named_modules():
        if isinstance(m, torch.nn.Conv2d):
            lvl = int(name.split(".")[1]) if name.startswith("up.") else (0 if name == "conv_out" else 4)
            side = 16 * 2 ** (4 - lvl)
            flops += 2.0 * side * side * m.weight.numel()
    for _ in range(2):
        dec.decode_tokens(tokens.to(dev, non_blocking=True)).cpu()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        frames = dec.decode_tokens(tokens.to(dev, non_blocking=True)).cpu()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    del dec
    torch.cuda.empty_cache()
    return {"metric": "decoded_frames_per_s", "value": batch / (ms / 1e3), "unit": "frames/s", "ms_per_batch": ms, "batch": batch,
            "frame": list(frames.shape[1:]), "gflop_per_frame": flops / 1e9, "achieved_tflops": flops * batch / (ms / 1e3) / 1e12,
            "what": "LFQ code lookup + MagViT2 decoder (40.5 M parameters, 19 3x3 convolutions as tcgen05 contractions with "
                    "TMA-offset taps) + uint8 mapping; random weights (the checkpoint is not available offline)"}


def gpu_reference_leg(dev, steps: int = 3, warmup: int = 1):
    """SURVEY.md §8(d) / BASELINE.md §5, "the bar to beat on the same box": the reference's algorithm in plain PyTorch on the
    SAME B200 under torch.autocast(bf16) — cuBLAS GEMMs, ATen element-wise kernels and (i) the reference's own math attention
    (BasicSelfAttention, attention.py:37-61), (ii) fused attention through scaled_dot_product_attention (what the reference
    gets from xformers, attention.py:139-155). It is the oracle module (a restatement pinned on the real reference), never
    part of the product path. Training: fwd + loss + bwd of the config-2 batch (no optimizer: flattering to the reference);
    generation: the reference algorithm (full-window recompute per MaskGIT step) at batch 64."""
    from oracle import stmaskgit_oracle as O
    cfg = O.OracleConfig(num_layers=L, num_heads=HEADS, d_model=D_MODEL, T=T, S=S, num_factored_vocabs=2, qk_norm=False,
                         action_network="concat+modulate")
    sd = O.make_state_dict(cfg, ["dom00"], [D_ACTION_CYCLE[0]], seed=0, std=0.02, action_dims=[ACTION_DIM_CYCLE[0]])
    params = {k: v.to(dev).requires_grad_(v.is_floating_point() and "action_preprocessor" not in k) for k, v in sd.items()}
    gen = torch.Generator().manual_seed(1234)
    ids, labels, actions = (t.to(dev) for t in synthetic_batch(gen, B_PER_GPU, D_ACTION_CYCLE[0]))
    dom = ["dom00"] * B_PER_GPU
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for impl in ("math", "sdpa"):
        O.ATTENTION_IMPL = impl
        try:
            for i in range(warmup + steps):
                if i == warmup:
                    torch.cuda.synchronize()
                    e0.record()
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss, _, _ = O.forward(ids, labels, actions, dom, params, cfg)
                loss.backward()
                for v in params.values():
                    v.grad = None
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[f"train_{impl}"] = {"value": B_PER_GPU * T * S / (ms / 1e3), "unit": "tokens/s", "ms_per_step": ms,
                                    "loss": float(loss.detach())}
        finally:
            O.ATTENTION_IMPL = "math"
        torch.cuda.empty_cache()
    # generation, reference algorithm, batch 64 (fused attention: the faster of the two above)
    Tp, Tn, K, Bg = 8, T - 8, 2, 64
    prompt = torch.randint(0, 262144, (Bg, Tp * S), generator=gen).to(dev)
    acts = torch.randn(Bg, T, D_ACTION_CYCLE[0], generator=gen).to(dev)
    nograd = {k: v.detach() for k, v in params.items()}
    O.ATTENTION_IMPL = "sdpa"
    try:
        for i in range(2):
            if i == 1:
                torch.cuda.synchronize()
                e0.record()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                toks, _ = O.generate(prompt, Tn * S, nograd, cfg, 16, 16, maskgit_steps=K, temperature=1.0, action_ids=acts,
                                     domain=["dom00"] * Bg)
            toks.cpu()
        e1.record()
        torch.cuda.synchronize()
    finally:
        O.ATTENTION_IMPL = "math"
    ms = e0.elapsed_time(e1)
    out["generate_sdpa"] = {"value": Bg * Tn / (ms / 1e3), "unit": "frames/s", "ms_per_generate_call": ms, "batch": Bg,
                            "maskgit_steps": K}
    out["what"] = ("the reference's algorithm as plain PyTorch (oracle module) on this GPU under torch.autocast(bf16): 32 layers, "
                   "B=8, T=16 fwd+loss+bwd without optimizer step; math = BasicSelfAttention, sdpa = fused attention; generation = "
                   "full-window recompute per MaskGIT step at batch 64")
    del params, nograd
    torch.cuda.empty_cache()
    return out


def generation_leg(model, dev, world, rank, domains, d_actions, sync_all, reps: int = 3, per_gpu: bool = True):
    """BASELINE configs[2]: 8 prompt frames -> 8 generated frames, 16x16 tokens, batch 64 split over the GPUs (replicas,
    no collective), maskgit_steps 2, temperature 1, through the public STMaskGIT.generate API. Host prompt in, host
    tokens out inside the timed region. Returns (frames/s over all ranks, ms per generate call)."""
    import torch.distributed as dist
    Tp, Tn, K = 8, T - 8, 2
    Bg = 64 if per_gpu else max(1, 64 // world)
    g = torch.Generator().manual_seed(4321 + rank)
    prompt = torch.randint(0, 262144, (Bg, Tp * S), generator=g).pin_memory()
    actions = torch.randn(Bg, T, d_actions[1], generator=g).pin_memory()
    dom = [domains[1]] * Bg
    model.eval()

    def run():
        toks = model.generate(prompt.to(dev, non_blocking=True), None, Tn * S, maskgit_steps=K, temperature=1.0,
                              action_ids=actions.to(dev, non_blocking=True), domain=dom, h=[16], w=[16])
        return toks.cpu()

    run()
    run()  # eager pass, then CUDA-graph capture of the one-frame passes
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        toks = run()
    e1.record()
    sync_all()
    assert bool((toks != 262144).all())
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    model.train()
    return world * Bg * Tn / (ms / 1e3), ms, Bg


# --------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hma_b200", choices=["hma_b200", "reference"])
    ap.add_argument("--layers", type=int, default=L, help=argparse.SUPPRESS)  # debugging only; default = full model
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--breakdown", default=None, help="write a per-stage CUDA-event breakdown JSON here")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from the host instead of replaying CUDA graphs")
    ap.add_argument("--no-generation", action="store_true", help="skip the MaskGIT generation leg (BASELINE configs[2])")
    ap.add_argument("--no-mar", action="store_true", help="skip the HMA-MAR leg (BASELINE configs[3])")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the plain-PyTorch-on-this-GPU reference leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the interactive-latency and data-pipeline legs")
    ap.add_argument("--no-config5", action="store_true", help="skip the long-context (T=32) training leg (BASELINE configs[4])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from hma_b200 import GenieConfig, STMaskGIT, _lib, ops
    from hma_b200.train import TrainStep
    _lib.call("hma_device_check")

    domains = [f"dom{i:02d}" for i in range(NUM_DOMAINS)]
    d_actions = [D_ACTION_CYCLE[i % 10] for i in range(NUM_DOMAINS)]
    adims = [ACTION_DIM_CYCLE[i % 10] for i in range(NUM_DOMAINS)]
    stats = [[[0.0] * a, [1.0] * a] for a in adims]
    cfg = GenieConfig(num_layers=args.layers, num_heads=HEADS, d_model=D_MODEL, T=T, S=S, num_factored_vocabs=2,
                      qk_norm=False, qkv_bias=False, use_mup=False, action_network="concat+modulate")
    torch.manual_seed(0)
    with torch.device(dev):
        model = STMaskGIT(cfg)
        model.init_action_projectors(domains, d_actions, stats, "concat+modulate")
    with torch.no_grad():  # non-degenerate weights (the reference init is ~uniform logits): N(0, 0.02) on matrices
        for k, p in model.named_parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.02)
    n_params = sum(p.numel() for p in model.parameters())
    step_fn = TrainStep(model, lr=1e-4, weight_decay=0.05, max_grad_norm=1.0, cuda_graphs=not args.no_graphs,
                        overlap_segments=int(os.environ.get("HMA_B200_OVERLAP_SEGMENTS", "1")))  # env: A/B measurements only

    total = args.warmup + args.steps
    gen = torch.Generator().manual_seed(1234 + rank)
    host, sched = [], []
    for i in range(2 * total):
        di = (rank + i) % NUM_DOMAINS
        ids, labels, actions = synthetic_batch(gen, B_PER_GPU, d_actions[di])
        host.append((ids.pin_memory(), labels.pin_memory(), actions.pin_memory()))
        sched.append([domains[(r + i) % NUM_DOMAINS] for r in range(world)])
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_resident(i, batch):
        return step_fn(batch[0], batch[1], batch[2], [sched[i][rank]] * B_PER_GPU, rank_domains=sched[i])

    # One CUDA graph per action domain (forward + loss + backward; ~1400 launches), captured here, before any timed
    # region, for the domains the schedule will visit: in training every domain is revisited thousands of times.
    if step_fn.cuda_graphs:
        seen = set()
        for i in range(2 * total):
            dname = sched[i][rank]
            if dname not in seen:
                seen.add(dname)
                b = tuple(t.to(dev) for t in host[i])
                step_fn.precapture(b[0], b[1], b[2], [dname] * B_PER_GPU)
        sync_all()

    # ---------------- leg 1: inputs resident in HBM
    resident = [tuple(t.to(dev) for t in b) for b in host[:total]]
    for i in range(args.warmup):
        run_resident(i, resident[i])
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ops.LAUNCHES
    dominant = "attn_spatial_bwd"  # largest share of the step (profiles/r02a_launches_summary.txt); DESIGN.md §3.2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()  # no-op unless run as `ncu --profile-from-start off ...` (profiles/ recipes)
    e0.record()
    for i in range(args.warmup, total):
        out = run_resident(i, resident[i])
    e1.record()
    sync_all()
    torch.cuda.profiler.stop()
    ms_resident = e0.elapsed_time(e1)
    launches = ops.LAUNCHES - launches0
    loss_val = float(out[0].item())
    # Per-launch CUDA events cannot ride inside graph replays, and between host launches they break the programmatic
    # dependent launch overlap: the roofline kernel (and every other stage) is timed over the SAME steps re-launched
    # from the host with an event pair around each launch, right after the timed region, in this process, with the
    # clock sampler still running.
    roofline_pass = "the same steps re-run with per-launch events right after the timed region"
    was_graph = step_fn.cuda_graphs
    step_fn.cuda_graphs = False
    ops.PROFILER = ops.Profiler()
    for i in range(args.warmup, total):
        run_resident(i, resident[i])
    step_fn.cuda_graphs = was_graph
    prof = ops.PROFILER.summary()
    ops.PROFILER = None

    # ---------------- leg 2: end to end from pinned host buffers, loss read back each step
    del resident
    for i in range(args.warmup):
        b = tuple(t.to(dev, non_blocking=True) for t in host[total + i])
        run_resident(total + i, b)[0].item()
    sync_all()
    e0.record()
    for i in range(args.warmup, total):
        b = tuple(t.to(dev, non_blocking=True) for t in host[total + i])
        run_resident(total + i, b).cpu()
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop()

    times = torch.tensor([ms_resident, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_resident, ms_e2e = times.tolist()
    tokens_per_step = world * B_PER_GPU * T * S
    value = tokens_per_step * args.steps / (ms_resident / 1e3)
    e2e = tokens_per_step * args.steps / (ms_e2e / 1e3)

    # ---------------- MaskGIT generation (the other half of BASELINE.json's metric)
    gen_fps = None
    if not args.no_generation:
        gen_fps, gen_ms, gen_b = generation_leg(model, dev, world, rank, domains, d_actions, sync_all)
        gen_strong = None
        if world > 1:  # the same 64 samples split over the GPUs (strong scaling), for reference
            gen_strong = generation_leg(model, dev, world, rank, domains, d_actions, sync_all, per_gpu=False)

    # ---------------- interactive single-frame loop (B=1) and the on-device data pipeline (rank 0 only: no collective)
    interactive = pipeline = pixel_decode = None
    if not args.no_extras and rank == 0:
        interactive = interactive_leg(model, dev, domains, d_actions)
        pipeline = pipeline_leg(dev)
        pixel_decode = decoder_leg(dev)
    sync_all()

    # ---------------- optional per-stage breakdown (one extra, untimed step)
    if args.breakdown and rank == 0:
        ops.PROFILER = ops.Profiler()
        b = tuple(t.to(dev) for t in host[0])
        run_resident(0, b)
        bd = ops.PROFILER.summary()
        ops.PROFILER = None
        with open(args.breakdown, "w") as f:
            json.dump({k: v for k, v in sorted(bd.items(), key=lambda kv: -kv[1]["total_ms"])}, f, indent=1)

    # ---------------- HMA-MAR (continuous tokens + diffusion head): training step and sampling, BASELINE configs[3]
    graphs_on = step_fn.cuda_graphs
    exchange_segments = step_fn.overlap_segments
    mar = None
    if not args.no_mar:
        del step_fn, run_resident
        model = None
        torch.cuda.empty_cache()
        mar = mar_leg(dev, world, rank, sync_all, args.layers)

    # ---------------- long context, heterogeneous action stems (BASELINE configs[4])
    cfg5 = None
    if not args.no_config5:
        step_fn = run_resident = model = None
        torch.cuda.empty_cache()
        cfg5 = config5_leg(dev, world, rank, sync_all, args.layers)

    # ---------------- the reference algorithm as plain PyTorch on this same GPU (rank 0, N = 1 only)
    gpu_ref = None
    if not args.no_gpu_reference and world == 1:
        step_fn = run_resident = model = None
        torch.cuda.empty_cache()
        gpu_ref = gpu_reference_leg(dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_roofline_kernel.json")) as f:
            traffic = json.load(f).get("traffic_bytes")  # dram read+write of one launch, from the committed ncu capture
    except Exception:
        pass
    rec = prof.get(dominant, {"launches": 0, "total_ms": 0.0, "robust_ms": 0.0, "work": 0.0})
    achieved = rec["work"] / (rec["total_ms"] / 1e3) / 1e12 if rec["total_ms"] else 0.0
    # every stage of the step, for context: share of the summed kernel time and achieved rate on its own bound
    # (work = algorithmic FLOPs for the contractions, algorithmic bytes for the streaming stages; hma_b200/ops.py). Shares and
    # rates use median launch time x launches (robust_ms): in this host-launched pass a kernel's event pair occasionally
    # includes a host stall
    hbm_gbs = peaks.get("hbm_gbs", 6650.0)
    main_stream = {k: v for k, v in prof.items() if v["launches"] and v["work"] / v["launches"] >= 1e8}  # drop the
    # one-tile adaLN chain: it runs on a side stream, where event pairs also time its waits on the main stream
    tot_ms = sum(v["robust_ms"] for v in main_stream.values()) or 1.0
    stages = []
    for kind, v in sorted(main_stream.items(), key=lambda kv: -kv[1]["robust_ms"])[:14]:
        is_flops = kind.startswith(("gemm", "attn_spatial"))
        rate = v["work"] / (v["robust_ms"] / 1e3) if v["robust_ms"] else 0.0
        stages.append({"stage": kind, "share": round(v["robust_ms"] / tot_ms, 4), "launches": v["launches"],
                       "median_us": round(v["robust_ms"] / max(v["launches"], 1) * 1e3, 1),
                       "mean_us": round(v["total_ms"] / max(v["launches"], 1) * 1e3, 1),
                       "achieved": round(rate / 1e12, 1) if is_flops else round(rate / 1e9, 0),
                       "unit": "TFLOP/s" if is_flops else "GB/s",
                       "frac_of_peak": round(rate / 1e12 / peak_tf, 3) if is_flops else round(rate / 1e9 / hbm_gbs, 3)})
    step_tf = (B_PER_GPU * train_flops_per_sample() * (args.layers / L)) / (ms_resident / args.steps / 1e3) / 1e12

    line = {
        "metric": "train_video_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_resident / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "layers": args.layers, "params": n_params,
                   "global_batch": world * B_PER_GPU, "tokens_per_step": tokens_per_step, "parallelism": f"dp{world}",
                   "l2": "per-step working set (~19 GB of activations) >> 126 MB L2; no explicit flush needed",
                   "loss": loss_val, "model_tflops_per_gpu": step_tf,
                   "cuda_graph": ("forward+loss+backward replayed from one CUDA graph per action domain (captured before the timed "
                                  "region; with grad_exchange_segments > 1: one graph per backward segment and the shared-gradient "
                                  "all-reduce of each segment launched between the replays); gradient exchange, clip and AdamW "
                                  "launched from the host")
                   if graphs_on else "off",
                   "grad_exchange_segments": exchange_segments},
        "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "attn_spatial_bwd_kernel (per-frame attention backward, 128 frames x 8 heads x 320 tokens, "
                                                    "head_dim 32: 33.55 algorithmic GFLOP and 150 MB per launch (qkv + dO + lse + delta in, dqkv "
                                                    "out), AI ~ the ridge)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                     "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)", "launches_timed": rec["launches"], "timed_in": roofline_pass, "avg_launch_us": (rec["total_ms"] / max(rec["launches"], 1)) * 1e3,
                     "peak_source": peak_src, "stages": stages},
    }
    if gen_fps is not None:
        line["generation"] = {
            "metric": "maskgit_generated_frames_per_s", "value": gen_fps, "unit": "frames/s", "ms_per_generate_call": gen_ms,
            "config": {"workload": "HMA-MagVit 32L generate(): 8 prompt frames -> 8 generated frames, 16x16 tokens, "
                                   "maskgit_steps 2, temperature 1.0, unmask_mode random", "batch_per_gpu": gen_b,
                       "global_batch": gen_b * world, "layers": args.layers, "parallelism": f"replicas x{world} (no collective)",
                       "algorithm": "frame-incremental decode: per-layer temporal K/V cache + CUDA-graph replay of the "
                                    "one-frame pass (reference algorithm recomputes the 16-frame window per MaskGIT step)",
                       "io": "pinned host prompt/actions in, host tokens out, inside the timed region",
                       "scaling": "weak (batch 64 per GPU)"}}
        gflops = decode_flops_per_sample(8, T - 8, 2) * (args.layers / L) * gen_b
        ach = gflops / (gen_ms / 1e3) / 1e12
        line["generation"]["roofline"] = {
            "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
            "traffic": None, "flops_per_generate_call": gflops,
            "what": "FLOPs the frame-incremental decode executes per generate() call (prefill of 8 frames + per new frame K one-frame "
                    "passes with the head + one commit pass; the first step of a frame shares a launch with the prefill / the "
                    "previous frame's commit, same FLOPs) / device time of the call, H2D prompt and D2H tokens included"}
        if gen_strong is not None:
            line["generation"]["strong_scaling_total_batch_64"] = {"value": gen_strong[0], "unit": "frames/s",
                                                                    "ms_per_generate_call": gen_strong[1], "batch_per_gpu": gen_strong[2]}
    if interactive is not None:
        line["interactive"] = interactive
    if pipeline is not None:
        pipeline["train_step_consumes_samples_per_s"] = B_PER_GPU * args.steps / (ms_resident / 1e3)
        line["pipeline"] = pipeline
    if pixel_decode is not None:
        line["pixel_decode"] = pixel_decode
    if gpu_ref is not None:
        gpu_ref["speedup_train_vs_sdpa"] = value / gpu_ref["train_sdpa"]["value"]
        gpu_ref["speedup_train_vs_math"] = value / gpu_ref["train_math"]["value"]
        if gen_fps is not None:
            gpu_ref["speedup_generate_vs_sdpa"] = gen_fps / gpu_ref["generate_sdpa"]["value"]
        line["gpu_reference"] = gpu_ref
    if cfg5 is not None:
        line["config5_long_context"] = cfg5
    if mar is not None:
        line["mar"] = mar
    if not args.no_cpu_baseline and world == 1:
        tps, sec, threads = cpu_oracle_train_tokens_per_s(2, 1)
        line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port",
                                "sample": "1 sample (4096 video tokens) of the same shape per step, 32 layers, fwd+bwd, "
                                          "fp32 oracle port of the reference; 2 timed steps after 1 warm-up"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
