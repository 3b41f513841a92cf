"""Training step over the CUDA engine without autograd in the loop: forward, fused loss, backward,
gradient synchronisation over NCCL, global-norm clip and AdamW — the device-side content of one
iteration of the reference trainer (train_multi.py:556-598).

Parameters live in one flat fp32 arena laid out [shared | domain 0 | domain 1 | ...]; the model's
nn.Parameters are views into it, so state_dict()/load_state_dict() keep the reference key layout.
A backward writes its gradients as one flat buffer [shared | active domain] with the same intra-range
layout, which makes the optimizer two streaming kernels and the gradient exchange two collectives:
  * all-reduce of the shared range (every rank contributes),
  * all-gather of the per-domain range (each rank trains ONE domain per batch, as the reference's
    MultiTaskBatchSampler guarantees; the reference instead all-reduces all ~375 M parameters, of
    which ~330 M are zeros — SURVEY.md §2.2, §8e).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, ops
from .engine import SMOOTHING, Engine


class ParamArena:
    def __init__(self, model):
        eng: Engine = model._engine
        cfg = model.config
        named = dict(model.named_parameters())
        dev = next(iter(named.values())).device
        d = eng.dims(1, cfg.T, cfg.S, True)
        domains: List[str] = list(getattr(model, "action_mlp", {}).keys())
        shared_used = eng.shared_param_names(named, d)
        dom_used = {dom: eng.domain_param_names(named, d, dom, True) for dom in domains}
        claimed = set(shared_used)
        for v in dom_used.values():
            claimed.update(v)
        leftovers = [k for k in named if k not in claimed]
        order: List[str] = list(shared_used)
        self.shared_size = sum(eng.padded_numel(named[k]) for k in shared_used)
        # every range is laid out [decayed | not decayed] (engine.decay_partition): decayed prefix lengths
        self.shared_decay = eng.decayed_numel(named, shared_used)
        self.dom_range: Dict[str, tuple] = {}
        self.dom_decay: Dict[int, int] = {}  # arena offset of a domain block -> length of its decayed prefix
        off = self.shared_size
        for dom in domains:
            size = sum(eng.padded_numel(named[k]) for k in dom_used[dom])
            self.dom_range[dom] = (off, size)
            self.dom_decay[off] = eng.decayed_numel(named, dom_used[dom])
            off += size
            order += dom_used[dom]
        order += leftovers
        total = sum(eng.padded_numel(named[k]) for k in order)
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.offsets: Dict[str, int] = {}
        off = 0
        with torch.no_grad():
            for k in order:
                p = named[k]
                view = self.flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self.offsets[k] = off
                off += eng.padded_numel(p)
        self.max_dom_size = max([s for _, s in self.dom_range.values()], default=0)


def shared_segment_ranges(names: Sequence[str], sizes: Sequence[int], num_layers: int, segments: int):
    """Ranges of the shared gradient range that become final after each backward segment. The backward runs layers
    L-1 .. 0; segment s covers layers [L - (s+1)*L/segments, L - s*L/segments); the readout and any loss-head parameters
    (computed first) belong to segment 0, the front end (embeddings, positions: computed last) to the last segment.
    names / sizes: the shared parameters in buffer order with their padded lengths. Returns
    (first layer of each segment, [[(offset, length), ...] per segment]) with adjacent ranges merged."""
    segments = max(1, min(segments, num_layers))
    lows = [num_layers - (s + 1) * num_layers // segments for s in range(segments)]
    lows[-1] = 0
    front = ("pos_embed", "token_embed.", "action_mask_tokens", "z_proj", "mask_token", "diffusion_pos_embed")

    def seg_of(name: str) -> int:
        if name.startswith("decoder.layers."):
            i = int(name.split(".")[2])
            for s_, lo in enumerate(lows):
                if i >= lo:
                    return s_
        if name.startswith(front):
            return segments - 1
        return 0

    out = [[] for _ in range(segments)]
    off = 0
    for k, n in zip(names, sizes):
        r = out[seg_of(k)]
        if r and r[-1][0] + r[-1][1] == off:
            r[-1] = (r[-1][0], r[-1][1] + n)
        else:
            r.append((off, n))
        off += n
    return lows, out


def exchange_gradients(grad: torch.Tensor, shared_size: int, max_dom_size: int, dom_range: Dict[str, tuple],
                       rank_domains: Sequence[Optional[str]], gathered: Optional[torch.Tensor], group=None,
                       shared_done: bool = False):
    """Gradient exchange of one step (device- and backend-agnostic: NCCL on GPUs, gloo in the CPU tests).

    grad = [shared | this rank's domain block] (sums, not means). After the call grad[:shared_size] holds the
    sum over ranks, and the returned list [(arena_offset, length, tensor)] holds, for every distinct domain
    trained by some rank this step, the sum of the blocks of the ranks that trained it. Equivalent to the
    reference's dense all-reduce over all parameters (train_multi.py:579,779-781), whose other entries are zero.
    """
    world = dist.get_world_size(group)
    if not shared_done:  # (the overlapped path all-reduced the shared range segment by segment during the backward)
        dist.all_reduce(grad[:shared_size], group=group)
    updates = []
    if max_dom_size:
        send = grad[shared_size:shared_size + max_dom_size]
        dist.all_gather_into_tensor(gathered.view(-1), send, group=group)
        first: Dict[str, int] = {}
        for r, rd in enumerate(rank_domains):
            if rd is None:
                continue
            n_r = dom_range[rd][1]
            if rd in first:
                gathered[first[rd], :n_r] += gathered[r, :n_r]
            else:
                first[rd] = r
        for rd, r in first.items():
            lo, n_r = dom_range[rd]
            updates.append((lo, n_r, gathered[r, :n_r]))
    assert len(rank_domains) == world
    return updates


class TrainStep:
    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.05,
                 max_grad_norm: Optional[float] = 1.0, process_group=None, cuda_graphs: bool = False,
                 mu_transfer: bool = False, overlap_segments: int = 1):
        """The optimizer is the reference's (train_multi.py:899-922): AdamW over two parameter groups — names containing
        "bias" or "layer_norm.weight" get weight_decay 0, everything else `weight_decay`. `mu_transfer=True` selects the
        reference's `mup.MuAdamW`: it divides the learning rate of matrix-like parameters by their width multiplier
        relative to the base shapes, which the reference hard-codes to d_model=256, num_heads=8
        (st_mask_git.py:755-760) — the only width these kernels are built for, so every multiplier is 1 and MuAdamW's
        update is AdamW's; wider models are rejected by engine.check_config before they get here.

        world_size > 1, overlap_segments > 1: the all-reduce of the shared gradient range is issued in that many pieces
        DURING the backward — after the backward of layers 24..31 their gradients go on the wire while layers 23..0 are still
        being computed, and so on (what DDP's buckets do in the reference, train_multi.py:779-781,990) — on NCCL's own stream;
        with CUDA graphs the forward+backward is captured as one graph per segment and the collectives are launched between
        the replays. Parity-tested on two GPUs (tests/ddp_parity_2gpu.py). It is OFF by default because it measured no gain
        on B200: 31.64 vs 31.65 ms/step at N=2 and 32.24 vs 32.12 at N=8 (4 segments vs 1) — the persistent GEMM / attention
        CTAs fill every SM's register file, so an NCCL kernel only gets SMs at kernel boundaries and then delays the compute
        CTAs queued behind it by as much as it hides.

        cuda_graphs=True: forward + loss + backward of each (domain, shape) is captured into a CUDA graph on its
        second use (or by precapture()) and replayed from static input buffers afterwards; the gradient exchange,
        the clip and AdamW stay ordinary launches. The ~1400 launches of a step otherwise cost ~30 ms of host time."""
        self.model = model
        self.engine: Engine = model._engine
        if mu_transfer:
            assert model.config.d_model == 256 and model.config.num_heads == 8, "MuAdamW: only the base width is built"
        self.mu_transfer = mu_transfer
        self.arena = ParamArena(model)
        model._arena = self.arena  # save_pretrained() clones arena-backed parameters (model._save_pretrained)
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_grad_norm
        dev = self.arena.flat.device
        self.m = torch.zeros_like(self.arena.flat)
        self.v = torch.zeros_like(self.arena.flat)
        self.grad = torch.zeros(self.arena.shared_size + self.arena.max_dom_size, device=dev, dtype=torch.float32)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.ones = torch.ones(1, device=dev, dtype=torch.float32)
        self.step_count = 0
        self.dom_steps: Dict[int, int] = {}  # arena offset of a domain block -> optimizer steps it has received
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        self.gathered = (torch.zeros(self.world, self.arena.max_dom_size, device=dev, dtype=torch.float32)
                         if self.world > 1 else None)
        self.overlap_segments = overlap_segments if (self.world > 1 or os.environ.get("HMA_B200_FORCE_SEGMENTS") == "1") else 1
        self._seg = None       # (first layer of each segment, ranges per segment), built on first use
        self._works: list = []  # in-flight all-reduces of this step
        self._p = None
        # nn.Dropout(mlp_drop) (st_transformer.py:24-27): keep masks are keyed by (seed, element); the device-side part of
        # the seed is redrawn every step, so a CUDA-graph replay gets new masks
        self._seed_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self.cuda_graphs = cuda_graphs
        self._graphs: Dict[tuple, dict] = {}
        self._warm = set()
        self._pool = None
        self._capture_stream = None

    def _params(self) -> Dict[str, torch.Tensor]:
        if self._p is None:
            p = {k: b for k, b in self.model.named_buffers()}
            p.update({k: v.data for k, v in self.model.named_parameters()})
            self._p = p
        return self._p

    def _adamw(self, lo: int, n: int, n_decay: int, g: torch.Tensor, scale: float, step: int) -> None:
        a = self.arena.flat
        _lib.call("hma_adamw_step", a.data_ptr() + 4 * lo, g.data_ptr(), self.m.data_ptr() + 4 * lo,
                  self.v.data_ptr() + 4 * lo, n, n_decay, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                  step, scale, self.sumsq.data_ptr() if self.max_norm is not None else None,
                  float(self.max_norm or 0.0), ops._s())

    # ------------------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        """Optimizer state for checkpoint / resume (what `accelerator.save_state` keeps of AdamW, train_multi.py:321):
        both moments over the whole arena, the shared step count and the per-domain step counts (keyed by domain name)."""
        off_to_dom = {lo: dom for dom, (lo, _) in self.arena.dom_range.items()}
        return {"m": self.m.detach().clone(), "v": self.v.detach().clone(), "step_count": self.step_count,
                "dom_steps": {off_to_dom[lo]: n for lo, n in self.dom_steps.items()},
                "param_offsets": dict(self.arena.offsets),
                "hyper": {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd,
                          "max_grad_norm": self.max_norm, "mu_transfer": self.mu_transfer}}

    def load_state_dict(self, state: dict) -> None:
        """Resume from state_dict(): the model's parameters must already hold the checkpointed values (they live in the
        arena; load them with model.load_state_dict / from_pretrained before building the TrainStep, or after — the
        parameters are views, so load_state_dict copies in place)."""
        if state["param_offsets"] != self.arena.offsets:
            raise ValueError("optimizer state was saved for a different parameter layout (domains / config differ)")
        self.m.copy_(state["m"])
        self.v.copy_(state["v"])
        self.step_count = int(state["step_count"])
        self.dom_steps = {self.arena.dom_range[dom][0]: int(n) for dom, n in state["dom_steps"].items()}
        hy = state.get("hyper", {})
        self.lr, self.eps, self.wd = hy.get("lr", self.lr), hy.get("eps", self.eps), hy.get("weight_decay", self.wd)
        self.betas, self.max_norm = tuple(hy.get("betas", self.betas)), hy.get("max_grad_norm", self.max_norm)

    # ------------------------------------------------------------------------------------------
    def _segments(self, p, d):
        if self._seg is None:
            eng = self.engine
            named = {k: v for k, v in self.model.named_parameters()}
            names = eng.shared_param_names(named, d)
            self._seg = shared_segment_ranges(names, [eng.padded_numel(named[k]) for k in names], d.num_layers, self.overlap_segments)
            self.overlap_segments = len(self._seg[0])  # at most one segment per layer
        return self._seg

    def _reduce_segment(self, s: int) -> None:
        """All-reduce the shared gradient ranges that segment `s` of the backward has completed (async: NCCL's stream waits
        for the work enqueued so far on the current stream, the backward goes on)."""
        if self.world == 1:  # HMA_B200_FORCE_SEGMENTS=1: segmented capture without a process group (single-GPU tests)
            return
        for off, n in self._seg[1][s]:
            self._works.append(dist.all_reduce(self.grad[off:off + n], group=self.pg, async_op=True))

    def _fwd_bwd_eager(self, p, ids, labels, action_ids, dom, d, seg_hook=None) -> torch.Tensor:
        eng = self.engine
        B, T, S = d.B, d.T, d.S
        mlp_drop = float(getattr(self.model.config, "mlp_drop", 0.0))
        drop = (mlp_drop, 0x5EED, self._seed_dev) if (mlp_drop > 0.0 and self.model.training) else None
        logits, sv = eng.forward(p, ids, action_ids, dom, d, training=True, drop=drop)
        loss_acc, lse, sums = ops.ce_fwd(logits, labels, ids, B, T, S, d.nv, d.vs, d.mask_id, SMOOTHING)
        dlogits = ops.ce_bwd(logits, labels, ids, B, T, S, d.nv, d.vs, d.mask_id, SMOOTHING, lse, sums, self.ones)
        self.grad.zero_()
        hook = None
        if seg_hook is not None:  # segment s ends after the backward of its first layer
            lows = self._segments(p, d)[0]
            hook = (set(lows[:-1]), lambda i: seg_hook(lows.index(i)))
        eng.backward(p, sv, dlogits, flat=self.grad, layer_hook=hook)
        if seg_hook is not None:
            seg_hook(self.overlap_segments - 1)
        return loss_acc

    def _fwd_bwd(self, p, ids, labels, action_ids, dom, d) -> torch.Tensor:
        """Forward, fused loss and backward into self.grad; returns the device tensor [loss, acc]. With world_size > 1 and
        overlap_segments > 1 the shared-range all-reduces are in flight (self._works) when this returns."""
        overlap = self.overlap_segments > 1
        if overlap:
            self._segments(p, d)
            self._works = []
        if not self.cuda_graphs:
            return self._fwd_bwd_eager(p, ids, labels, action_ids, dom, d, self._reduce_segment if overlap else None)
        key = (dom, d.B, d.T, d.S, None if action_ids is None else tuple(action_ids.shape[1:]))
        rec = self._graphs.get(key)
        if rec is None:
            if key not in self._warm:  # first use: eager (fills the bf16 weight-cast tables, one-time kernel attributes)
                self._warm.add(key)
                return self._fwd_bwd_eager(p, ids, labels, action_ids, dom, d, self._reduce_segment if overlap else None)
            rec = {"ids": ids.clone(), "labels": labels.clone(),
                   "actions": None if action_ids is None else action_ids.to(torch.float32).clone()}
            torch.cuda.synchronize()
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            n0 = ops.LAUNCHES
            # one graph per backward segment (a single graph without overlap): the collectives are launched between replays
            graphs = []
            if self._capture_stream is None:  # ONE capture stream for every graph: the caching allocator reuses a freed
                self._capture_stream = torch.cuda.Stream()  # block only on the stream it was allocated on
            stream = self._capture_stream
            stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(stream):
                cur = torch.cuda.CUDAGraph()
                cur.capture_begin(pool=self._pool)

                def cut(s_):
                    nonlocal cur
                    cur.capture_end()
                    graphs.append(cur)
                    if s_ < self.overlap_segments - 1:
                        cur = torch.cuda.CUDAGraph()
                        cur.capture_begin(pool=self._pool)

                if overlap:
                    rec["out"] = self._fwd_bwd_eager(p, rec["ids"], rec["labels"], rec["actions"], dom, d, cut)
                else:
                    rec["out"] = self._fwd_bwd_eager(p, rec["ids"], rec["labels"], rec["actions"], dom, d)
                    cut(0)
            torch.cuda.current_stream().wait_stream(stream)
            rec["launches"] = ops.LAUNCHES - n0
            ops.LAUNCHES = n0  # counted per replay below
            rec["graphs"] = graphs
            self._graphs[key] = rec
        else:
            rec["ids"].copy_(ids)
            rec["labels"].copy_(labels)
            if action_ids is not None:
                rec["actions"].copy_(action_ids)
        for s_, g in enumerate(rec["graphs"]):
            g.replay()
            if overlap:
                self._reduce_segment(s_)
        ops.LAUNCHES += rec["launches"]
        return rec["out"]

    def precapture(self, input_ids: torch.Tensor, labels: torch.Tensor, action_ids: Optional[torch.Tensor], domain) -> None:
        """Warm and capture the graph of this (domain, shape) without touching the parameters (no optimizer step)."""
        assert self.cuda_graphs
        model, eng, cfg = self.model, self.engine, self.model.config
        B = input_ids.shape[0]
        T = cfg.T
        S = input_ids.numel() // (B * T)
        dom = model._domain0(domain, action_ids)
        d = eng.dims(B, T, S, action_ids is not None)
        ids = input_ids.reshape(B, T, S).contiguous()
        lab = labels.reshape(B, T * S).contiguous()
        for _ in range(2):
            self._fwd_bwd(self._params(), ids, lab, action_ids, dom, d)

    def __call__(self, input_ids: torch.Tensor, labels: torch.Tensor, action_ids: Optional[torch.Tensor], domain,
                 rank_domains: Optional[Sequence[str]] = None) -> torch.Tensor:
        """One optimisation step. Returns a device tensor [loss, acc]. `rank_domains[r]` is the domain rank r
        trains this step (defaults to an all_gather_object when world_size > 1)."""
        model, eng, cfg = self.model, self.engine, self.model.config
        B = input_ids.shape[0]
        T = cfg.T
        S = input_ids.numel() // (B * T)
        ids = input_ids.reshape(B, T, S).contiguous()
        labels = labels.reshape(B, T * S).contiguous()
        dom = model._domain0(domain, action_ids)
        d = eng.dims(B, T, S, action_ids is not None)
        p = self._params()
        self._seed_dev.random_()  # new dropout masks every step (no-op cost when mlp_drop == 0)
        loss_acc = self._fwd_bwd(p, ids, labels, action_ids, dom, d)
        self._apply(dom, rank_domains)
        return loss_acc

    def _apply(self, dom: Optional[str], rank_domains: Optional[Sequence[str]] = None) -> None:
        """Gradient exchange, global-norm clip and AdamW on self.grad = [shared | domain block of `dom`]."""
        eng = self.engine
        self.step_count += 1
        shared = self.arena.shared_size
        dom_lo, dom_n = self.arena.dom_range.get(dom, (0, 0)) if dom is not None else (0, 0)
        g_shared = self.grad[:shared]
        g_dom = self.grad[shared:shared + dom_n]
        scale = 1.0 / self.world
        updates = []  # (arena offset, length, gradient tensor)
        if self.world > 1:
            if rank_domains is None:
                gathered: List[Optional[str]] = [None] * self.world
                dist.all_gather_object(gathered, dom, group=self.pg)
                rank_domains = gathered
            for w in self._works:  # the segment-wise all-reduces of the shared range issued during the backward
                w.wait()
            shared_done = bool(self._works)
            self._works = []
            updates = exchange_gradients(self.grad, shared, self.arena.max_dom_size, self.arena.dom_range, rank_domains,
                                         self.gathered, self.pg, shared_done=shared_done)
        elif dom_n:
            updates.append((dom_lo, dom_n, g_dom))
        if self.max_norm is not None:
            self.sumsq.zero_()
            _lib.call("hma_sumsq", g_shared.data_ptr(), shared, self.sumsq.data_ptr(), ops._s())
            for _, n_r, g in updates:
                _lib.call("hma_sumsq", g.data_ptr(), n_r, self.sumsq.data_ptr(), ops._s())
        self._adamw(0, shared, self.arena.shared_decay, g_shared, scale, self.step_count)
        for lo, n_r, g in updates:
            # torch.optim.AdamW counts steps per parameter and skips parameters without a gradient, so a domain's
            # bias correction follows the number of updates THAT domain has received (train_multi.py:593-598)
            self.dom_steps[lo] = self.dom_steps.get(lo, 0) + 1
            self._adamw(lo, n_r, self.arena.dom_decay[lo], g, scale, self.dom_steps[lo])
        # parameters changed underneath torch's version counters: drop the cached bf16 inference copies
        eng.weights._versions.clear()
        eng._stem_w0.clear()
