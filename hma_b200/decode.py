"""Frame-incremental MaskGIT decode: per-layer temporal K/V cache + CUDA-graph replay of the one-frame pass.

The reference's `maskgit_generate` (st_mask_git.py:337-467) re-runs `compute_logits` on the whole
T-frame window for each of the K MaskGIT steps of each generated frame. Frames before `out_t` cannot
change during those steps and, the temporal attention being causal (st_transformer.py:111) and the
spatial attention per-frame, they influence frame `out_t` only through their temporal keys/values.
A `DecodeSession` therefore
  * runs the context frames through the network ONCE ("prefill"), keeping every layer's temporal K/V;
  * runs only the frame being generated for each MaskGIT step ("step": 1/T of the reference's work);
  * runs a finished frame once more to add its K/V to the cache before the next frame ("commit") — together with the
    first MaskGIT step of that next frame, whose input (a fully masked frame) is known in advance ("commit_step": one
    pass over two frames instead of two passes over one; the logits wait in the session until `step(first=True)`);
  * likewise the prefill can carry the frame to generate as one more, fully masked, frame ("prefill_step").
Same logits as the full-window recompute up to bf16 summation order (tests/test_decode_gpu.py).

A one-frame pass is ~450 launches of a few microseconds each, i.e. CPU-launch-bound, so each
(frame index, mode) pass is captured once into a CUDA graph and replayed; all buffers a graph touches
(token ids, K/V cache, per-frame action conditioning, logits) are owned by the session and static.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops
from .engine import Engine


class DecodeSession:
    def __init__(self, model, B: int, T: int, S: int, dom: Optional[str], d_action: int, device: torch.device,
                 use_graphs: bool = True):
        self.model = model
        self.eng: Engine = model._engine
        self.B, self.T, self.S, self.dom, self.d_action = B, T, S, dom, d_action
        self.device = device
        self.use_graphs = use_graphs
        cfg = model.config
        self.d1 = self.eng.dims(B, 1, S, dom is not None)
        n, L = self.d1.n, cfg.num_layers
        self.kv = torch.zeros(L, T, B * n, 512, device=device, dtype=torch.bfloat16)
        self.ids = torch.zeros(B, 1, S, device=device, dtype=torch.long)
        self.d2 = self.eng.dims(B, 2, S, dom is not None)
        self.ids2 = torch.zeros(B, 2, S, device=device, dtype=torch.long)
        self.mask_id = cfg.image_vocab_size
        self.first_logits: Optional[torch.Tensor] = None  # logits of the first MaskGIT step of frame `first_t`, computed
        self.first_t = -1                                 # by the pass that committed the frame before it
        self.act2 = self.mods2 = None
        self.act_tb = self.hmods_tb = self.mods_tb = None
        if dom is not None:
            self.act_tb = torch.zeros(T * B, 256, device=device, dtype=torch.float32)
            if self.d1.modulate:
                self.hmods_tb = torch.zeros(L, T * B, 256, device=device, dtype=torch.bfloat16)
                self.mods_tb = torch.zeros(L, T * B, 512, device=device, dtype=torch.float32)
                self.mods2 = torch.zeros(L, B * 2, 512, device=device, dtype=torch.float32)
            self.act2 = torch.zeros(B * 2, 256, device=device, dtype=torch.float32)
        self.graphs: Dict[Tuple[int, str], torch.cuda.CUDAGraph] = {}
        self.outputs: Dict[Tuple[int, str], Optional[torch.Tensor]] = {}
        self.warm = set()
        self.pool = None
        self.filled = 0
        self._p: Optional[Dict[str, torch.Tensor]] = None
        self._sig = None
        self._prefill: Dict[tuple, dict] = {}   # (n_ctx, skip_normalization) -> static inputs + captured prefill graph
        self._prefill_warm = set()

    # ------------------------------------------------------------------------------------------
    def signature(self, p: Dict[str, torch.Tensor]):
        """Addresses the captured graphs have baked in (fp32 biases / norms are read in place)."""
        return tuple(p[k].data_ptr() for k in ("pos_embed_TSC", "out_x_proj.bias", "decoder.layers.0.mlp.fc1.bias")
                     if k in p)

    def _begin_eager(self, p, ids: torch.Tensor, actions: Optional[torch.Tensor], n_ctx: int, skip_normalization: bool):
        """`ids` holds the n_ctx context frames plus the (fully masked) frame to generate: one "prefill_step" pass fills the
        cache AND returns that frame's first-step logits (its provisional K/V at cache frame n_ctx are never read: every
        later pass reads [0, n_prev) and the frame's own commit overwrites them)."""
        eng, B, T, S, dom = self.eng, self.B, self.T, self.S, self.dom
        nf = ids.shape[1]
        dp = eng.dims(B, nf, S, dom is not None)
        a_ctx = actions[:, :nf].contiguous() if actions is not None else None
        logits, _ = eng.forward(p, ids, a_ctx, dom, dp, False, skip_normalization, t0=0, kv=self.kv,
                                mode="prefill_step" if nf == n_ctx + 1 else "prefill")
        if dom is not None:
            a_tb = actions.transpose(0, 1).reshape(T * B, -1).to(torch.float32).contiguous()  # (t, b) row order
            act, c_bf = eng.action_stem(p, a_tb, dom, skip_normalization)
            self.act_tb.copy_(act)
            if self.d1.modulate:
                eng.modulation_all_layers(p, c_bf, dom, self.d1.num_layers, False, hmods=self.hmods_tb, mods=self.mods_tb)
        return logits

    def begin(self, p: Dict[str, torch.Tensor], prompt_THW: torch.Tensor, n_ctx: int, actions: Optional[torch.Tensor],
              skip_normalization: bool, first_step: bool = True) -> None:
        """Prefill: context frames [0, n_ctx) -> K/V cache; action conditioning of every frame of the window. With CUDA
        graphs the whole prefill is replayed from static inputs from its third use on (the interactive loop of
        sim/simulator.py:233-372 re-prompts a sliding window every step: at B=1 it is purely launch-bound).
        `first_step`: the pass also carries frame n_ctx as a fully masked frame and keeps its logits for
        `step(..., first=True)` — the first MaskGIT step of the frame about to be generated costs no pass of its own."""
        eng, B, T, S, dom = self.eng, self.B, self.T, self.S, self.dom
        sig = self.signature(p)
        if sig != self._sig:  # parameters were re-allocated: captured graphs point at dead memory
            self.graphs.clear()
            self.outputs.clear()
            self._prefill.clear()
            self._sig = sig
        self._p = p
        first_step = bool(first_step and n_ctx < T)
        ids = prompt_THW[:, :n_ctx].reshape(B, n_ctx, S)
        if first_step:
            ids = torch.cat([ids, torch.full((B, 1, S), self.mask_id, dtype=ids.dtype, device=ids.device)], dim=1)
        nf = ids.shape[1]
        key = (n_ctx, bool(skip_normalization), first_step)
        if not self.use_graphs or key not in self._prefill_warm:
            self._prefill_warm.add(key)
            logits = self._begin_eager(p, ids.contiguous(), actions, n_ctx, skip_normalization)
        else:
            eng.prepare_weights(p, eng.dims(B, nf, S, dom is not None), dom, False)  # refresh bf16 copies outside the graph
            rec = self._prefill.get(key)
            if rec is None:
                rec = {"ids": ids.contiguous().clone(), "actions": None if actions is None else actions.to(torch.float32).clone()}
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                if self.pool is None:
                    self.pool = torch.cuda.graph_pool_handle()
                with torch.cuda.graph(g, pool=self.pool):
                    rec["logits"] = self._begin_eager(p, rec["ids"], rec["actions"], n_ctx, skip_normalization)
                rec["graph"] = g
                self._prefill[key] = rec
            else:
                rec["ids"].copy_(ids)
                if actions is not None:
                    rec["actions"].copy_(actions)
            rec["graph"].replay()
            logits = rec["logits"]
        self.filled = n_ctx
        self.first_logits, self.first_t = (logits, n_ctx) if first_step else (None, -1)

    def _cond2(self, t: int):
        """Action conditioning of frames t, t + 1 in (b, t) row order (the tables are (t, b)); static buffers."""
        if self.dom is None:
            return None
        B = self.B
        self.act2.view(B, 2, 256).copy_(self.act_tb[t * B:(t + 2) * B].view(2, B, 256).transpose(0, 1))
        if self.mods2 is not None:
            L = self.mods2.shape[0]
            self.mods2.view(L, B, 2, 512).copy_(self.mods_tb[:, t * B:(t + 2) * B].view(L, 2, B, 512).transpose(1, 2))
        return (self.act2, self.mods2)

    def _cond(self, t: int):
        if self.dom is None:
            return None
        B = self.B
        act = self.act_tb[t * B:(t + 1) * B]
        mods = self.mods_tb[:, t * B:(t + 1) * B] if self.mods_tb is not None else None
        return (act, mods)

    def _run(self, t: int, mode: str):
        if mode == "commit_step":
            logits, _ = self.eng.forward(self._p, self.ids2, None, self.dom, self.d2, False, t0=t, kv=self.kv, mode=mode,
                                         frame_cond=self._cond2(t))
            return logits
        logits, _ = self.eng.forward(self._p, self.ids, None, self.dom, self.d1, False, t0=t, kv=self.kv, mode=mode,
                                     frame_cond=self._cond(t))
        return logits

    def _pass(self, frame_ids: torch.Tensor, t: int, mode: str) -> Optional[torch.Tensor]:
        assert t == self.filled, f"decode session holds {self.filled} frames of context, asked for frame {t}"
        if mode == "commit_step":
            self.ids2[:, 0].copy_(frame_ids.reshape(self.B, self.S))
            self.ids2[:, 1].fill_(self.mask_id)
        else:
            self.ids.copy_(frame_ids.reshape(self.B, 1, self.S))
        key = (t, mode)
        if not self.use_graphs:
            return self._run(t, mode)
        if mode not in self.warm:  # first use of a mode: run eagerly (lazy one-time kernel attributes, caches)
            self.warm.add(mode)
            return self._run(t, mode)
        g = self.graphs.get(key)
        if g is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            if self.pool is None:
                self.pool = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(g, pool=self.pool):
                out = self._run(t, mode)
            self.graphs[key] = g
            self.outputs[key] = out  # keeps the graph's output block alive (other graphs share the pool)
        g.replay()
        return self.outputs[key]

    def step(self, frame_ids: torch.Tensor, t: int, first: bool = False) -> torch.Tensor:
        """Logits fp32 [B*S, nv*vs] of window frame t given the cached context (valid until the next pass). `first`: this
        is the first MaskGIT step of the frame, i.e. `frame_ids` is fully masked — then the logits the previous frame's
        commit pass already produced are returned without another pass."""
        if first and self.first_t == t and self.first_logits is not None:
            self.first_t = -1
            return self.first_logits
        self.first_t = -1
        return self._pass(frame_ids, t, "step")

    def commit(self, frame_ids: torch.Tensor, t: int, prefetch_next: bool = False) -> None:
        """Add finished frame t to the context. `prefetch_next`: the same pass also runs frame t + 1 as a fully masked frame
        and keeps its logits for `step(..., first=True)`."""
        if prefetch_next and t + 1 < self.T:
            self.first_logits = self._pass(frame_ids, t, "commit_step")
            self.first_t = t + 1
        else:
            self._pass(frame_ids, t, "commit")
            self.first_t = -1
        self.filled = t + 1
