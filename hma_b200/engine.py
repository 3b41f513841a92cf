"""Host-side schedule of the ST-MaskGIT hot path: which CUDA stage runs on which buffer, in what
order, forward and backward. Mirrors, stage by stage, the reference call stack
(SURVEY.md §3.2): STMaskGIT.compute_logits (st_mask_git.py:632-686) -> STTransformerDecoder /
STBlock.forward (st_transformer.py:79-114,172-177) -> compute_video_loss_and_acc (:603-630), and
the autograd transpose of all of it.

Memory layout: the residual stream is fp32 [B*T*n, 256] in (b, t, s) token order for the whole
network (n = S video tokens + A action tokens per frame); every GEMM operand is a bf16 matrix
over the same token order, so neither the spatial nor the temporal stage ever transposes
(the temporal kernel strides over frames in place). Weights are fp32 master parameters
(reference state_dict layout); bf16 copies (and their transposes, for input gradients) are made
per step.

All math runs in libhma_b200.so; torch only allocates buffers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import os

import torch

from . import ops
from .ops import EPI_BF16, EPI_DGELU, EPI_DSILU, EPI_GELU, EPI_RESID, EPI_SILU

Tensor = torch.Tensor
C = 256
SMOOTHING = 0.01  # st_mask_git.py:620


@dataclass
class Dims:
    B: int
    T: int
    S: int
    A: int  # action tokens per frame actually concatenated (0 if none)
    heads: int
    nv: int
    vs: int
    mask_id: int
    scale: float
    num_layers: int
    modulate: bool
    readout_alpha: float
    additive: bool = False  # action_network containing "mlp": the action embedding is added to every token of its frame

    @property
    def n(self) -> int:
        return self.S + self.A

    @property
    def M(self) -> int:
        return self.B * self.T

    @property
    def N(self) -> int:
        return self.B * self.T * self.n


def check_config(cfg) -> None:
    """Reject, loudly, what the CUDA path does not implement (no silent fallback)."""
    if cfg.d_model != 256 or cfg.num_heads != 8:
        raise NotImplementedError(f"hma_b200 kernels are built for d_model=256, num_heads=8 (got {cfg.d_model}, {cfg.num_heads})")
    if int(cfg.d_model * cfg.mlp_ratio) != 1024:
        raise NotImplementedError("mlp_ratio must be 4.0")
    if cfg.jointly_predict_actions or not cfg.jointly_predict_states:
        raise NotImplementedError("jointly_predict_actions / jointly_predict_states=False are not implemented")
    net = cfg.action_network
    if "cross_attention" in net and "mlp" not in net:
        raise NotImplementedError(f"action_network={net!r} is not implemented (supported: mlp, concat, modulate and their "
                                  "combinations; st_transformer.py:93-104 checks 'mlp' first, then 'cross_attention', then 'modulate')")
    if cfg.num_factored_vocabs not in (1, 2) or cfg.factored_vocab_size % 256 != 0 or \
            cfg.num_factored_vocabs * cfg.factored_vocab_size > 1024:
        raise NotImplementedError("unsupported factorised vocabulary")


class Weights:
    """bf16 operand copies of the fp32 master parameters (plain for forward, transposed for dgrad),
    refreshed by ONE batched cast/transpose launch per step."""

    def __init__(self):
        self.plain: Dict[str, Tensor] = {}
        self.trans: Dict[str, Tensor] = {}
        self._versions: Dict[str, int] = {}
        self._tables: Dict[tuple, tuple] = {}

    def prepare(self, params: Dict[str, Tensor], names: List[str], need_t: bool, force: bool) -> None:
        todo = []
        for name in names:
            w = params[name]
            have = name in self.plain and (not need_t or name in self.trans)
            if have and not force and self._versions.get(name) == w._version:
                continue
            todo.append(name)
        if not todo:
            return
        key = (tuple(todo), need_t)
        srcs = tuple(params[n].data_ptr() for n in todo)
        cached = self._tables.get(key)
        if cached is None or cached[0] != srcs:
            rows = []
            for n in todo:
                w = params[n]
                assert w.dim() == 2 and w.is_contiguous() and w.dtype == torch.float32
                R, Cc = w.shape
                if n not in self.plain:
                    self.plain[n] = torch.empty(R, Cc, device=w.device, dtype=torch.bfloat16)
                if need_t and n not in self.trans:
                    self.trans[n] = torch.empty(Cc, R, device=w.device, dtype=torch.bfloat16)
                rows.append([w.data_ptr(), self.plain[n].data_ptr(), self.trans[n].data_ptr() if need_t else 0, R, Cc])
            dev = params[todo[0]].device
            table = torch.tensor(rows, dtype=torch.int64).to(dev)
            cached = (srcs, table, max(r[3] for r in rows), max(r[4] for r in rows))
            self._tables[key] = cached
        ops.cast_transpose_batched(cached[1], len(todo), cached[2], cached[3])
        for n in todo:
            self._versions[n] = params[n]._version


def layer_matrix_names(i: int, dom: Optional[str], modulate: bool) -> List[str]:
    p = f"decoder.layers.{i}."
    names = [p + "spatial_attn.qkv.weight", p + "spatial_attn.proj.weight", p + "temporal_attn.qkv.weight",
             p + "temporal_attn.proj.weight", p + "mlp.fc1.weight", p + "mlp.fc2.weight"]
    if modulate:
        q = p + f"action_projectors.{dom}."
        names += [q + "linear_out.weight", q + "adaLN_modulation.0.weight", q + "adaLN_modulation.2.weight"]
    return names


NO_DECAY = ("bias", "layer_norm.weight")  # train_multi.py:906: substrings of parameter names that get weight_decay 0


def is_no_decay(name: str) -> bool:
    """The reference's optimizer grouping rule (train_multi.py:906-917), verbatim: a parameter whose NAME contains "bias"
    or "layer_norm.weight" is not decayed. (No HMA module is called `layer_norm`, so norm1/norm2/LayerNorm gains ARE
    decayed by the reference; kept.)"""
    return any(nd in name for nd in NO_DECAY)


def decay_partition(names: List[str]) -> List[str]:
    """Stable partition [decayed | not decayed]: inside every flat range (parameter arena, moments, gradient buffer) the
    decayed tensors come first, so the optimizer needs one boundary per range instead of a per-tensor table."""
    return [k for k in names if not is_no_decay(k)] + [k for k in names if is_no_decay(k)]


def stem_pad(da: int) -> int:
    return (da + 127) // 128 * 128


class Engine:
    """One engine per model; owns the bf16 weight cache."""

    def __init__(self, cfg):
        self.check(cfg)
        self.cfg = cfg
        self.weights = Weights()
        self._stem_w0: Dict[str, tuple] = {}
        self._side: Dict[int, torch.cuda.Stream] = {}
        # HMA_B200_FUSE_LN=1: LayerNorms emitted by the preceding residual GEMM's epilogue (ops.gemm_nt_ln) instead of separate
        # row-wise kernels. Off by default: measured on config 2 the fused step is 30.37 ms against 30.04 ms — inside a step
        # the row-wise kernel finds the stream in L2, while the epilogue's extra work sits on the GEMM's critical path
        # (DESIGN.md 3.9).
        self.fuse_ln = os.environ.get("HMA_B200_FUSE_LN", "0") == "1"

    @staticmethod
    def check(cfg) -> None:
        check_config(cfg)

    def _side_stream(self, dev) -> "torch.cuda.Stream":
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if idx not in self._side:
            self._side[idx] = torch.cuda.Stream(device=idx)
        return self._side[idx]

    # ------------------------------------------------------------------------------------------
    def dims(self, B: int, T: int, S: int, with_actions: bool) -> Dims:
        cfg = self.cfg
        net = cfg.action_network
        A = cfg.action_token_size if (with_actions and "concat" in net) else 0
        hd = cfg.d_model // cfg.num_heads
        scale = 8.0 / hd if cfg.use_mup else hd ** -0.5  # attention.py:27
        return Dims(B=B, T=T, S=S, A=A, heads=cfg.num_heads, nv=cfg.num_factored_vocabs, vs=cfg.factored_vocab_size,
                    mask_id=cfg.image_vocab_size, scale=scale, num_layers=cfg.num_layers,
                    modulate=with_actions and "modulate" in net and "mlp" not in net and "cross_attention" not in net,
                    readout_alpha=(256.0 / cfg.d_model) if cfg.use_mup else 1.0,  # st_mask_git.py:755-760,788-789
                    additive=with_actions and "mlp" in net)

    def _prepare(self, p: Dict[str, Tensor], d: Dims, dom: Optional[str], training: bool) -> None:
        names = ["out_x_proj.weight"]
        for i in range(d.num_layers):
            names += layer_matrix_names(i, dom, d.modulate)
        if dom is not None:
            names.append(f"action_mlp.{dom}.model.3.weight")
        self.weights.prepare(p, names, need_t=training, force=training)
        if dom is not None:
            w0 = p[f"action_mlp.{dom}.model.0.weight"]
            key = f"action_mlp.{dom}.model.0.weight"
            cached = self._stem_w0.get(key)
            if training or cached is None or cached[0] != w0._version:
                self._stem_w0[key] = (w0._version, ops.action_prep(w0.detach().contiguous(), stem_pad(w0.shape[1])))

    # ------------------------------------------------------------------------------------------
    def action_stem(self, p: Dict[str, Tensor], a2d: Tensor, dom: str, skip_normalization: bool, sv: Optional[dict] = None):
        """ActionStat + BasicMLP (st_mask_git.py:645-649) on rows [rows, d_a] -> (act fp32 [rows,256], bf16 copy)."""
        Wp = self.weights.plain
        da = a2d.shape[1]
        mean = std = None
        if not skip_normalization:
            mean, std = p[f"action_preprocessor.{dom}.mean"], p[f"action_preprocessor.{dom}.std"]
        q = f"action_mlp.{dom}.model."
        a_prep = ops.action_prep(a2d, stem_pad(da), mean, std)
        h1 = ops.gemm_nt(a_prep, self._stem_w0[q + "0.weight"][1], EPI_RESID, bias=p[q + "0.bias"])
        h1n, st_stem = ops.ln_relu_fwd(h1, p[q + "1.weight"], p[q + "1.bias"])
        act = ops.gemm_nt(h1n, Wp[q + "3.weight"], EPI_RESID, bias=p[q + "3.bias"])
        c_bf = ops.ln_fwd(act, 0)
        if sv is not None:
            sv.update(a_prep=a_prep, h1=h1, h1n=h1n, st_stem=st_stem, c_bf=c_bf, da=da)
        return act, c_bf

    def modulation_all_layers(self, p: Dict[str, Tensor], c_bf: Tensor, dom: str, num_layers: int, want_z: bool,
                              hmods: Optional[Tensor] = None, mods: Optional[Tensor] = None):
        """adaLN_modulation (Linear -> SiLU -> Linear, st_mask_git.py:61-63,70) of every layer for rows c_bf.
        Runs on the caller's current stream. Returns (hmods bf16 [L,rows,256], zmods or None, mods fp32 [L,rows,512])."""
        Wp = self.weights.plain
        rows, dev = c_bf.shape[0], c_bf.device
        if hmods is None:
            hmods = torch.empty(num_layers, rows, C, device=dev, dtype=torch.bfloat16)
        zmods = torch.empty(num_layers, rows, C, device=dev, dtype=torch.bfloat16) if want_z else None
        if mods is None:
            mods = torch.empty(num_layers, rows, 2 * C, device=dev, dtype=torch.float32)
        events = []
        for i in range(num_layers):
            ap = f"decoder.layers.{i}.action_projectors.{dom}."
            ops.gemm_nt(c_bf, Wp[ap + "adaLN_modulation.0.weight"], EPI_SILU, bias=p[ap + "adaLN_modulation.0.bias"],
                        out=hmods[i], out2=zmods[i] if want_z else None)
            ops.gemm_nt(hmods[i], Wp[ap + "adaLN_modulation.2.weight"], EPI_RESID,
                        bias=p[ap + "adaLN_modulation.2.bias"], out=mods[i])
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            events.append(ev)
        return hmods, zmods, mods, events

    def prepare_weights(self, p: Dict[str, Tensor], d: Dims, dom: Optional[str], training: bool) -> None:
        self._prepare(p, d, dom, training)

    def forward(self, p: Dict[str, Tensor], ids: Tensor, actions: Optional[Tensor], dom: Optional[str], d: Dims,
                training: bool, skip_normalization: bool = False, *, t0: int = 0, kv=None, mode: str = "full",
                frame_cond=None, front=None, head: bool = True, drop=None):
        """ids: i64 [B, T, S] (contiguous). Returns (logits fp32 [B*T*S, nv*vs], saved or None).

        Frame-incremental decode (inference only; `kv` = per-layer K/V cache tensors [Tmax, B*n, 512]):
          mode "prefill": the T frames given are frames [t0, t0+T) of the window; their temporal K/V are
                          appended to the cache; no head (returns None logits).
          mode "step"   : T == 1, the frame is window frame t0; temporal attention reads cache frames [0, t0).
          mode "commit" : as "step", but appends this frame's K/V to the cache and skips everything after
                          the last layer's temporal K/V (no logits).
          mode "commit_step": T == 2 = a finished frame t0 and the (still fully masked) frame t0 + 1 in ONE pass: frame t0's
                          K/V join the cache layer by layer and frame t0 + 1 attends to them, so the pass both commits t0 and
                          returns the logits of t0 + 1's first MaskGIT step — two passes of B*n rows become one of 2*B*n.
          mode "prefill_step": as "prefill", but the last of the T frames is the (fully masked) frame to generate: no early
                          exit, and the logits of that frame are returned (the reference's first full-window step).
        `frame_cond` (step/commit/commit_step): precomputed (act fp32 [B*T,256] or None, mods fp32 [L,B*T,512] or None) of the
        frame(s) in (b, t) row order, replacing the action stem + adaLN chain. `actions`, if given, must hold exactly the T
        frames computed. With commit_step / prefill_step the logits cover the LAST frame only: fp32 [B*S, nv*vs].

        Other front ends / heads over the same trunk (STMAR, mar.py): `front(act)` returns the fp32 residual stream
        [B*T*n, 256] in place of the token embedding (`ids` is then ignored); `head=False` returns the trunk output
        instead of logits. `drop = (p, seed, seed_dev)`: nn.Dropout(p) after the GELU and after fc2
        (st_transformer.py:24-27; training only); seed_dev is an optional device u64 mixed into the seed at run time."""
        assert mode in ("full", "prefill", "step", "commit", "commit_step", "prefill_step")
        assert mode == "full" or (not training and kv is not None)
        W = self.weights
        has_act = actions is not None or frame_cond is not None
        if frame_cond is None:
            self._prepare(p, d, dom if actions is not None else None, training)
        Wp, sv = W.plain, {}
        B, T, S, n, M, N = d.B, d.T, d.S, d.n, d.M, d.N
        act = c_bf = None
        if frame_cond is not None:
            assert (mode in ("step", "commit") and T == 1) or (mode == "commit_step" and T == 2)
            act = frame_cond[0]
        elif actions is not None:
            a2d = actions.reshape(M, -1).to(torch.float32).contiguous()
            act, c_bf = self.action_stem(p, a2d, dom, skip_normalization, sv if training else None)
        pos = p["pos_embed_TSC"]
        pos_n = pos.shape[2]
        if t0:
            pos = pos[:, t0:]
        if front is not None:
            x = front(act if d.A else None)
        else:
            E1 = p.get("token_embed.factored_embeds.1.weight") if d.nv == 2 else None
            x = ops.embed_fwd(ids, p["token_embed.factored_embeds.0.weight"], E1, p["token_embed.mask_token_embed"],
                              act if d.A else None, pos, pos_n, B, T, S, d.A, d.vs, d.mask_id)
        drop_p = drop[0] if (drop is not None and training) else 0.0
        # adaLN_modulation (Linear -> SiLU -> Linear on the [B*T, 256] action embedding) depends only on the
        # stem output: all layers' shift/scale are produced up-front on a side stream, where these
        # one-tile GEMMs fill the tails of the main stream's kernels instead of serialising with them.
        hmods = zmods = mods = None
        mod_events = []
        if d.modulate and frame_cond is not None:
            mods = frame_cond[1]
        elif d.modulate:
            main = torch.cuda.current_stream()
            side = self._side_stream(x.device)
            fork = torch.cuda.Event()
            fork.record(main)
            with torch.cuda.stream(side):
                side.wait_event(fork)
                hmods, zmods, mods, mod_events = self.modulation_all_layers(p, c_bf, dom, d.num_layers, training)
        layers = []
        qk = bool(self.cfg.qk_norm)  # norm1 / norm2 are Identity and q, k get a shared per-head LayerNorm (attention.py:32-35)
        # Every LayerNorm of the trunk follows a residual Linear with 256 output columns: that GEMM's epilogue emits the
        # normalised bf16 operand of the next stage itself (ops.gemm_nt_ln), so the stream is not re-read by a row-wise pass.
        fuse_ln = self.fuse_ln and not qk
        a1_next = None  # (LN1 output, stats) of the layer about to run, when the previous layer's fc2 produced them
        for i in range(d.num_layers):
            lp = f"decoder.layers.{i}."
            L = {}
            # ---- spatial attention, pre-norm (st_transformer.py:85-86)
            if qk:
                a1, st1 = ops.ln_fwd(x, 0), None
            elif a1_next is not None:
                (a1, st1), a1_next = a1_next, None
            else:
                a1, st1 = ops.ln_fwd(x, 1, gamma=p[lp + "norm1.weight"], beta=p[lp + "norm1.bias"], eps=1e-5, want_stats=True)
            qkv_s = ops.gemm_nt(a1, Wp[lp + "spatial_attn.qkv.weight"], EPI_BF16, bias=p.get(lp + "spatial_attn.qkv.bias"))
            qkv_s_raw = qkv_t_raw = None
            if qk:
                qkv_s_raw = qkv_s
                qkv_s = ops.qk_norm_fwd(qkv_s_raw, p[lp + "spatial_attn.norm.weight"], p[lp + "spatial_attn.norm.bias"])
            att_s, lse = ops.attn_spatial_fwd(qkv_s, M, n, d.heads, d.scale, want_lse=training)
            plain = not d.modulate and not d.additive  # x2 is x1: the temporal stage's bf16 operand comes out of this GEMM
            at = torch.empty(N, C, device=x.device, dtype=torch.bfloat16)
            if d.modulate and mod_events:
                torch.cuda.current_stream().wait_event(mod_events[i])
            if d.modulate and fuse_ln:  # the projection's epilogue also emits ModulateLayer's (1 + scale) * LN(x1) + shift
                x1, am, stm = ops.gemm_nt_ln(att_s, Wp[lp + "spatial_attn.proj.weight"], resid=x, out=None if training else x,
                                             bias=p.get(lp + "spatial_attn.proj.bias"), ln_mode=2, mod=mods[i], rows_per_group=n,
                                             eps=1e-6, want_stats=True)
            else:
                am = None
                x1 = ops.gemm_nt(att_s, Wp[lp + "spatial_attn.proj.weight"], EPI_RESID, bias=p.get(lp + "spatial_attn.proj.bias"),
                                 resid=x, out=None if training else x, out2=at if plain else None)
            # ---- per-layer action conditioning (st_transformer.py:102-104; st_mask_git.py:66-76)
            if d.modulate:
                ap = lp + f"action_projectors.{dom}."
                mod = mods[i]
                hmod = hmods[i] if training else None
                zmod = zmods[i] if training else None
                if am is None:
                    am, stm = ops.ln_fwd(x1, 2, mod=mod, rows_per_group=n, eps=1e-6, want_stats=True)
                x2 = ops.gemm_nt(am, Wp[ap + "linear_out.weight"], EPI_RESID, bias=p[ap + "linear_out.bias"], resid=x1,
                                 out=None if training else x1, out2=at)
                if training:
                    L.update(zmod=zmod, hmod=hmod, mod=mod, am=am, stm=stm)
            elif d.additive:  # st_transformer.py:93-97: x += action embedding of the frame, broadcast over its tokens
                x2 = ops.group_add(x1, act, n, out=None if training else x1)
            else:
                x2 = x1
            # ---- causal temporal attention, no pre-norm (st_transformer.py:111): `at` = bf16(x2), written by the GEMM
            # that produced x2 (or cast here after the additive conditioning)
            if d.additive and not d.modulate:
                ops.ln_fwd(x2, 0, out=at)
            qkv_t = ops.gemm_nt(at, Wp[lp + "temporal_attn.qkv.weight"], EPI_BF16, bias=p.get(lp + "temporal_attn.qkv.bias"))
            if qk:
                qkv_t_raw = qkv_t
                qkv_t = ops.qk_norm_fwd(qkv_t_raw, p[lp + "temporal_attn.norm.weight"], p[lp + "temporal_attn.norm.bias"])
            if mode in ("step", "commit"):
                lse_t = None
                att_t = ops.attn_temporal_cached(qkv_t, kv[i], t0, d.heads, d.scale)
            elif mode == "commit_step":
                # both frames' K/V go to the cache first (frame t0 + 1's are provisional: its own commit overwrites them),
                # then frame f of the pass attends to cache frames [0, t0 + f) and to itself
                lse_t = None
                ops.kv_cache_append(qkv_t, B, T, n, kv[i], t0)
                att_t = ops.attn_temporal_cached(qkv_t, kv[i], t0, d.heads, d.scale, frames=T, n=n)
            else:
                att_t, lse_t = ops.attn_temporal_fwd(qkv_t, B, T, n, d.heads, d.scale, want_lse=False)  # bwd recomputes the rows
            if mode in ("prefill", "commit", "prefill_step"):
                ops.kv_cache_append(qkv_t, B, T, n, kv[i], t0)
                if i == d.num_layers - 1 and mode != "prefill_step":
                    return None, None  # nothing downstream of the last layer's K/V is needed
            # ---- MLP, pre-norm, erf-GELU (st_transformer.py:24-27,112); norm2 comes out of the temporal projection
            if fuse_ln:
                x3, a2, st2 = ops.gemm_nt_ln(att_t, Wp[lp + "temporal_attn.proj.weight"], resid=x2, out=None if training else x2,
                                             bias=p.get(lp + "temporal_attn.proj.bias"), ln_mode=1, gamma=p[lp + "norm2.weight"],
                                             beta=p[lp + "norm2.bias"], eps=1e-5, want_stats=True)
            else:
                x3 = ops.gemm_nt(att_t, Wp[lp + "temporal_attn.proj.weight"], EPI_RESID, bias=p.get(lp + "temporal_attn.proj.bias"),
                                 resid=x2, out=None if training else x2)
                if qk:
                    a2, st2 = ops.ln_fwd(x3, 0), None
                else:
                    a2, st2 = ops.ln_fwd(x3, 1, gamma=p[lp + "norm2.weight"], beta=p[lp + "norm2.bias"], eps=1e-5, want_stats=True)
            z = torch.empty(N, 1024, device=x.device, dtype=torch.bfloat16) if training else None
            h = ops.gemm_nt(a2, Wp[lp + "mlp.fc1.weight"], EPI_GELU, bias=p.get(lp + "mlp.fc1.bias"), out2=z)
            if drop_p > 0.0:
                ops.dropout_bf16_(h, drop_p, drop[1] + 2 * i, drop[2])
                x4 = ops.dropout_add_f32(ops.gemm_nt(h, Wp[lp + "mlp.fc2.weight"], EPI_RESID, bias=p.get(lp + "mlp.fc2.bias")),
                                         x3, drop_p, drop[1] + 2 * i + 1, seed_dev=drop[2])
            elif fuse_ln and i + 1 < d.num_layers:  # fc2's epilogue emits the next layer's norm1
                nl = f"decoder.layers.{i + 1}."
                x4, a_n, st_n = ops.gemm_nt_ln(h, Wp[lp + "mlp.fc2.weight"], resid=x3, out=None if training else x3,
                                               bias=p.get(lp + "mlp.fc2.bias"), ln_mode=1, gamma=p[nl + "norm1.weight"],
                                               beta=p[nl + "norm1.bias"], eps=1e-5, want_stats=True)
                a1_next = (a_n, st_n)
            else:
                x4 = ops.gemm_nt(h, Wp[lp + "mlp.fc2.weight"], EPI_RESID, bias=p.get(lp + "mlp.fc2.bias"), resid=x3,
                                 out=None if training else x3)
            if training:
                L.update(x0=x, a1=a1, st1=st1, qkv_s=qkv_s, att_s=att_s, lse=lse, x1=x1, at=at, qkv_t=qkv_t, att_t=att_t,
                         x3=x3, a2=a2, st2=st2, z=z, h=h, lse_t=lse_t, qkv_s_raw=qkv_s_raw, qkv_t_raw=qkv_t_raw)
                layers.append(L)
            x = x4
        # ---- head on the video tokens only (st_mask_git.py:681-683)
        if mode in ("commit_step", "prefill_step"):  # the last frame of every sample only: row (b, T-1, s) of (b, t, s)
            ah = ops.ln_fwd(x[(T - 1) * n:], 0, rows=B * S, src_group=T * n, dst_group=S)
        else:
            ah = ops.ln_fwd(x, 0, rows=M * S, src_group=n, dst_group=S) if d.A else ops.ln_fwd(x, 0)
        logits = ops.gemm_nt(ah, Wp["out_x_proj.weight"], EPI_RESID, bias=p["out_x_proj.bias"], alpha=d.readout_alpha)
        if training:
            sv.update(layers=layers, ah=ah, ids=ids, dom=dom, dims=d, pos_n=pos_n, has_actions=actions is not None,
                      drop=(drop_p, drop[1], drop[2]) if drop_p > 0.0 else None)
        return logits, (sv if training else None)

    # ------------------------------------------------------------------------------------------
    def shared_param_names(self, p: Dict[str, Tensor], d: Dims) -> List[str]:
        """Parameters every batch touches (embeddings, trunk, head), in gradient-buffer order."""
        names = self.front_param_names(p, d)
        for i in range(d.num_layers):
            lp = f"decoder.layers.{i}."
            for k in ("norm1.weight", "norm1.bias", "spatial_attn.qkv.weight", "spatial_attn.qkv.bias",
                      "spatial_attn.norm.weight", "spatial_attn.norm.bias",
                      "spatial_attn.proj.weight", "spatial_attn.proj.bias", "temporal_attn.qkv.weight",
                      "temporal_attn.qkv.bias", "temporal_attn.norm.weight", "temporal_attn.norm.bias",
                      "temporal_attn.proj.weight", "temporal_attn.proj.bias", "norm2.weight",
                      "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"):
                if lp + k in p:
                    names.append(lp + k)
        names += ["out_x_proj.weight", "out_x_proj.bias"]
        return decay_partition(names + self.head_param_names(p, d))

    def front_param_names(self, p: Dict[str, Tensor], d: Dims) -> List[str]:
        names = ["pos_embed_TSC", "token_embed.mask_token_embed", "token_embed.factored_embeds.0.weight"]
        if d.nv == 2:
            names.append("token_embed.factored_embeds.1.weight")
        return names

    def head_param_names(self, p: Dict[str, Tensor], d: Dims) -> List[str]:
        return []

    def domain_param_names(self, p: Dict[str, Tensor], d: Dims, dom: Optional[str], has_actions: bool) -> List[str]:
        """Parameters only batches of domain `dom` touch (action stem + per-layer ModulateLayers)."""
        names: List[str] = []
        if not has_actions or dom is None or not (d.A or d.modulate or d.additive):
            return names
        q = f"action_mlp.{dom}.model."
        names += [q + k for k in ("0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias")]
        if d.modulate:
            for i in range(d.num_layers):
                ap = f"decoder.layers.{i}.action_projectors.{dom}."
                names += [ap + "linear_out.weight", ap + "linear_out.bias", ap + "adaLN_modulation.0.weight",
                          ap + "adaLN_modulation.0.bias", ap + "adaLN_modulation.2.weight", ap + "adaLN_modulation.2.bias"]
        return decay_partition(names)

    def active_param_names(self, p: Dict[str, Tensor], d: Dims, dom: Optional[str], has_actions: bool) -> List[str]:
        return self.shared_param_names(p, d) + self.domain_param_names(p, d, dom, has_actions)

    @staticmethod
    def padded_numel(t: Tensor) -> int:
        return (t.numel() + 3) // 4 * 4  # keep every tensor 16-byte aligned inside flat buffers

    def decayed_numel(self, p: Dict[str, Tensor], names: List[str]) -> int:
        """Padded length of the decayed prefix of a range laid out by `names` (see decay_partition)."""
        return sum(self.padded_numel(p[k]) for k in names if not is_no_decay(k))

    def alloc_grads(self, p: Dict[str, Tensor], d: Dims, dom: Optional[str], has_actions: bool, dev,
                    flat: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """fp32 gradients for every active parameter, as views of one flat buffer laid out [shared | domain]
        (`flat`, if given, must be zeroed and large enough)."""
        names = self.active_param_names(p, d, dom, has_actions)
        sizes = [p[k].numel() for k in names]
        padded = [self.padded_numel(p[k]) for k in names]
        if flat is None:
            flat = torch.zeros(sum(padded), device=dev, dtype=torch.float32)
        assert flat.numel() >= sum(padded)
        g: Dict[str, Tensor] = {}
        off = 0
        for k, s, ps in zip(names, sizes, padded):
            g[k] = flat[off:off + s].view(p[k].shape)
            off += ps
        return g

    def backward(self, p: Dict[str, Tensor], sv: dict, dlogits: Tensor, flat: Optional[Tensor] = None, *,
                 g: Optional[Dict[str, Tensor]] = None, front_bwd=None, layer_hook=None) -> Dict[str, Tensor]:
        """dlogits: bf16 [B*T*S, nv*vs] (gradient of the out_x_proj output). Returns the gradient dict of alloc_grads
        (`g`, if given, is used as is). `front_bwd(dx, dact)` replaces the token-embedding backward (mar.py).
        `layer_hook = (layers, fn)`: fn(i) is called on the main stream right after the backward of layer i has been
        enqueued, for i in `layers` — every SHARED gradient of layers >= i is then final (train.py overlaps their all-reduce
        with the rest of the backward, and splits CUDA-graph capture there; the side stream is joined around the call)."""
        d: Dims = sv["dims"]
        dom = sv["dom"]
        Wt = self.weights.trans
        B, T, S, n, M, N = d.B, d.T, d.S, d.n, d.M, d.N
        dev = dlogits.device
        if g is None:
            g = self.alloc_grads(p, d, dom, sv["has_actions"], dev, flat)
        drop = sv.get("drop")

        def g2(name):  # gradient viewed as a matrix / vector
            return g[name]

        # ---- head
        assert d.readout_alpha == 1.0, "readout scaling != 1 is not implemented for training"
        ops.gemm_wgrad(dlogits, sv["ah"], g2("out_x_proj.weight"))
        ops.colsum_bf16(dlogits, g2("out_x_proj.bias"))
        dxh = ops.gemm_nt(dlogits, Wt["out_x_proj.weight"], EPI_RESID)
        dx = ops.rows_scatter(dxh, M, S, n) if d.A else dxh
        dact = torch.zeros(M, C, device=dev, dtype=torch.float32) if sv["has_actions"] else None

        main = torch.cuda.current_stream()
        side = self._side_stream(dev)
        if d.modulate:
            fork = torch.cuda.Event()
            fork.record(main)  # dact zero-fill and the gradient buffer memset precede the side-stream chain
            side.wait_event(fork)
        dy = None  # bf16 copy of dx whose column sums are already in the consumer's bias gradient
        side_keep = []
        qk = bool(self.cfg.qk_norm)
        for i in reversed(range(d.num_layers)):
            L = sv["layers"][i]
            lp = f"decoder.layers.{i}."
            # The seven weight gradients of the block are collected and contracted by ONE persistent launch at the end of
            # the block (ops.gemm_wgrad_grouped): nothing in the backward chain waits for them, their operands (the dy / dz /
            # dqkv of each stage and the saved activations) are not touched again, and one launch of ~3 long work items per
            # CTA replaces seven single-wave launches that are mostly prologue and atomics.
            wg = []

            def wgrad(G, X, dW, wg=wg):
                wg.append((G, X, dW))
            # ---- MLP
            if drop is not None:  # x4 = x3 + drop(fc2(drop(gelu(z)))): the keep masks are regenerated from the seeds
                dy = ops.dropout_cast_bf16(dx, drop[0], drop[1] + 2 * i + 1, drop[2])
                if lp + "mlp.fc2.bias" in g:
                    ops.colsum_bf16(dy, g[lp + "mlp.fc2.bias"])
            elif dy is None:
                dy = ops.cast_colsum(dx, g.get(lp + "mlp.fc2.bias"))
            wgrad(dy, L["h"], g2(lp + "mlp.fc2.weight"))
            dz = ops.gemm_nt(dy, Wt[lp + "mlp.fc2.weight"], EPI_DGELU, aux=L["z"],
                             colsum=g.get(lp + "mlp.fc1.bias") if drop is None else None)
            if drop is not None:  # the GELU-output dropout mask applies to dz: fc1's bias gradient is taken after it
                ops.dropout_bf16_(dz, drop[0], drop[1] + 2 * i, drop[2])
                if lp + "mlp.fc1.bias" in g:
                    ops.colsum_bf16(dz, g[lp + "mlp.fc1.bias"])
            wgrad(dz, L["a2"], g2(lp + "mlp.fc1.weight"))
            da2 = ops.gemm_nt(dz, Wt[lp + "mlp.fc1.weight"], EPI_BF16)
            if qk:
                dy = ops.ln_bwd(da2, None, None, 0, dx, want_next=True, colsum_next=g.get(lp + "temporal_attn.proj.bias"))
            else:
                dy = ops.ln_bwd(da2, L["x3"], L["st2"], 1, dx, gamma=p[lp + "norm2.weight"], dgamma=g2(lp + "norm2.weight"),
                                dbeta=g2(lp + "norm2.bias"), want_next=True, colsum_next=g.get(lp + "temporal_attn.proj.bias"))
            # ---- temporal attention
            wgrad(dy, L["att_t"], g2(lp + "temporal_attn.proj.weight"))
            datt = ops.gemm_nt(dy, Wt[lp + "temporal_attn.proj.weight"], EPI_BF16)
            dqkv = ops.attn_temporal_bwd(L["qkv_t"], L["att_t"], datt, L["lse_t"], B, T, n, d.heads, d.scale)
            if qk:
                ops.qk_norm_bwd(L["qkv_t_raw"], p[lp + "temporal_attn.norm.weight"], dqkv, g2(lp + "temporal_attn.norm.weight"),
                                g2(lp + "temporal_attn.norm.bias"))
            wgrad(dqkv, L["at"], g2(lp + "temporal_attn.qkv.weight"))
            if lp + "temporal_attn.qkv.bias" in g:
                ops.colsum_bf16(dqkv, g2(lp + "temporal_attn.qkv.bias"))
            # dx += dqkv . W; its bf16 copy and column sums (operand and bias gradient of the stage below) come out of
            # the same epilogue
            dy = torch.empty(N, C, device=dev, dtype=torch.bfloat16)
            nb = g2(lp + f"action_projectors.{dom}.linear_out.bias") if d.modulate else g.get(lp + "spatial_attn.proj.bias")
            ops.gemm_nt(dqkv, Wt[lp + "temporal_attn.qkv.weight"], EPI_RESID, resid=dx, out=dx, out2=dy, colsum=nb)
            # ---- modulate
            if d.modulate:
                ap = lp + f"action_projectors.{dom}."
                wgrad(dy, L["am"], g2(ap + "linear_out.weight"))
                dam = ops.gemm_nt(dy, Wt[ap + "linear_out.weight"], EPI_BF16)
                dmod = torch.zeros(M, 2 * C, device=dev, dtype=torch.float32)
                dy = ops.ln_bwd(dam, L["x1"], L["stm"], 2, dx, mod=L["mod"], rows_per_group=n, dmod=dmod, want_next=True,
                                colsum_next=g.get(lp + "spatial_attn.proj.bias"))
                # adaLN_modulation backward: a chain of one-tile kernels -> side stream
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    dmod_bf = ops.cast_bf16(dmod)
                    ops.gemm_wgrad(dmod_bf, L["hmod"], g2(ap + "adaLN_modulation.2.weight"))
                    ops.colsum_f32(dmod, g2(ap + "adaLN_modulation.2.bias"))
                    dzm = ops.gemm_nt(dmod_bf, Wt[ap + "adaLN_modulation.2.weight"], EPI_DSILU, aux=L["zmod"],
                                      colsum=g2(ap + "adaLN_modulation.0.bias"))
                    ops.gemm_wgrad(dzm, sv["c_bf"], g2(ap + "adaLN_modulation.0.weight"))
                    ops.gemm_nt(dzm, Wt[ap + "adaLN_modulation.0.weight"], EPI_RESID, resid=dact, out=dact)
                    side_keep.append((dmod, dmod_bf, dzm))  # alive until the side stream has been joined
            elif d.additive:
                ops.group_colsum(dx, dact, n)
            # ---- spatial attention
            wgrad(dy, L["att_s"], g2(lp + "spatial_attn.proj.weight"))
            # delta = rowsum(dO * O) per (token, head) comes out of the epilogue of the GEMM that produces dO (a head's 32
            # channels are one epilogue chunk), so the attention kernel's prologue reads 8 bytes per row instead of 128
            delta = torch.empty(N, d.heads, device=dev, dtype=torch.float32)
            datt = ops.gemm_nt(dy, Wt[lp + "spatial_attn.proj.weight"], EPI_BF16, aux=L["att_s"], rowdot=delta)
            dqkv = ops.attn_spatial_bwd(L["qkv_s"], None, datt, L["lse"], M, n, d.heads, d.scale, delta=delta)
            if qk:
                ops.qk_norm_bwd(L["qkv_s_raw"], p[lp + "spatial_attn.norm.weight"], dqkv, g2(lp + "spatial_attn.norm.weight"),
                                g2(lp + "spatial_attn.norm.bias"))
            wgrad(dqkv, L["a1"], g2(lp + "spatial_attn.qkv.weight"))
            if lp + "spatial_attn.qkv.bias" in g:
                ops.colsum_bf16(dqkv, g2(lp + "spatial_attn.qkv.bias"))
            da1 = ops.gemm_nt(dqkv, Wt[lp + "spatial_attn.qkv.weight"], EPI_BF16)
            nxt = f"decoder.layers.{i - 1}.mlp.fc2.bias" if i > 0 else None
            if qk:
                dy = ops.ln_bwd(da1, None, None, 0, dx, want_next=i > 0, colsum_next=g.get(nxt) if nxt else None)
            else:
                dy = ops.ln_bwd(da1, L["x0"], L["st1"], 1, dx, gamma=p[lp + "norm1.weight"], dgamma=g2(lp + "norm1.weight"),
                                dbeta=g2(lp + "norm1.bias"), want_next=i > 0, colsum_next=g.get(nxt) if nxt else None)
            ops.gemm_wgrad_grouped(wg)
            sv["layers"][i] = None  # release this layer's activations
            if layer_hook is not None and i in layer_hook[0]:
                if d.modulate:  # a capture segment must end with every forked stream joined
                    join = torch.cuda.Event()
                    join.record(side)
                    main.wait_event(join)
                layer_hook[1](i)
                main = torch.cuda.current_stream()
                if d.modulate:
                    fork = torch.cuda.Event()
                    fork.record(main)
                    side.wait_event(fork)
        if d.modulate:
            join = torch.cuda.Event()
            join.record(side)
            main.wait_event(join)
            side_keep.clear()

        # ---- embedding / positional / action-token gradients
        if front_bwd is not None:
            front_bwd(dx, dact if d.A else None)
        else:
            ops.embed_bwd(sv["ids"], dx, sv["pos_n"], B, T, S, d.A, d.vs, d.mask_id,
                          g["token_embed.factored_embeds.0.weight"], g.get("token_embed.factored_embeds.1.weight"),
                          g["token_embed.mask_token_embed"], dact if d.A else None, g["pos_embed_TSC"])
        # ---- action stem backward
        if sv["has_actions"] and (d.A or d.modulate or d.additive):
            q = f"action_mlp.{dom}.model."
            dact_bf = ops.cast_bf16(dact)
            ops.gemm_wgrad(dact_bf, sv["h1n"], g2(q + "3.weight"))
            ops.colsum_f32(dact, g2(q + "3.bias"))
            dh1n = ops.gemm_nt(dact_bf, Wt[q + "3.weight"], EPI_RESID)
            dh1 = ops.ln_relu_bwd(dh1n, sv["h1"], sv["st_stem"], p[q + "1.weight"], p[q + "1.bias"], g2(q + "1.weight"),
                                  g2(q + "1.bias"))
            ops.colsum_f32(dh1, g2(q + "0.bias"))
            kpad = sv["a_prep"].shape[1]
            dw0 = torch.zeros(C, kpad, device=dev, dtype=torch.float32)
            ops.gemm_wgrad(ops.cast_bf16(dh1), sv["a_prep"], dw0)
            g[q + "0.weight"].copy_(dw0[:, : sv["da"]])
        return g
