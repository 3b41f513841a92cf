"""STMaskGIT: drop-in for the reference model API (hma/model/st_mask_git.py:150-769) on top of the
B200 CUDA path. Same constructor/config, same method signatures and argument meaning, same
parameter names and shapes (so `state_dict()` / `load_state_dict()` / `save_pretrained` interoperate
with reference checkpoints), same error behaviour (asserts, in-place prompt
update). The sub-modules below are parameter containers only: no torch op computes anything on the
hot path, which is scheduled by hma_b200.engine and executed by libhma_b200.so. There is no CPU
path: calling the model without a CUDA device raises.
"""
import math
from typing import Optional

import torch
import torch.nn as nn

try:  # same (de)serialisation mixin as the reference (st_mask_git.py:8,150)
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass

try:
    from transformers.utils import ModelOutput
except Exception:  # pragma: no cover
    class ModelOutput(dict):  # type: ignore
        def __getattr__(self, k):
            return self[k]

from . import ops
from .config import GenieConfig
from .engine import SMOOTHING, Engine


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's names (never called)
# ------------------------------------------------------------------------------------------------
def _xavier(m: nn.Module, gain: float) -> None:
    for mod in m.modules():
        if isinstance(mod, nn.Linear):
            nn.init.xavier_uniform_(mod.weight, gain=gain)
            if mod.bias is not None:
                nn.init.zeros_(mod.bias)


class _Attention(nn.Module):  # attention.py:10-35
    def __init__(self, cfg: GenieConfig):
        super().__init__()
        d = cfg.d_model
        self.qkv = nn.Linear(d, 3 * d, bias=cfg.qkv_bias)
        self.proj = nn.Linear(d, d, bias=cfg.proj_bias)
        if cfg.qk_norm:
            self.norm = nn.LayerNorm(d // cfg.num_heads, eps=1e-5)


class _Mlp(nn.Module):  # st_transformer.py:9-22
    def __init__(self, cfg: GenieConfig):
        super().__init__()
        hidden = int(cfg.d_model * cfg.mlp_ratio)
        self.fc1 = nn.Linear(cfg.d_model, hidden, bias=cfg.mlp_bias)
        self.fc2 = nn.Linear(hidden, cfg.d_model, bias=cfg.mlp_bias)


class _Block(nn.Module):  # st_transformer.py:30-77
    def __init__(self, cfg: GenieConfig):
        super().__init__()
        self.norm1 = nn.Identity() if cfg.qk_norm else nn.LayerNorm(cfg.d_model, eps=1e-5)
        self.spatial_attn = _Attention(cfg)
        self.temporal_attn = _Attention(cfg)
        self.norm2 = nn.Identity() if cfg.qk_norm else nn.LayerNorm(cfg.d_model, eps=1e-5)
        self.mlp = _Mlp(cfg)
        self.action_projectors = None


class _Decoder(nn.Module):  # st_transformer.py:117-168
    def __init__(self, cfg: GenieConfig):
        super().__init__()
        self.layers = nn.ModuleList([_Block(cfg) for _ in range(cfg.num_layers)])
        _xavier(self, 0.1)


class _ModulateLayer(nn.Module):  # st_mask_git.py:51-64
    def __init__(self, d: int):
        super().__init__()
        self.linear_out = nn.Linear(d, d, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.Linear(d, d), nn.SiLU(), nn.Linear(d, 2 * d, bias=True))
        _xavier(self, 0.1)


class _ActionMLP(nn.Module):  # st_mask_git.py:90-98
    def __init__(self, d_action: int, d: int):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(d_action, d), nn.LayerNorm(d), nn.ReLU(), nn.Linear(d, d))
        _xavier(self, 0.01)


class _ActionStat(nn.Module):  # st_mask_git.py:128-147
    def __init__(self, info):
        super().__init__()
        self.register_buffer("mean", torch.tensor(info[0], dtype=torch.float32))
        self.register_buffer("std", torch.tensor(info[1], dtype=torch.float32))

    def unnormalize(self, actions):
        d = self.mean.numel()
        b, t, D = actions.shape
        a = actions.reshape(b, t, D // d, d) * (self.std + 1e-10) + self.mean
        return a.reshape(b, t, D)


class _FactorizedEmbedding(nn.Module):  # factorization_utils.py:6-29
    def __init__(self, cfg: GenieConfig):
        super().__init__()
        self.factored_embeds = nn.ParameterList(
            [nn.Embedding(cfg.factored_vocab_size, cfg.d_model) for _ in range(cfg.num_factored_vocabs)])
        self.mask_token_embed = nn.Parameter(torch.zeros(1, cfg.d_model))


def cosine_schedule(u: float) -> float:  # st_mask_git.py:116-125
    return math.cos(u * math.pi / 2)


class _LazyParams:
    """Read-only name -> tensor view of a module's parameters and buffers for the inference paths. The name table
    is built once (a model with 40 action domains has ~8000 parameters; rebuilding a dict of detached tensors per
    call cost ~20 ms); tensors are detached at lookup time, so in-place updates, .to() moves and the autograd
    version counter (which tells the engine when to refresh its bf16 weight copies) are always current."""

    def __init__(self, module: nn.Module):
        self._t = dict(module.named_parameters())
        self._t.update(dict(module.named_buffers()))

    def __getitem__(self, k):
        return self._t[k].detach()

    def get(self, k, default=None):
        t = self._t.get(k)
        return default if t is None else t.detach()

    def __contains__(self, k):
        return k in self._t


# ------------------------------------------------------------------------------------------------
# autograd bridge: one Function for logits + loss, so DDP / optimizers see ordinary .grad tensors
# ------------------------------------------------------------------------------------------------
class _ForwardLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, names, ids, labels, actions, dom, dims, drop, *params):
        p = dict(model._buffers_dict())
        p.update({k: t.detach() for k, t in zip(names, params)})
        eng: Engine = model._engine
        logits, sv = eng.forward(p, ids, actions, dom, dims, training=True, drop=drop)
        loss_acc, lse, sums = ops.ce_fwd(logits, labels, ids, dims.B, dims.T, dims.S, dims.nv, dims.vs, dims.mask_id,
                                         SMOOTHING)
        ctx.model, ctx.names, ctx.p, ctx.sv = model, names, p, sv
        ctx.ce = (logits, labels, ids, lse, sums)
        loss, acc = loss_acc[0], loss_acc[1]
        ctx.mark_non_differentiable(acc, logits)
        return loss, acc, logits

    @staticmethod
    def backward(ctx, dloss, _dacc, _dlogits):
        model, names, p, sv = ctx.model, ctx.names, ctx.p, ctx.sv
        d = sv["dims"]
        logits, labels, ids, lse, sums = ctx.ce
        dl = dloss.detach().to(torch.float32).reshape(1).contiguous()
        dlogits = ops.ce_bwd(logits, labels, ids, d.B, d.T, d.S, d.nv, d.vs, d.mask_id, SMOOTHING, lse, sums, dl)
        grads = model._engine.backward(p, sv, dlogits)
        ctx.sv = ctx.ce = None
        return (None,) * 8 + tuple(grads.get(k) for k in names)


class _Logits(torch.autograd.Function):
    """compute_logits as a differentiable call (st_mask_git.py:632-686 is an ordinary autograd forward in the reference):
    the trunk's saved activations ride on the Function; backward takes d(logits) in whatever layout autograd hands back."""

    @staticmethod
    def forward(ctx, model, names, ids, actions, dom, dims, drop, skip_norm, *params):
        p = dict(model._buffers_dict())
        p.update({k: t.detach() for k, t in zip(names, params)})
        logits, sv = model._engine.forward(p, ids, actions, dom, dims, training=True, skip_normalization=skip_norm, drop=drop)
        ctx.model, ctx.names, ctx.p, ctx.sv = model, names, p, sv
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model, names, p, sv = ctx.model, ctx.names, ctx.p, ctx.sv
        grads = model._engine.backward(p, sv, dlogits.to(torch.bfloat16).contiguous())
        ctx.sv = None
        return (None,) * 8 + tuple(grads.get(k) for k in names)


class _VideoLoss(torch.autograd.Function):
    """compute_video_loss_and_acc (st_mask_git.py:603-630) on logits rows [B*T*S, nv*vs]: factorised cross entropy with
    label smoothing over the masked positions of frames >= 1, and its gradient w.r.t. the logits."""

    @staticmethod
    def forward(ctx, rows, labels, ids, B, T, S, nv, vs, mask_id):
        la, lse, sums = ops.ce_fwd(rows, labels, ids, B, T, S, nv, vs, mask_id, SMOOTHING)
        ctx.save_for_backward(rows, labels, ids, lse, sums)
        ctx.dims = (B, T, S, nv, vs, mask_id)
        loss, acc = la[0], la[1]
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, dloss, _dacc):
        rows, labels, ids, lse, sums = ctx.saved_tensors
        B, T, S, nv, vs, mask_id = ctx.dims
        dl = dloss.detach().to(torch.float32).reshape(1).contiguous()
        d = ops.ce_bwd(rows, labels, ids, B, T, S, nv, vs, mask_id, SMOOTHING, lse, sums, dl)
        return (d.to(rows.dtype),) + (None,) * 8


class STMaskGIT(nn.Module, PyTorchModelHubMixin):
    # "incremental": frame-incremental decode with a temporal K/V cache (hma_b200/decode.py);
    # "full": the reference algorithm (whole-window compute_logits per MaskGIT step, st_mask_git.py:384,394)
    decode_algorithm = "incremental"
    decode_cuda_graphs = True

    def __init__(self, config: GenieConfig):
        super().__init__()
        self.h = self.w = math.isqrt(config.S)
        assert self.h ** 2 == config.S, "Expected S to be square"
        self.decoder = _Decoder(config)
        self.pos_embed_TSC = nn.Parameter(torch.zeros(1, config.T, config.S + config.action_token_size, config.d_model))
        self.mask_token_id = config.image_vocab_size
        self.seq_len = config.S
        self.relevant_action_mask = None
        self._build_io(config)
        self.config = config
        self.action_mask_tokens = nn.Parameter(torch.zeros(1, config.T, 1, config.d_model))
        self._engine = self._make_engine(config)
        self._sessions = {}
        self._lazy = None
        self._arena = None
        if (config.init_actions or config.use_actions) and config.action_domains is not None:
            self.init_action_projectors(config.action_domains, config.d_actions, config.action_stats, config.action_network)

    def _build_io(self, config) -> None:
        """Token embedding and readout (st_mask_git.py:184-192); STMAR replaces both (st_mar.py:56-66)."""
        self.token_embed = _FactorizedEmbedding(config)
        self.out_x_proj = nn.Linear(config.d_model, config.factored_vocab_size * config.num_factored_vocabs)
        _xavier(self.out_x_proj, 0.01) if config.use_mup else None

    def _make_engine(self, config) -> Engine:
        return Engine(config)

    # ---------------------------------------------------------------- construction (st_mask_git.py:201-251)
    def init_action_projectors(self, domains, d_actions, action_stats, action_network: str = "mlp", use_diffusion: bool = False):
        assert len(domains) == len(d_actions) == len(action_stats), \
            f"{len(domains)=} {len(d_actions)=} {len(action_stats)=}"
        cfg = self.config
        cfg.init_actions = True
        cfg.action_domains, cfg.d_actions, cfg.action_stats = list(domains), list(d_actions), action_stats
        cfg.action_network = action_network
        self._engine.check(cfg)
        dev = self.pos_embed_TSC.device
        self.action_preprocessor = nn.ModuleDict()
        self.action_mlp = nn.ModuleDict()
        self.action_out_projectors = nn.ModuleDict()
        for dom, da, stat in zip(domains, d_actions, action_stats):
            self.action_preprocessor[dom] = _ActionStat(stat)
            self.action_mlp[dom] = _ActionMLP(da, cfg.d_model)
            if not use_diffusion:
                self.action_out_projectors[dom] = nn.Linear(cfg.d_model, da)
        for layer in self.decoder.layers:
            layer.action_projectors = nn.ModuleDict()
            for dom in domains:
                # same precedence as st_mask_git.py:240-251: "mlp" (additive, Identity projector) wins over "modulate"
                if "modulate" in action_network and "mlp" not in action_network and "cross_attention" not in action_network:
                    layer.action_projectors[dom] = _ModulateLayer(cfg.d_model)
                else:
                    layer.action_projectors[dom] = nn.Identity()
        self.to(dev)
        self._lazy = None

    def _save_pretrained(self, save_directory) -> None:
        """save_pretrained (the reference's checkpoint call, train_multi.py:310-321). After a TrainStep has re-homed the
        parameters into its flat arena they are views of one storage, which safetensors refuses to serialise: clone
        them first (same keys, shapes and values; the file is identical to one written before re-homing)."""
        if getattr(self, "_arena", None) is None:
            return super()._save_pretrained(save_directory)
        from pathlib import Path

        from huggingface_hub import constants
        from safetensors.torch import save_file
        tensors = {k: v.detach().clone().contiguous() for k, v in self.state_dict().items()}
        save_file(tensors, str(Path(save_directory) / constants.SAFETENSORS_SINGLE_FILE), metadata={"format": "pt"})

    def _inference_params(self) -> _LazyParams:
        if self._lazy is None:
            self._lazy = _LazyParams(self)
        return self._lazy

    # ---------------------------------------------------------------- plumbing
    def _buffers_dict(self):
        return {k: b for k, b in self.named_buffers()}

    def _params_dict(self):
        return {k: v for k, v in self.named_parameters()}

    def _require_cuda(self, t: torch.Tensor) -> None:
        if not t.is_cuda or not self.pos_embed_TSC.is_cuda:
            raise RuntimeError("hma_b200.STMaskGIT runs on a CUDA device only (no CPU path exists); "
                               "move the model and its inputs to cuda")

    def _hw(self, kwargs):
        h, w = self.h, self.w
        if "h" in kwargs:
            assert "w" in kwargs
            h, w = kwargs["h"][0], kwargs["w"][0]
        return int(h), int(w)

    def _domain0(self, domain, action_ids):
        if action_ids is None:
            return None
        dom = domain[0] if not isinstance(domain, str) else domain  # st_mask_git.py:648,669 (only [0] is used)
        if dom not in self.action_mlp:
            raise KeyError(f"unknown action domain {dom!r}")
        return dom

    def _logits_nograd(self, x_THW, action_ids, domain, kwargs):
        B, T, H, W = x_THW.shape
        dom = self._domain0(domain, action_ids)
        if action_ids is not None:
            assert action_ids.shape[1] == T, "action_ids must provide one action vector per frame"
        d = self._engine.dims(B, T, H * W, action_ids is not None)
        p = self._inference_params()
        ids = x_THW.reshape(B, T, H * W).contiguous()
        logits, _ = self._engine.forward(p, ids, action_ids, dom, d, training=False,
                                         skip_normalization=kwargs.get("skip_normalization", False))
        return logits, d

    @staticmethod
    def _as_CTHW(logits, B, T, H, W):
        return logits.view(B, T, H, W, -1).permute(0, 4, 1, 2, 3)  # "B T (H W) C -> B C T H W" as a view

    # ---------------------------------------------------------------- st_mask_git.py:632-686
    def compute_logits(self, x_THW: torch.Tensor, action_ids: Optional[torch.Tensor] = None, domain=None, **kwargs):
        self._require_cuda(x_THW)
        B, T, H, W = x_THW.shape
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # differentiable, like the reference's (forward() is the faster training call: it fuses head, loss and backward)
            dom = self._domain0(domain, action_ids)
            if action_ids is not None:
                assert action_ids.shape[1] == T, "action_ids must provide one action vector per frame"
            d = self._engine.dims(B, T, H * W, action_ids is not None)
            named = [(k, v) for k, v in self.named_parameters()]
            drop = None
            if self.training and self.config.mlp_drop > 0.0:
                drop = (float(self.config.mlp_drop), int(torch.randint(0, 2 ** 62, ()).item()), None)
            logits = _Logits.apply(self, [k for k, _ in named], x_THW.reshape(B, T, H * W).contiguous(), action_ids, dom, d, drop,
                                   bool(kwargs.get("skip_normalization", False)), *[v for _, v in named])
        else:
            logits, _ = self._logits_nograd(x_THW, action_ids, domain, kwargs)
        return self._as_CTHW(logits, B, T, H, W), None

    # ---------------------------------------------------------------- st_mask_git.py:603-630
    def compute_video_loss_and_acc(self, logits_CTHW, targets_THW, relevant_mask_THW):
        """Same contract as the reference: logits [B, nv*vs, T, H, W], targets [B, T*H*W], mask [B, T-1, H, W].
        (forward() does not call this; it fuses the loss with the head.) Differentiable w.r.t. the logits."""
        B, Cv, T, H, W = logits_CTHW.shape
        cfg = self.config
        rows = logits_CTHW.permute(0, 2, 3, 4, 1).reshape(B * T * H * W, Cv).float().contiguous()
        ids = torch.zeros(B, T, H * W, dtype=torch.long, device=rows.device)
        ids[:, 1:][relevant_mask_THW.reshape(B, T - 1, H * W).bool()] = self.mask_token_id
        labels = targets_THW.reshape(B, T * H * W).contiguous()
        return _VideoLoss.apply(rows, labels, ids, B, T, H * W, cfg.num_factored_vocabs, cfg.factored_vocab_size,
                                self.mask_token_id)

    # ---------------------------------------------------------------- st_mask_git.py:688-735
    def forward(self, input_ids, labels, action_ids=None, domain="default", **kwargs):
        self._require_cuda(input_ids)
        T = self.config.T
        H, W = self._hw(kwargs)
        B = input_ids.shape[0]
        ids = input_ids.reshape(B, T, H * W).contiguous()
        labels = labels.reshape(B, T * H * W).contiguous()
        dom = self._domain0(domain, action_ids)
        if action_ids is not None:
            assert action_ids.shape[1] == T, "action_ids must provide one action vector per frame"
        d = self._engine.dims(B, T, H * W, action_ids is not None)
        if action_ids is not None:
            # st_mask_git.py:703-710: the reference draws an action-drop mask from the CPU generator on every forward.
            # It only matters with jointly_predict_actions (not implemented), but the two draws are kept so that the
            # CPU RNG stream of a training script stays in step with the reference.
            drop_ratio = torch.rand(len(action_ids), 1, 1)
            self.relevant_action_mask = (torch.rand(len(action_ids), T, 1) < drop_ratio).unsqueeze(-1)
        if torch.is_grad_enabled():
            named = [(k, v) for k, v in self.named_parameters()]
            names = [k for k, _ in named]
            drop = None
            if self.training and self.config.mlp_drop > 0.0:  # nn.Dropout after GELU and after fc2 (st_transformer.py:24-27)
                drop = (float(self.config.mlp_drop), int(torch.randint(0, 2 ** 62, ()).item()), None)
            loss, acc, logits = _ForwardLoss.apply(self, names, ids, labels, action_ids, dom, d, drop, *[v for _, v in named])
        else:
            logits, _ = self._logits_nograd(ids.view(B, T, H, W), action_ids, domain, kwargs)
            la, _, _ = ops.ce_fwd(logits, labels, ids, B, T, H * W, d.nv, d.vs, d.mask_id, SMOOTHING)
            loss, acc = la[0], la[1]
        return ModelOutput(loss=loss, acc=acc, logits=self._as_CTHW(logits, B, T, H, W))

    # ---------------------------------------------------------------- st_mask_git.py:331-335
    def init_mask(self, prompt_THW, t=1):
        return torch.zeros(prompt_THW.size(0), t * self.seq_len, dtype=torch.bool, device=prompt_THW.device)

    # ---------------------------------------------------------------- frame-incremental decode state
    def _decode_session(self, prompt_THW, n_ctx: int, action_ids, domain, kwargs):
        """A DecodeSession for this window shape with frames [0, n_ctx) of `prompt_THW` prefilled."""
        from .decode import DecodeSession
        B, T, H, W = prompt_THW.shape
        dom = self._domain0(domain, action_ids)
        if action_ids is not None:
            assert action_ids.shape[1] == T, "action_ids must provide one action vector per frame"
        key = (B, T, H * W, dom, action_ids.shape[-1] if action_ids is not None else 0, prompt_THW.device,
               self.decode_cuda_graphs)
        sess = self._sessions.get(key)
        if sess is None:
            self._sessions.clear()  # one live session: its K/V cache is the large allocation
            sess = DecodeSession(self, B, T, H * W, dom, key[4], prompt_THW.device, use_graphs=self.decode_cuda_graphs)
            self._sessions[key] = sess
        p = self._inference_params()
        sess.begin(p, prompt_THW, n_ctx, action_ids, kwargs.get("skip_normalization", False))
        return sess

    # ---------------------------------------------------------------- st_mask_git.py:337-467
    @torch.no_grad()
    def maskgit_generate(self, prompt_THW: torch.LongTensor, out_t: int, maskgit_steps: int = 1, temperature: float = 0.0,
                         unmask_mode: str = "random", action_ids=None, domain="default", **kwargs):
        self._require_cuda(prompt_THW)
        assert out_t, "maskgit_generate requires out_t > 0"
        assert torch.all(prompt_THW[:, out_t:] == self.mask_token_id), \
            f"when generating z{out_t}, frames {out_t} and later must be masked"
        if unmask_mode not in ("greedy", "random"):
            raise NotImplementedError(f"Expected `unmask_mode` to be one of ['greedy', 'random'], got {unmask_mode}")
        B, T, H, W = prompt_THW.shape
        S = H * W
        cfg = self.config
        nv, vs = cfg.num_factored_vocabs, cfg.factored_vocab_size
        if not prompt_THW.is_contiguous():
            raise ValueError("prompt_THW must be contiguous (it is updated in place)")
        frame = prompt_THW[:, out_t].view(B, S)
        unmasked = torch.zeros(B, S, dtype=torch.uint8, device=prompt_THW.device)
        orig_logits = None
        samples = None
        session = kwargs.pop("_session", None)
        if session is None and self.decode_algorithm == "incremental":
            session = self._decode_session(prompt_THW, out_t, action_ids, domain, kwargs)
        for step in range(maskgit_steps):
            if session is not None:
                # only frame out_t runs; context comes from the K/V cache (and the first step's logits from the pass that
                # committed the previous frame, when generate() asked for them)
                logits = session.step(frame, out_t, first=step == 0)
                lf = logits.view(B, S, nv * vs)
            else:
                logits, _ = self._logits_nograd(prompt_THW, action_ids, domain, kwargs)
                lf = logits.view(B, T, S, nv * vs)[:, out_t]  # strided view of frame out_t
            if orig_logits is None:
                orig_logits = lf.clone()
            noise = None
            if temperature > 1e-8:
                # Appendix C of SURVEY.md: one Exp(1) tensor [B*S, vs] per vocabulary half, high half first
                noise = torch.stack([torch.empty(B * S, vs, device=logits.device, dtype=torch.float32).exponential_(1)
                                     for _ in range(nv)])
            new, conf = ops.sample_tokens(lf, nv, vs, noise, temperature if noise is not None else 1.0)
            if step != maskgit_steps - 1:
                n = math.ceil(cosine_schedule((step + 1) / maskgit_steps) * S)
                if unmask_mode == "greedy":
                    keys = conf
                else:
                    keys = torch.rand(B, H, W, device=logits.device, dtype=torch.float32).view(B, S)
                samples = ops.rank_remask(keys, unmasked, new, frame, n, self.mask_token_id)
            else:
                samples = ops.rank_remask(None, unmasked, new, frame, -1, self.mask_token_id)
        factored = orig_logits.view(B, H, W, nv, vs).permute(0, 4, 3, 1, 2)  # B vs nv H W
        return samples.view(B, H, W), factored, None

    # ---------------------------------------------------------------- st_mask_git.py:253-329
    def generate(self, input_ids: torch.LongTensor, attention_mask, max_new_tokens: int, min_new_tokens: int = None,
                 return_logits: bool = False, return_with_actions: bool = False, maskgit_steps: int = 1,
                 temperature: float = 0.0, action_ids: torch.Tensor = None, domain: str = "default", **kwargs):
        assert min_new_tokens in (None, max_new_tokens), \
            "Expecting `min_new_tokens`, if specified, to match `max_new_tokens`."
        if "h" not in kwargs or "w" not in kwargs:
            # the reference raises UnboundLocalError here (st_mask_git.py:283-289); fail with a message instead
            raise TypeError("generate() requires h=[..] and w=[..] keyword arguments")
        if return_with_actions:
            raise NotImplementedError("return_with_actions needs jointly_predict_actions, which is not implemented")
        h, w = kwargs["h"][0], kwargs["w"][0]
        S = h * w
        n_new = max_new_tokens // S
        B = input_ids.size(0)
        inputs = input_ids.clone().reshape(B, -1, h, w)
        full = torch.cat([inputs, torch.full((B, n_new, h, w), self.mask_token_id, dtype=torch.long,
                                             device=input_ids.device)], dim=1).contiguous()
        all_logits = []
        session = None
        if self.decode_algorithm == "incremental" and n_new > 0:
            self._require_cuda(input_ids)
            session = self._decode_session(full, inputs.size(1), action_ids, domain, kwargs)
        last = inputs.size(1) + n_new - 1
        for t in range(inputs.size(1), inputs.size(1) + n_new):
            sample_HW, fl, _ = self.maskgit_generate(full, t, maskgit_steps=maskgit_steps, temperature=temperature,
                                                     action_ids=action_ids, domain=domain, _session=session, **kwargs)
            full[:, t] = sample_HW
            all_logits.append(fl)
            if session is not None and t != last:
                # the finished frame joins the context of the next one; the same pass runs the next frame's first step
                session.commit(full[:, t], t, prefetch_next=True)
        tokens = full.reshape(B, -1)
        if return_logits:
            return tokens, torch.stack(all_logits, dim=3)
        return tokens
