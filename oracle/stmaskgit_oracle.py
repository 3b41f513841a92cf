"""CPU oracle for the ST-MaskGIT hot path: a plain fp32 PyTorch restatement of the reference.

TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
`--impl reference` legs may import this module; the product (hma_b200/) never does.

Parity pin: the reference ships no tests or golden vectors, so this restatement is
pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the authoring container by
oracle/make_golden.py (which imports /root/reference with the two import shims in
oracle/ref_shims/) and committed under tests/golden/. tests/test_oracle.py checks the restatement
against those fixtures, and against the live reference when /root/reference is present.

Everything is functional: weights come from a state_dict with the reference's key layout
(SURVEY.md Appendix A), so the same tensors can be loaded into the reference, this oracle and the
CUDA model. Each function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class OracleConfig:
    """Subset of hma/config.py:8-61 that the path reads."""

    num_layers: int
    num_heads: int
    d_model: int
    T: int = 12
    S: int = 256
    image_vocab_size: int = 262144
    num_factored_vocabs: int = 2
    use_mup: bool = False
    qkv_bias: bool = False
    proj_bias: bool = True
    qk_norm: bool = False
    mlp_ratio: float = 4.0
    mlp_bias: bool = True
    action_network: str = "concat+modulate"
    action_token_size: int = 64
    jointly_predict_actions: bool = False
    jointly_predict_states: bool = True
    action_domains: Optional[List[str]] = None
    d_actions: Optional[List[int]] = None

    @property
    def factored_vocab_size(self) -> int:  # config.py:77-81
        return round(self.image_vocab_size ** (1.0 / self.num_factored_vocabs))

    @property
    def mask_token_id(self) -> int:  # st_mask_git.py:181
        return self.image_vocab_size

    @property
    def attn_scale(self) -> float:  # attention.py:27
        hd = self.d_model // self.num_heads
        return 8.0 / hd if self.use_mup else hd ** -0.5


# --------------------------------------------------------------------------------------------
# embedding (factorization_utils.py:31-54, 57-68)
# --------------------------------------------------------------------------------------------
def factorize_token_ids(ids: Tensor, nv: int, vs: int) -> Tensor:
    """factorization_utils.py:57-68: digit k of the id in base `vs` (little endian)."""
    return torch.stack([(ids // (vs ** k)) % vs for k in range(nv)], dim=-1)


def token_embed(ids_BTS: Tensor, sd: SD, cfg: OracleConfig) -> Tensor:
    """factorization_utils.py:31-54: sum of per-factor embeddings; mask id -> mask_token_embed."""
    nv, vs = cfg.num_factored_vocabs, cfg.factored_vocab_size
    is_mask = ids_BTS == cfg.mask_token_id
    safe = torch.where(is_mask, torch.zeros_like(ids_BTS), ids_BTS)
    digits = factorize_token_ids(safe, nv, vs)
    e = sum(F.embedding(digits[..., k], sd[f"token_embed.factored_embeds.{k}.weight"]) for k in range(nv))
    return torch.where(is_mask[..., None], sd["token_embed.mask_token_embed"].expand_as(e), e)


# --------------------------------------------------------------------------------------------
# action stem (st_mask_git.py:134-138 ActionStat, 90-102 BasicMLP)
# --------------------------------------------------------------------------------------------
def action_stem(a_BTD: Tensor, sd: SD, dom: str, skip_normalization: bool = False) -> Tensor:
    if not skip_normalization:
        mean, std = sd[f"action_preprocessor.{dom}.mean"], sd[f"action_preprocessor.{dom}.std"]
        d = mean.numel()
        B, T, D = a_BTD.shape
        a = a_BTD.reshape(B, T, D // d, d)
        a_BTD = ((a - mean) / (std + 1e-10)).reshape(B, T, D)
    p = f"action_mlp.{dom}.model."
    h = F.linear(a_BTD, sd[p + "0.weight"], sd[p + "0.bias"])
    h = F.layer_norm(h, (h.shape[-1],), sd[p + "1.weight"], sd[p + "1.bias"], 1e-5)
    return F.linear(F.relu(h), sd[p + "3.weight"], sd[p + "3.bias"])


# --------------------------------------------------------------------------------------------
# attention (attention.py:37-61, the in-repo restatement of the xformers call at :139-155)
# --------------------------------------------------------------------------------------------
# "math": BasicSelfAttention (attention.py:37-61), what the reference runs with XFORMERS_DISABLED — the parity pin.
# "sdpa": the fused-attention variant (attention.py:139-155 calls xformers' memory_efficient_attention, i.e. FlashAttention;
#         torch's scaled_dot_product_attention is the same algorithm) — only used by bench.py's GPU reference leg.
ATTENTION_IMPL = "math"
CHECKPOINT_LAYERS = False  # torch.utils.checkpoint around every ST block (full-size GPU parity tests: caps activation memory)


def self_attention(x_BNC: Tensor, sd: SD, prefix: str, cfg: OracleConfig, causal: bool) -> Tensor:
    B, N, C = x_BNC.shape
    H = cfg.num_heads
    qkv = F.linear(x_BNC, sd[prefix + "qkv.weight"], sd.get(prefix + "qkv.bias"))
    q, k, v = qkv.reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    if cfg.qk_norm:  # attention.py:32-35,43-48: one LayerNorm(head_dim) shared by q and k
        w, b = sd[prefix + "norm.weight"], sd[prefix + "norm.bias"]
        q = F.layer_norm(q, (C // H,), w, b, 1e-5)
        k = F.layer_norm(k, (C // H,), w, b, 1e-5)
    if ATTENTION_IMPL == "sdpa":
        o = F.scaled_dot_product_attention(q.to(v.dtype), k.to(v.dtype), v, is_causal=causal, scale=cfg.attn_scale)
        return F.linear(o.transpose(1, 2).reshape(B, N, C), sd[prefix + "proj.weight"], sd.get(prefix + "proj.bias"))
    s = (q * cfg.attn_scale) @ k.transpose(-2, -1)
    if causal:  # attention.py:51-55
        keep = torch.ones(N, N, dtype=torch.bool, device=x_BNC.device).tril()
        s = s.masked_fill(~keep, -torch.finfo(s.dtype).max)
    o = (s.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[prefix + "proj.weight"], sd.get(prefix + "proj.bias"))


# --------------------------------------------------------------------------------------------
# per-layer action conditioning (st_mask_git.py:51-76 ModulateLayer)
# --------------------------------------------------------------------------------------------
def adaln_shift_scale(a_BTC: Tensor, sd: SD, prefix: str):
    """st_mask_git.py:61-63,73: Linear -> SiLU -> Linear, chunk(2) = (shift, scale)."""
    h = F.silu(F.linear(a_BTC, sd[prefix + "adaLN_modulation.0.weight"], sd[prefix + "adaLN_modulation.0.bias"]))
    m = F.linear(h, sd[prefix + "adaLN_modulation.2.weight"], sd[prefix + "adaLN_modulation.2.bias"])
    return m.chunk(2, dim=-1)


def modulate_layer(x_BTSC: Tensor, a_BTC: Tensor, sd: SD, prefix: str) -> Tensor:
    """st_mask_git.py:66-76 on the (B,T,S,C) view: the reference reshapes to (b,s,t,d) and broadcasts
    the per-(b,t) shift/scale over s; the slice `c[:, None, :x_shape[2]]` keeps every frame because
    x_shape[2] is d_model >= T."""
    C = x_BTSC.shape[-1]
    shift, scale = adaln_shift_scale(a_BTC, sd, prefix)
    xn = F.layer_norm(x_BTSC, (C,), None, None, 1e-6)
    y = xn * (1 + scale[:, :, None, :]) + shift[:, :, None, :]
    return F.linear(y, sd[prefix + "linear_out.weight"], sd[prefix + "linear_out.bias"])


# --------------------------------------------------------------------------------------------
# ST block (st_transformer.py:79-114) and stack (:172-177)
# --------------------------------------------------------------------------------------------
def _maybe_ln(x: Tensor, sd: SD, key: str) -> Tensor:
    """st_transformer.py:50,75: LayerNorm(eps 1e-5) unless qk_norm (then Identity, no params)."""
    if key + ".weight" in sd:
        return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], 1e-5)
    return x


def st_block(x_BTSC: Tensor, a_BTC: Optional[Tensor], sd: SD, i: int, dom: Optional[str], cfg: OracleConfig) -> Tensor:
    B, T, S, C = x_BTSC.shape
    p = f"decoder.layers.{i}."
    # spatial attention over the S tokens of each frame, pre-norm (st_transformer.py:85-86)
    x = x_BTSC.reshape(B * T, S, C)
    x = x + self_attention(_maybe_ln(x, sd, p + "norm1"), sd, p + "spatial_attn.", cfg, causal=False)
    x = x.reshape(B, T, S, C)
    # action conditioning (st_transformer.py:91-106)
    if a_BTC is not None and dom is not None:
        if "mlp" in cfg.action_network:  # additive, :93-97 (projector is Identity)
            x = x + a_BTC[:, :T, None, :]
        elif "modulate" in cfg.action_network:  # :102-104
            x = x + modulate_layer(x, a_BTC, sd, p + f"action_projectors.{dom}.")
    # causal temporal attention per spatial slot, NO pre-norm (st_transformer.py:111)
    xt = x.permute(0, 2, 1, 3).reshape(B * S, T, C)
    xt = xt + self_attention(xt, sd, p + "temporal_attn.", cfg, causal=True)
    # MLP, pre-norm, erf GELU (st_transformer.py:24-27,112)
    h = _maybe_ln(xt, sd, p + "norm2")
    h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd.get(p + "mlp.fc1.bias")))
    xt = xt + F.linear(h, sd[p + "mlp.fc2.weight"], sd.get(p + "mlp.fc2.bias"))
    return xt.reshape(B, S, T, C).permute(0, 2, 1, 3)


# --------------------------------------------------------------------------------------------
# compute_logits (st_mask_git.py:632-686)
# --------------------------------------------------------------------------------------------
def hidden_states(x_THW: Tensor, action_ids: Optional[Tensor], domain: Optional[Sequence[str]], sd: SD,
                  cfg: OracleConfig, skip_normalization: bool = False, layers: Optional[int] = None) -> Tensor:
    B, T = x_THW.shape[:2]
    x = token_embed(x_THW.reshape(B, T, -1), sd, cfg)  # :640-641
    a = None
    dom = domain[0] if domain is not None else None  # :648,669
    if action_ids is not None:
        a = action_stem(action_ids, sd, dom, skip_normalization)  # :645-649
        if "concat" in cfg.action_network:  # :651-661
            x = torch.cat([x, a[:, :T, None].expand(-1, -1, cfg.action_token_size, -1)], dim=2)
    x = x + sd["pos_embed_TSC"][:, :T, : x.shape[2]]  # :670-672
    for i in range(cfg.num_layers if layers is None else layers):
        if CHECKPOINT_LAYERS and torch.is_grad_enabled():
            from torch.utils.checkpoint import checkpoint
            x = checkpoint(lambda xx, aa, i=i: st_block(xx, aa, sd, i, dom, cfg), x, a, use_reentrant=False)
        else:
            x = st_block(x, a, sd, i, dom, cfg)
    return x


def readout(x: Tensor, sd: SD, cfg: OracleConfig) -> Tensor:
    """st_mask_git.py:191-192,772-789: nn.Linear, or FixedMuReadout = Linear(output_mult*x/width_mult)
    with output_mult=1 and width_mult = d_model/256 (base shape hard-coded at :755-760)."""
    if cfg.use_mup:
        x = x / (cfg.d_model / 256.0)
    return F.linear(x, sd["out_x_proj.weight"], sd["out_x_proj.bias"])


def compute_logits(x_THW: Tensor, action_ids: Optional[Tensor], domain, sd: SD, cfg: OracleConfig,
                   skip_normalization: bool = False) -> Tensor:
    B, T, H, W = x_THW.shape
    x = hidden_states(x_THW, action_ids, domain, sd, cfg, skip_normalization)
    x = x[:, :, : H * W]  # drop the action tokens (:681)
    logits = readout(x, sd, cfg)  # [B,T,S,nv*vs]
    return logits.reshape(B, T, H, W, -1).permute(0, 4, 1, 2, 3)  # "B T (H W) C -> B C T H W" (:683)


# --------------------------------------------------------------------------------------------
# loss (st_mask_git.py:603-630; factorization_utils.py:85-96)
# --------------------------------------------------------------------------------------------
def video_loss_and_acc(logits_CTHW: Tensor, labels_flat: Tensor, relevant_mask_THW: Tensor, cfg: OracleConfig):
    B, C, T, H, W = logits_CTHW.shape
    nv, vs = cfg.num_factored_vocabs, cfg.factored_vocab_size
    targets = labels_flat.reshape(B, T, H, W)[:, 1:]
    lg = logits_CTHW[:, :, 1:].reshape(B, nv, vs, T - 1, H, W).permute(0, 2, 1, 3, 4, 5)  # b v nv t h w
    ft = factorize_token_ids(targets, nv, vs).permute(0, 4, 1, 2, 3)  # b nv t h w
    loss = F.cross_entropy(lg, ft, reduction="none", label_smoothing=0.01).sum(dim=1)
    acc = (lg.argmax(dim=1) == ft).all(dim=1)
    n = relevant_mask_THW.sum()
    return (loss * relevant_mask_THW).sum() / n, (acc * relevant_mask_THW).sum().float() / n


def forward(input_ids: Tensor, labels: Tensor, action_ids: Optional[Tensor], domain, sd: SD, cfg: OracleConfig,
            h: Optional[int] = None, w: Optional[int] = None):
    """st_mask_git.py:688-735 for jointly_predict_actions=False (the action-mask RNG at :703-710 is then
    dead code: relevant_action_mask is only read under jointly_predict_actions, :655)."""
    B = input_ids.shape[0]
    hh = h or math.isqrt(cfg.S)
    ww = w or math.isqrt(cfg.S)
    x_THW = input_ids.reshape(B, cfg.T, hh, ww)
    logits = compute_logits(x_THW, action_ids, domain, sd, cfg)
    relevant = x_THW[:, 1:] == cfg.mask_token_id
    loss, acc = video_loss_and_acc(logits, labels, relevant, cfg)
    return loss, acc, logits


# --------------------------------------------------------------------------------------------
# MaskGIT decode (st_mask_git.py:337-467) and the AR driver (:253-329)
# --------------------------------------------------------------------------------------------
def cosine_schedule(u: float) -> float:  # :116-125
    return math.cos(u * math.pi / 2)


@torch.no_grad()
def maskgit_generate(prompt_THW: Tensor, out_t: int, sd: SD, cfg: OracleConfig, maskgit_steps: int = 1,
                     temperature: float = 0.0, unmask_mode: str = "random", action_ids=None, domain=None,
                     generator: Optional[torch.Generator] = None, noise: Optional[dict] = None):
    """Restates :337-467. RNG is drawn in the reference's order: per step, per half
    (high first) an Exp(1) tensor [B*H*W, vs] when sampling, then a U(0,1) tensor [B,H,W] when
    re-masking in "random" mode. `noise` may inject those tensors ({"exp": [[hi, lo], ...], "rand": [...]})
    so a device implementation can be compared bit-exactly on identical noise."""
    assert out_t > 0
    assert torch.all(prompt_THW[:, out_t:] == cfg.mask_token_id)
    B, T, H, W = prompt_THW.shape
    S = H * W
    nv, vs = cfg.num_factored_vocabs, cfg.factored_vocab_size
    dev = prompt_THW.device
    unmasked = torch.zeros(B, S, dtype=torch.bool, device=dev)
    orig = None
    samples = None
    for step in range(maskgit_steps):
        logits = compute_logits(prompt_THW, action_ids, domain, sd, cfg)[:, :, out_t]  # [B, nv*vs, H, W]
        if orig is None:
            orig = logits.clone()
        fl = logits.reshape(B, nv, vs, H, W).permute(0, 2, 1, 3, 4)  # b vs nv h w
        probs = fl.softmax(dim=1)
        samples = torch.zeros(B, H, W, dtype=torch.long, device=dev)
        conf = torch.ones(B, H, W, device=dev)
        for j, k in enumerate(reversed(range(nv))):  # .flip(2).unbind(2): high factor first (:408)
            p = probs[:, :, k]  # b vs h w
            if temperature <= 1e-8:
                s = p.argmax(dim=1)
            else:
                # Categorical(probs=p/temperature).sample() == multinomial(renormalised p) ==
                # argmax(p / Exp(1)); the temperature cancels in the renormalisation (:411-416)
                p2 = p.permute(0, 2, 3, 1).reshape(-1, vs)
                p2 = p2 / temperature
                p2 = p2 / p2.sum(-1, keepdim=True)
                if noise is not None:
                    q = noise["exp"][step][j]
                else:
                    q = torch.empty_like(p2).exponential_(1, generator=generator)
                s = (p2 / q).argmax(dim=-1).reshape(B, H, W)
            samples = samples * vs + s
            conf = conf * torch.gather(p, 1, s.unsqueeze(1)).squeeze(1)
        prev_unmasked = unmasked.clone()
        prev_flat = prompt_THW[:, out_t].reshape(B, S)
        flat = samples.reshape(B, S)
        if step != maskgit_steps - 1:
            n = math.ceil(cosine_schedule((step + 1) / maskgit_steps) * S)  # :428
            if unmask_mode == "greedy":
                keys = conf.reshape(B, S).clone()
            elif unmask_mode == "random":
                if noise is not None:
                    keys = noise["rand"][step].reshape(B, S).clone()
                else:
                    keys = torch.rand(B, H, W, generator=generator, device=dev).reshape(B, S)
            else:
                raise NotImplementedError(unmask_mode)
            keys[unmasked] = torch.inf
            order = torch.argsort(keys, dim=1, stable=True)
            unmasked.scatter_(1, order[:, n:], True)
            flat.scatter_(1, order[:, :n], cfg.mask_token_id)
        flat[prev_unmasked] = prev_flat[prev_unmasked]
        samples = flat.reshape(B, H, W)
        prompt_THW[:, out_t] = samples  # in place on the caller's tensor (:453)
    factored_logits = orig.reshape(B, nv, vs, H, W).permute(0, 2, 1, 3, 4)
    return samples, factored_logits


@torch.no_grad()
def generate(input_ids: Tensor, max_new_tokens: int, sd: SD, cfg: OracleConfig, h: int, w: int, maskgit_steps: int = 1,
             temperature: float = 0.0, action_ids=None, domain=None, generator=None, unmask_mode: str = "random"):
    """st_mask_git.py:253-329."""
    B = input_ids.shape[0]
    S = h * w
    n_new = max_new_tokens // S
    prompt = input_ids.clone().reshape(B, -1, h, w)
    full = torch.cat([prompt, torch.full((B, n_new, h, w), cfg.mask_token_id, dtype=torch.long, device=prompt.device)], dim=1)
    all_logits = []
    for t in range(prompt.shape[1], prompt.shape[1] + n_new):
        s, fl = maskgit_generate(full, t, sd, cfg, maskgit_steps, temperature, unmask_mode, action_ids, domain,
                                 generator)
        full[:, t] = s
        all_logits.append(fl)
    return full.reshape(B, -1), torch.stack(all_logits, dim=3)


# --------------------------------------------------------------------------------------------
# deterministic weights with the reference's key layout
# --------------------------------------------------------------------------------------------
def make_state_dict(cfg: OracleConfig, domains: Sequence[str], d_actions: Sequence[int], seed: int = 0,
                    std: float = 0.05, action_dims: Optional[Sequence[int]] = None) -> SD:
    """Non-degenerate random weights (the reference init is near-degenerate): every >=2-D
    tensor ~ N(0, std^2), norm weights ~ 1 + N(0, 0.1^2), biases ~ N(0, 0.02^2). Key order is fixed so
    the same seed gives the same tensors everywhere."""
    g = torch.Generator().manual_seed(seed)
    d, L, nv, vs = cfg.d_model, cfg.num_layers, cfg.num_factored_vocabs, cfg.factored_vocab_size
    hd = d // cfg.num_heads
    hidden = int(d * cfg.mlp_ratio)
    sd: SD = {}

    def mat(*shape, s=std):
        return torch.randn(*shape, generator=g) * s

    def vec(n, s=0.02):
        return torch.randn(n, generator=g) * s

    def gain(n):
        return 1.0 + torch.randn(n, generator=g) * 0.1

    sd["pos_embed_TSC"] = mat(1, cfg.T, cfg.S + cfg.action_token_size, d)
    sd["action_mask_tokens"] = mat(1, cfg.T, 1, d)
    sd["token_embed.mask_token_embed"] = mat(1, d)
    for k in range(nv):
        sd[f"token_embed.factored_embeds.{k}.weight"] = mat(vs, d, s=0.5)
    for i in range(L):
        p = f"decoder.layers.{i}."
        if not cfg.qk_norm:
            sd[p + "norm1.weight"], sd[p + "norm1.bias"] = gain(d), vec(d)
        for att in ("spatial_attn.", "temporal_attn."):
            sd[p + att + "qkv.weight"] = mat(3 * d, d)
            if cfg.qkv_bias:
                sd[p + att + "qkv.bias"] = vec(3 * d)
            sd[p + att + "proj.weight"] = mat(d, d)
            if cfg.proj_bias:
                sd[p + att + "proj.bias"] = vec(d)
            if cfg.qk_norm:
                sd[p + att + "norm.weight"], sd[p + att + "norm.bias"] = gain(hd), vec(hd)
        if not cfg.qk_norm:
            sd[p + "norm2.weight"], sd[p + "norm2.bias"] = gain(d), vec(d)
        sd[p + "mlp.fc1.weight"] = mat(hidden, d)
        sd[p + "mlp.fc2.weight"] = mat(d, hidden, s=std / 2)
        if cfg.mlp_bias:
            sd[p + "mlp.fc1.bias"], sd[p + "mlp.fc2.bias"] = vec(hidden), vec(d)
        if "modulate" in cfg.action_network:
            for dom in domains:
                q = p + f"action_projectors.{dom}."
                sd[q + "linear_out.weight"], sd[q + "linear_out.bias"] = mat(d, d), vec(d)
                sd[q + "adaLN_modulation.0.weight"], sd[q + "adaLN_modulation.0.bias"] = mat(d, d), vec(d)
                sd[q + "adaLN_modulation.2.weight"], sd[q + "adaLN_modulation.2.bias"] = mat(2 * d, d), vec(2 * d)
    sd["out_x_proj.weight"], sd["out_x_proj.bias"] = mat(nv * vs, d, s=0.2), vec(nv * vs)
    for j, (dom, da) in enumerate(zip(domains, d_actions)):
        adim = da if action_dims is None else action_dims[j]
        sd[f"action_preprocessor.{dom}.mean"] = torch.randn(adim, generator=g) * 0.3
        sd[f"action_preprocessor.{dom}.std"] = 0.5 + torch.rand(adim, generator=g)
        q = f"action_mlp.{dom}.model."
        sd[q + "0.weight"], sd[q + "0.bias"] = mat(d, da, s=0.3), vec(d)
        sd[q + "1.weight"], sd[q + "1.bias"] = gain(d), vec(d)
        sd[q + "3.weight"], sd[q + "3.bias"] = mat(d, d, s=0.08), vec(d)
        sd[f"action_out_projectors.{dom}.weight"], sd[f"action_out_projectors.{dom}.bias"] = mat(da, d), vec(da)
    return sd
