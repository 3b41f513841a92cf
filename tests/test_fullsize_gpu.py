"""Size-independent properties at BASELINE.json's FULL sizes (the oracle is too slow there):
config 2 = HMA-MagVit 32 layers, B=8, T=16, 16x16 tokens + 64 action tokens; config 3 = generate 8 -> 8 frames, B=64.

  causality               logits of frames <= t do not move (bit-exact) when tokens / actions of frames > t change
                          (st_transformer.py:111 causal temporal attention; spatial attention is per frame)
  batch independence      a sample's logits do not move (bit-exact) when the other samples change
  window truncation       forward on the first 8 frames == the first 8 frames of the 16-frame forward (<= 2e-3 * max|logit|:
                          a different tile packing of the temporal kernel changes the summation order)
  incremental == full     frame-incremental decode logits == full-window logits (<= 1e-2 * max|logit|)
  directional derivative  (L(w + e d) - L(w - e d)) / 2e == <grad, d> within 5 % for random directions on several tensors
  MaskGIT invariants      nothing left masked, prompt frames untouched, greedy decode deterministic and independent of the
                          other samples in the batch, re-masking follows the cosine schedule
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

T, S, A = 16, 256, 64
DOMAINS, D_ACTIONS, ADIMS = ["dA", "dB", "dC"], [7, 14, 70], [7, 14, 7]


@pytest.fixture(scope="module")
def model():
    from hma_b200 import GenieConfig, STMaskGIT
    cfg = GenieConfig(num_layers=32, num_heads=8, d_model=256, T=T, S=S, num_factored_vocabs=2, qk_norm=False, qkv_bias=False,
                      use_mup=False, action_network="concat+modulate")
    torch.manual_seed(0)
    with torch.device("cuda"):
        m = STMaskGIT(cfg)
        m.init_action_projectors(DOMAINS, D_ACTIONS, [[[0.0] * a, [1.0] * a] for a in ADIMS], "concat+modulate")
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.03)
    return m


def _batch(B, seed, dom=1, masked_from=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 262144, (B, T, 16, 16), generator=g)
    if masked_from is not None:
        x[:, masked_from:] = 262144
    a = torch.randn(B, T, D_ACTIONS[dom], generator=g)
    return x.cuda(), a.cuda(), [DOMAINS[dom]] * B


def test_causality_and_batch_independence_full_size(model):
    x, a, dom = _batch(8, 1)
    with torch.no_grad():
        ref, _ = model.compute_logits(x, action_ids=a, domain=dom)
        ref = ref.clone()
        t = 9
        x2, a2 = x.clone(), a.clone()
        x2[:, t + 1:] = torch.randint(0, 262144, x2[:, t + 1:].shape, device="cuda")
        x2[:, t + 3] = 262144
        a2[:, t + 1:] += 1.5
        got, _ = model.compute_logits(x2, action_ids=a2, domain=dom)
        assert torch.equal(got[:, :, : t + 1], ref[:, :, : t + 1]), "future frames leaked into the past"
        assert not torch.equal(got[:, :, t + 1:], ref[:, :, t + 1:])
        x3, a3 = x.clone(), a.clone()
        x3[1:] = torch.randint(0, 262144, x3[1:].shape, device="cuda")
        a3[1:] = a3[1:] * -0.7
        got, _ = model.compute_logits(x3, action_ids=a3, domain=dom)
        assert torch.equal(got[0], ref[0]), "a sample's logits depend on its batch mates"


def test_window_truncation_and_incremental_decode_full_size(model):
    x, a, dom = _batch(8, 2)
    with torch.no_grad():
        full, _ = model.compute_logits(x, action_ids=a, domain=dom)
        full = full.float().clone()
        scale = full.abs().max().item()
        short, _ = model.compute_logits(x[:, :8].contiguous(), action_ids=a[:, :8].contiguous(), domain=dom)
        d = (short.float() - full[:, :, :8]).abs().max().item()
        assert d <= 2e-3 * scale, (d, scale)
        out_t = 11
        prompt = x.clone()
        prompt[:, out_t:] = 262144
        want, _ = model.compute_logits(prompt, action_ids=a, domain=dom)
        want = want[:, :, out_t].permute(0, 2, 3, 1).reshape(8 * S, -1).float()
        model._sessions.clear()
        sess = model._decode_session(prompt, out_t, a, dom, {})
        got = sess.step(prompt[:, out_t], out_t).float()
        d = (got - want).abs().max().item()
        assert d <= 1e-2 * want.abs().max().item(), d
        model._sessions.clear()


def test_directional_derivative_full_size(model):
    """Whole-model gradient check on the 32-layer step: central difference of the loss along random directions."""
    g = torch.Generator().manual_seed(5)
    labels = torch.randint(0, 262144, (8, T * S), generator=g)
    mask = torch.rand(8, T, S, generator=g) < 0.5
    mask[:, 0] = False
    ids = torch.where(mask.view(8, -1), torch.full_like(labels, 262144), labels).cuda()
    labels = labels.cuda()
    acts = torch.randn(8, T, D_ACTIONS[2], generator=g).cuda()
    dom = [DOMAINS[2]] * 8
    model.zero_grad(set_to_none=True)
    out = model(ids, labels, action_ids=acts, domain=dom)
    out.loss.backward()
    named = dict(model.named_parameters())
    names = ["decoder.layers.5.mlp.fc1.weight", "decoder.layers.20.spatial_attn.qkv.weight", "decoder.layers.31.temporal_attn.proj.weight",
             "pos_embed_TSC", f"decoder.layers.12.action_projectors.{DOMAINS[2]}.adaLN_modulation.2.weight",
             "token_embed.factored_embeds.1.weight", "decoder.layers.0.norm1.weight"]
    L0 = out.loss.item()
    checked = 0
    for k in names:
        p = named[k]
        gk = p.grad.float()
        rnd = torch.randn(p.shape, generator=torch.Generator().manual_seed(sum(map(ord, k)) % 1000)).cuda()
        aligned = gk * (math.sqrt(gk.numel()) / gk.norm().clamp_min(1e-20))  # RMS 1, along the gradient: best signal / noise
        for label, d in (("random", rnd), ("aligned", aligned)):
            want = (gk * d).sum().item()
            # step: aim at |dL| ~ 2 % of L, but never move the tensor by more than 20 % of its own scale
            eps = min(0.2 * p.detach().abs().mean().item() / d.abs().mean().item(), 0.02 * L0 / max(abs(want), 1e-12))
            if abs(want) * eps < 3e-3 * L0:
                continue  # the loss would move by less than its bf16 evaluation noise: not checkable by differences
            vals = []
            with torch.no_grad():
                for sgn in (1.0, -1.0):
                    p.add_(d, alpha=sgn * eps)
                    vals.append(model(ids, labels, action_ids=acts, domain=dom).loss.item())
                    p.add_(d, alpha=-sgn * eps)
            got = (vals[0] - vals[1]) / (2 * eps)
            assert abs(got - want) <= 5e-2 * abs(want), (k, label, got, want, eps)
            checked += 1
    assert checked >= 6, checked


def test_maskgit_invariants_full_size(model):
    from hma_b200.model import cosine_schedule
    B, Tp, K = 64, 8, 4
    x, a, dom = _batch(B, 7, masked_from=Tp)
    prompt_frames = x[:, :Tp].clone()
    kw = dict(maskgit_steps=K, temperature=0.0, unmask_mode="greedy", action_ids=a, domain=dom)
    model._sessions.clear()
    p1 = x.clone()
    s1, fl, _ = model.maskgit_generate(p1, Tp, **kw)
    assert fl.shape == (B, 512, 2, 16, 16)
    assert (s1 != 262144).all() and torch.equal(p1[:, Tp], s1) and torch.equal(p1[:, :Tp], prompt_frames)
    assert (p1[:, Tp + 1:] == 262144).all()
    p2 = x.clone()
    s2, _, _ = model.maskgit_generate(p2, Tp, **kw)
    assert torch.equal(s1, s2), "greedy decode is not deterministic"
    # batch independence of the greedy decode: the first 8 samples alone give the same tokens
    p3 = x[:8].clone()
    s3, _, _ = model.maskgit_generate(p3, Tp, maskgit_steps=K, temperature=0.0, unmask_mode="greedy", action_ids=a[:8], domain=dom[:8])
    assert torch.equal(s3, s1[:8])
    # cosine re-masking schedule: after step k exactly ceil(cos(pi/2 (k+1)/K) S) tokens of the frame are masked
    from hma_b200 import ops
    seen = []
    orig = ops.rank_remask

    def spy(keys, unmasked, samples, frame, n_mask, mask_id):
        out = orig(keys, unmasked, samples, frame, n_mask, mask_id)
        seen.append((n_mask, (frame == mask_id).sum(1)))
        return out

    ops.rank_remask = spy
    try:
        model.maskgit_generate(x.clone(), Tp, **kw)
    finally:
        ops.rank_remask = orig
    assert len(seen) == K
    for k, (n_mask, counts) in enumerate(seen[:-1]):
        assert n_mask == math.ceil(cosine_schedule((k + 1) / K) * S)
        assert (counts == n_mask).all()
    assert (seen[-1][1] == 0).all()
    # generate(): 8 -> 8 frames, sampled; prompt preserved, everything unmasked
    torch.manual_seed(3)
    toks = model.generate(x[:, :Tp].reshape(B, -1), None, (T - Tp) * S, maskgit_steps=2, temperature=1.0, action_ids=a, domain=dom,
                          h=[16], w=[16])
    assert toks.shape == (B, T * S) and (toks != 262144).all()
    assert torch.equal(toks[:, : Tp * S], prompt_frames.reshape(B, -1))
    model._sessions.clear()
