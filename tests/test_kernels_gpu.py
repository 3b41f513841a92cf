"""Each CUDA stage against a plain fp32 torch restatement of the reference op (GPU only)."""
import math

import pytest
import torch
import torch.nn.functional as F

from hma_b200 import _lib

pytestmark = pytest.mark.gpu
S_ = _lib.current_stream


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("frames,n", [(3, 320), (5, 256), (2, 128), (4, 64), (2, 208)])
def test_attn_spatial_fwd(frames, n):
    torch.manual_seed(0)
    H, hd = 8, 32
    C = H * hd
    qkv = (torch.randn(frames * n, 3 * C, device="cuda") * 1.0).bfloat16()
    out = torch.zeros(frames * n, C, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(frames, H, n, device="cuda")
    scale = hd ** -0.5
    _lib.call("hma_attn_spatial_fwd", qkv.data_ptr(), 3 * C, frames, n, H, 0, C, 2 * C, scale, out.data_ptr(), C,
              lse.data_ptr(), S_())
    torch.cuda.synchronize()
    q, k, v = qkv.float().reshape(frames, n, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    ref = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(frames * n, C)
    assert relerr(out, ref) < 2e-2
    ref_lse = torch.logsumexp(s, dim=-1) / math.log(2.0)
    assert (lse - ref_lse).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,T,n", [(2, 4, 320), (1, 16, 40), (2, 32, 24), (1, 5, 7), (2, 12, 320), (1, 128, 3), (1, 1, 16)])
def test_attn_temporal_fwd_bwd(B, T, n):
    torch.manual_seed(1)
    H, hd = 8, 32
    C = H * hd
    rows = B * T * n
    qkv = torch.randn(rows, 3 * C, device="cuda").bfloat16()
    dout = torch.randn(rows, C, device="cuda").bfloat16()
    out = torch.zeros(rows, C, device="cuda", dtype=torch.bfloat16)
    dqkv = torch.zeros(rows, 3 * C, device="cuda", dtype=torch.bfloat16)
    scale = 0.25
    lse = torch.zeros(rows, H, device="cuda")
    _lib.call("hma_attn_temporal_fwd", qkv.data_ptr(), 3 * C, B, T, n, H, 0, C, 2 * C, scale, out.data_ptr(), C,
              lse.data_ptr(), S_())
    _lib.call("hma_attn_temporal_bwd", qkv.data_ptr(), 3 * C, out.data_ptr(), C, dout.data_ptr(), C, lse.data_ptr(), B,
              T, n, H, 0, C, 2 * C, scale, dqkv.data_ptr(), 3 * C, S_())
    torch.cuda.synchronize()
    x = qkv.float().reshape(B, T, n, 3, H, hd).requires_grad_(True)
    q, k, v = x.permute(3, 0, 2, 4, 1, 5)  # [B, n, H, T, hd]
    s = (q @ k.transpose(-1, -2)) * scale
    mask = torch.ones(T, T, dtype=torch.bool, device="cuda").tril()
    s = s.masked_fill(~mask, float("-inf"))
    o = (s.softmax(-1) @ v).permute(0, 3, 1, 2, 4).reshape(rows, C)  # [B,T,n,H,hd]
    assert relerr(out, o) < 2e-2
    o.backward(dout.float())
    assert relerr(dqkv, x.grad.reshape(rows, 3 * C)) < 2e-2


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_ln_fwd_bwd(mode):
    torch.manual_seed(2)
    groups, rpg, C = 6, 320, 256
    rows = groups * rpg
    x = (torch.randn(rows, C, device="cuda") * 2 + 0.3)
    gamma = 1 + 0.1 * torch.randn(C, device="cuda")
    beta = 0.1 * torch.randn(C, device="cuda")
    mod = 0.3 * torch.randn(groups, 2 * C, device="cuda")
    y = torch.zeros(rows, C, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(rows, 2, device="cuda")
    eps = 1e-5 if mode == 1 else 1e-6
    _lib.call("hma_ln_fwd", x.data_ptr(), C, rows, mode, gamma.data_ptr(), beta.data_ptr(), mod.data_ptr(), rpg, eps,
              y.data_ptr(), C, stats.data_ptr(), 0, 0, S_())
    torch.cuda.synchronize()
    xr = x.clone().requires_grad_(True)
    gr, br, mr = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True), mod.clone().requires_grad_(True)
    if mode == 0:
        ref = xr
    elif mode == 1:
        ref = F.layer_norm(xr, (C,), gr, br, eps)
    else:
        shift, scale = mr.chunk(2, dim=-1)
        ref = F.layer_norm(xr, (C,), None, None, eps).reshape(groups, rpg, C) * (1 + scale[:, None]) + shift[:, None]
        ref = ref.reshape(rows, C)
    assert relerr(y, ref) < 1e-2
    if mode == 0:
        return
    dy = torch.randn(rows, C, device="cuda").bfloat16()
    dx = torch.randn(rows, C, device="cuda")
    dx0 = dx.clone()
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dmod = torch.zeros(groups, 2 * C, device="cuda")
    dyn = torch.zeros(rows, C, device="cuda", dtype=torch.bfloat16)
    csum = torch.zeros(C, device="cuda")
    _lib.call("hma_ln_bwd", dy.data_ptr(), C, x.data_ptr(), C, stats.data_ptr(), rows, mode, gamma.data_ptr(),
              mod.data_ptr(), rpg, dx.data_ptr(), C, dgamma.data_ptr(), dbeta.data_ptr(), dmod.data_ptr(),
              dyn.data_ptr(), csum.data_ptr(), S_())
    torch.cuda.synchronize()
    assert torch.equal(dyn, dx.bfloat16())
    assert relerr(csum, dyn.float().sum(0)) < 1e-3
    ref.backward(dy.float())
    assert relerr(dx - dx0, xr.grad) < 2e-3
    if mode == 1:
        assert relerr(dgamma, gr.grad) < 2e-3 and relerr(dbeta, br.grad) < 2e-3
    else:
        assert relerr(dmod, mr.grad) < 2e-3


def test_colsum_cast_transpose():
    torch.manual_seed(3)
    G = torch.randn(1000, 768, device="cuda").bfloat16()
    out = torch.ones(768, device="cuda")
    _lib.call("hma_colsum_bf16", G.data_ptr(), 768, 1000, 768, out.data_ptr(), S_())
    W = torch.randn(300, 70, device="cuda")
    Wb = torch.zeros(300, 70, device="cuda", dtype=torch.bfloat16)
    Wt = torch.zeros(70, 300, device="cuda", dtype=torch.bfloat16)
    _lib.call("hma_cast_transpose", W.data_ptr(), 300, 70, Wb.data_ptr(), Wt.data_ptr(), 0.5, S_())
    torch.cuda.synchronize()
    assert relerr(out, 1 + G.float().sum(0)) < 1e-3
    assert torch.equal(Wb, (W * 0.5).bfloat16()) and torch.equal(Wt, (W * 0.5).bfloat16().t())


# ------------------------------------------------------------------------------------------------
def _embed_ref(ids, E0, E1, me, act, pos, B, T, S, A, vs, mask_id):
    is_mask = ids == mask_id
    safe = torch.where(is_mask, torch.zeros_like(ids), ids)
    e = E0[safe % vs] + E1[(safe // vs) % vs]
    e = torch.where(is_mask[..., None], me.expand_as(e), e).reshape(B, T, S, -1)
    if A:
        e = torch.cat([e, act.reshape(B, T, 1, -1).expand(-1, -1, A, -1)], dim=2)
    return e + pos[None, :T, : S + A]


@pytest.mark.parametrize("A", [64, 0])
def test_embed_fwd_bwd(A):
    torch.manual_seed(4)
    B, T, S, C, vs = 3, 4, 256, 256, 512
    mask_id = vs * vs
    ids = torch.randint(0, mask_id, (B * T * S,), device="cuda")
    ids[torch.rand(B * T * S, device="cuda") < 0.4] = mask_id
    E0, E1 = torch.randn(vs, C, device="cuda"), torch.randn(vs, C, device="cuda")
    me, act = torch.randn(1, C, device="cuda"), torch.randn(B * T, C, device="cuda")
    pos = torch.randn(T + 2, S + 64, C, device="cuda")
    n = S + A
    x = torch.zeros(B * T * n, C, device="cuda")
    _lib.call("hma_embed_fwd", ids.data_ptr(), E0.data_ptr(), E1.data_ptr(), me.data_ptr(), act.data_ptr(),
              pos.data_ptr(), S + 64, B, T, S, A, vs, mask_id, x.data_ptr(), S_())
    torch.cuda.synchronize()
    leaves = [t.clone().requires_grad_(True) for t in (E0, E1, me, act, pos)]
    ref = _embed_ref(ids, *leaves, B, T, S, A, vs, mask_id)
    assert torch.allclose(x.reshape(B, T, n, C), ref, atol=1e-6)
    dx = torch.randn(B * T * n, C, device="cuda")
    grads = [torch.zeros_like(t) for t in (E0, E1, me, act, pos)]
    _lib.call("hma_embed_bwd", ids.data_ptr(), dx.data_ptr(), S + 64, B, T, S, A, vs, mask_id, grads[0].data_ptr(),
              grads[1].data_ptr(), grads[2].data_ptr(), grads[3].data_ptr() if A else None, grads[4].data_ptr(), S_())
    torch.cuda.synchronize()
    ref.backward(dx.reshape(B, T, n, C))
    for g, l in zip(grads, leaves):
        if l.grad is None:
            assert g.abs().max() == 0
        else:
            assert relerr(g, l.grad) < 1e-4


def test_ce_fwd_bwd():
    torch.manual_seed(5)
    B, T, S, nv, vs = 2, 4, 256, 2, 512
    mask_id = vs ** nv
    rows = B * T * S
    logits = torch.randn(rows, nv * vs, device="cuda") * 3
    labels = torch.randint(0, mask_id, (rows,), device="cuda")
    ids = labels.clone()
    ids[torch.rand(rows, device="cuda") < 0.6] = mask_id
    # make some rows exactly right so that acc > 0
    for r in range(0, rows, 3):
        logits[r, labels[r] % vs] += 20
        logits[r, vs + (labels[r] // vs) % vs] += 20
    lse = torch.zeros(rows, nv, device="cuda")
    sums, la = torch.zeros(3, device="cuda"), torch.zeros(2, device="cuda")
    _lib.call("hma_ce_fwd", logits.data_ptr(), nv * vs, labels.data_ptr(), ids.data_ptr(), B, T, S, nv, vs, mask_id,
              0.01, lse.data_ptr(), sums.data_ptr(), la.data_ptr(), S_())
    dloss = torch.tensor([0.7], device="cuda")
    dlogits = torch.zeros(rows, nv * vs, device="cuda", dtype=torch.bfloat16)
    _lib.call("hma_ce_bwd", logits.data_ptr(), nv * vs, labels.data_ptr(), ids.data_ptr(), B, T, S, nv, vs, mask_id,
              0.01, lse.data_ptr(), sums.data_ptr(), dloss.data_ptr(), dlogits.data_ptr(), nv * vs, S_())
    torch.cuda.synchronize()
    lg = logits.clone().requires_grad_(True)
    l3 = lg.reshape(B, T, S, nv, vs)[:, 1:]
    tg = torch.stack([(labels // vs ** k) % vs for k in range(nv)], -1).reshape(B, T, S, nv)[:, 1:]
    per = F.cross_entropy(l3.reshape(-1, vs), tg.reshape(-1), reduction="none", label_smoothing=0.01)
    per = per.reshape(B, T - 1, S, nv).sum(-1)
    acc = (l3.argmax(-1) == tg).all(-1)
    m = (ids.reshape(B, T, S)[:, 1:] == mask_id)
    loss = (per * m).sum() / m.sum()
    assert abs(la[0].item() - loss.item()) < 1e-4 * abs(loss.item())
    assert abs(la[1].item() - ((acc * m).sum().float() / m.sum()).item()) < 1e-6
    (loss * 0.7).backward()
    assert relerr(dlogits, lg.grad) < 1e-2


def test_sample_and_remask():
    torch.manual_seed(6)
    B, S, nv, vs = 3, 256, 2, 512
    mask_id = vs ** nv
    T = 4
    logits_all = torch.randn(B, T, S, nv * vs, device="cuda") * 2
    lg = logits_all[:, 2]  # strided view: frame 2
    samples = torch.zeros(B * S, dtype=torch.long, device="cuda")
    conf = torch.zeros(B * S, device="cuda")
    probs = lg.reshape(B * S, nv, vs).softmax(-1)
    # greedy
    _lib.call("hma_sample_tokens", lg.data_ptr(), lg.stride(0), lg.stride(1), B, S, nv, vs, None, 1.0, samples.data_ptr(),
              conf.data_ptr(), S_())
    torch.cuda.synchronize()
    hi, lo = probs[:, 1].argmax(-1), probs[:, 0].argmax(-1)
    assert torch.equal(samples, hi * vs + lo)
    ref_conf = probs[:, 1].gather(1, hi[:, None])[:, 0] * probs[:, 0].gather(1, lo[:, None])[:, 0]
    assert torch.allclose(conf, ref_conf, rtol=1e-4)
    # sampling with injected Exp(1) noise (hi first)
    q = torch.empty(nv, B * S, vs, device="cuda").exponential_(1)
    _lib.call("hma_sample_tokens", lg.data_ptr(), lg.stride(0), lg.stride(1), B, S, nv, vs, q.data_ptr(), 1.0,
              samples.data_ptr(), conf.data_ptr(), S_())
    torch.cuda.synchronize()
    hi, lo = (probs[:, 1] / q[0]).argmax(-1), (probs[:, 0] / q[1]).argmax(-1)
    assert (samples != hi * vs + lo).float().mean().item() < 2e-3  # ties at the last ulp only

    # rank + re-mask
    keys = torch.rand(B, S, device="cuda")
    unmasked = torch.rand(B, S, device="cuda") < 0.3
    prompt = torch.randint(0, mask_id, (B, T, S), device="cuda")
    prev = prompt[:, 2].clone()
    n = 100
    um = unmasked.to(torch.uint8).clone()
    new = torch.randint(0, mask_id, (B, S), device="cuda")
    outs = torch.zeros(B, S, dtype=torch.long, device="cuda")
    _lib.call("hma_rank_remask", keys.data_ptr(), um.data_ptr(), new.data_ptr(), prompt[:, 2].data_ptr(),
              prompt.stride(0), B, S, n, mask_id, outs.data_ptr(), S_())
    torch.cuda.synchronize()
    k2 = keys.clone()
    k2[unmasked] = float("inf")
    order = torch.argsort(k2, dim=1, stable=True)
    ref_um = unmasked.clone()
    ref_um.scatter_(1, order[:, n:], True)
    ref = new.clone()
    ref.scatter_(1, order[:, :n], mask_id)
    ref[unmasked] = prev[unmasked]
    assert torch.equal(outs, ref) and torch.equal(prompt[:, 2], ref) and torch.equal(um.bool(), ref_um)
    # last step: no ranking
    um2 = ref_um.to(torch.uint8).clone()
    before = prompt[:, 2].clone()
    _lib.call("hma_rank_remask", None, um2.data_ptr(), new.data_ptr(), prompt[:, 2].data_ptr(), prompt.stride(0), B, S,
              -1, mask_id, outs.data_ptr(), S_())
    torch.cuda.synchronize()
    ref2 = new.clone()
    ref2[ref_um] = before[ref_um]
    assert torch.equal(outs, ref2) and torch.equal(prompt[:, 2], ref2)


@pytest.mark.parametrize("frames,n", [(3, 320), (2, 256), (2, 128), (3, 64), (2, 208)])
def test_attn_spatial_bwd(frames, n):
    torch.manual_seed(7)
    H, hd = 8, 32
    C = H * hd
    rows = frames * n
    qkv = torch.randn(rows, 3 * C, device="cuda").bfloat16()
    dout = torch.randn(rows, C, device="cuda").bfloat16()
    out = torch.zeros(rows, C, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(frames, H, n, device="cuda")
    dqkv = torch.zeros(rows, 3 * C, device="cuda", dtype=torch.bfloat16)
    scale = 0.25
    _lib.call("hma_attn_spatial_fwd", qkv.data_ptr(), 3 * C, frames, n, H, 0, C, 2 * C, scale, out.data_ptr(), C,
              lse.data_ptr(), S_())
    _lib.call("hma_attn_spatial_bwd", qkv.data_ptr(), 3 * C, out.data_ptr(), C, dout.data_ptr(), C, lse.data_ptr(),
              frames, n, H, 0, C, 2 * C, scale, dqkv.data_ptr(), 3 * C, None, S_())
    torch.cuda.synchronize()
    # the same call with delta = rowsum(dO * O) per (token, head) handed in (what the engine does: the GEMM that produces
    # dO emits it) instead of computed in the kernel's prologue: bit-identical
    delta = (dout.float() * out.float()).view(rows, H, hd).sum(-1).contiguous()
    dqkv2 = torch.zeros_like(dqkv)
    _lib.call("hma_attn_spatial_bwd", qkv.data_ptr(), 3 * C, None, 0, dout.data_ptr(), C, lse.data_ptr(),
              frames, n, H, 0, C, 2 * C, scale, dqkv2.data_ptr(), 3 * C, delta.data_ptr(), S_())
    torch.cuda.synchronize()
    assert relerr(dqkv2, dqkv.float()) < 1e-3
    x = qkv.float().reshape(frames, n, 3, H, hd).requires_grad_(True)
    q, k, v = x.permute(2, 0, 3, 1, 4)
    o = (((q @ k.transpose(-1, -2)) * scale).softmax(-1) @ v).permute(0, 2, 1, 3).reshape(rows, C)
    o.backward(dout.float())
    g = x.grad.reshape(rows, 3 * C)
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        e = relerr(dqkv[:, sl], g[:, sl])
        assert e < 3e-2, (name, e)


# ------------------------------------------------------------------------------------------------
def test_qk_norm_fwd_bwd():
    """Per-head LayerNorm(32) of q and k with one shared affine (attention.py:32-35,47-52), v passed through."""
    from hma_b200 import ops
    torch.manual_seed(9)
    rows, H, hd = 1000, 8, 32
    C = H * hd
    qkv = (torch.randn(rows, 3 * C, device="cuda") * 1.7 + 0.4).bfloat16()
    gamma = 1 + 0.2 * torch.randn(hd, device="cuda")
    beta = 0.1 * torch.randn(hd, device="cuda")
    out = ops.qk_norm_fwd(qkv, gamma, beta)
    x = qkv.float().requires_grad_(True)
    q, k, v = x[:, :C], x[:, C:2 * C], x[:, 2 * C:]
    g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ln = lambda t: F.layer_norm(t.reshape(rows, H, hd), (hd,), g, b, 1e-5).reshape(rows, C)
    ref = torch.cat([ln(q), ln(k), v], dim=1)
    assert relerr(out, ref) < 1e-2
    assert torch.equal(out[:, 2 * C:], qkv[:, 2 * C:])
    dout = torch.randn(rows, 3 * C, device="cuda").bfloat16()
    ref.backward(dout.float())
    dqkv = dout.clone()
    dg = torch.zeros(hd, device="cuda")
    db = torch.zeros(hd, device="cuda")
    ops.qk_norm_bwd(qkv, gamma, dqkv, dg, db)
    torch.cuda.synchronize()
    assert relerr(dqkv, x.grad) < 1.5e-2
    assert relerr(dg, g.grad) < 5e-3 and relerr(db, b.grad) < 5e-3
