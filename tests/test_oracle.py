"""Pins the CPU oracle (oracle/stmaskgit_oracle.py) against outputs of the real reference:
the committed fixtures (tests/golden/*.pt, made by oracle/make_golden.py) and, when
/root/reference is present, the live reference. CPU only."""
import math
from pathlib import Path

import pytest
import torch

from oracle import reference_loader
from oracle import stmaskgit_oracle as O

GOLDEN = Path(__file__).parent / "golden"
VARIANTS = {
    "tiny_magvit": dict(num_layers=2, num_heads=8, d_model=256, T=4, S=256, use_mup=False, qk_norm=False,
                        qkv_bias=False, action_network="concat+modulate"),
    "tiny_mup_qknorm": dict(num_layers=2, num_heads=8, d_model=256, T=4, S=256, use_mup=True, qk_norm=True,
                            qkv_bias=True, action_network="concat+modulate"),
}


def _setup(name):
    rec = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    cfg = O.OracleConfig(num_factored_vocabs=2, **VARIANTS[name])
    sd = O.make_state_dict(cfg, rec["domains"], rec["d_actions"], seed=rec["seed"], action_dims=rec["action_dims"])
    return rec, cfg, sd


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_forward_backward_matches_reference_fixture(name):
    rec, cfg, sd = _setup(name)
    torch.set_num_threads(8)
    for dom in rec["domains"]:
        r = rec[dom]
        params = {k: v.clone().requires_grad_(v.is_floating_point() and "action_preprocessor" not in k)
                  for k, v in sd.items()}
        loss, acc, logits = O.forward(r["input_ids"], r["labels"], r["actions"], [dom, dom], params, cfg)
        assert torch.allclose(loss, r["loss"], rtol=1e-5, atol=1e-6), (loss.item(), r["loss"].item())
        assert torch.equal(acc, r["acc"])
        torch.testing.assert_close(logits[:, :, :, ::4, ::4], r["logits_sub"], rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(logits[0, :, -1], r["logits_b0_last"], rtol=1e-4, atol=1e-4)
        loss.backward()
        for k, gn in r["grad_norms"].items():
            g = params[k].grad
            assert g is not None, k
            assert math.isclose(g.norm().item(), gn, rel_tol=2e-4, abs_tol=1e-7), (k, g.norm().item(), gn)
            sl = g.reshape(-1)[:: max(1, g.numel() // 64)][:64]
            torch.testing.assert_close(sl, r["grad_slices"][k], rtol=2e-3, atol=1e-6)
        # every parameter the reference left without a gradient must be untouched here too
        for k, p in params.items():
            if p.requires_grad and k not in r["grad_norms"]:
                assert p.grad is None or p.grad.abs().max() == 0, k


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_maskgit_matches_reference_fixture(name):
    rec, cfg, sd = _setup(name)
    h = w = math.isqrt(cfg.S)
    B = 2
    dom = rec["domains"][0]
    r = rec[dom]
    for tag, steps, temp, mode in (("greedy1", 1, 0.0, "random"), ("greedy3", 3, 0.0, "greedy"),
                                   ("sample2", 2, 1.0, "random")):
        prompt = r["labels"].reshape(B, cfg.T, h, w).clone()
        prompt[:, -1] = cfg.mask_token_id
        torch.manual_seed(777)
        s, fl = O.maskgit_generate(prompt, cfg.T - 1, sd, cfg, steps, temp, mode, r["actions"], [dom, dom])
        assert torch.equal(s, r[f"gen_{tag}_samples"]), tag
        assert torch.equal(prompt[:, -1], r[f"gen_{tag}_prompt_after"]), tag
        torch.testing.assert_close(fl[:, :, :, ::4, ::4], r[f"gen_{tag}_logits_sub"], rtol=1e-4, atol=1e-4)


def test_oracle_generate_matches_reference_fixture():
    rec, cfg, sd = _setup("tiny_magvit")
    h = w = math.isqrt(cfg.S)
    dom = rec["domains"][0]
    r = rec[dom]
    torch.manual_seed(4242)
    toks, logits = O.generate(r["labels"][:, : 2 * cfg.S], 2 * cfg.S, sd, cfg, h, w, maskgit_steps=2, temperature=0.0,
                              action_ids=r["actions"], domain=[dom, dom])
    assert torch.equal(toks, rec["generate_tokens"])
    assert logits.shape == (2, 512, 2, 2, h, w)


def test_maskgit_schedule_properties():
    """masked counts for K=4 are 237, 182, 98; monotone; nothing left masked at the end."""
    S = 256
    counts = [math.ceil(O.cosine_schedule((k + 1) / 4) * S) for k in range(3)]
    assert counts == [237, 182, 98]
    rec, cfg, sd = _setup("tiny_magvit")
    dom = rec["domains"][0]
    r = rec[dom]
    prompt = r["labels"].reshape(2, cfg.T, 16, 16).clone()
    prompt[:, -1] = cfg.mask_token_id
    torch.manual_seed(0)
    s, _ = O.maskgit_generate(prompt, cfg.T - 1, sd, cfg, 4, 1.0, "random", r["actions"], [dom, dom])
    assert (s != cfg.mask_token_id).all()


@pytest.mark.skipif(not reference_loader.available(), reason="/root/reference not present")
def test_oracle_matches_live_reference_causality_and_layout():
    """Live reference: state_dict loads strictly; frames <= t do not depend on frames > t."""
    from oracle.make_golden import build_reference, synthetic_batch, DOMAINS

    model, cfg, sd = build_reference("tiny_magvit")
    x, labels, actions = synthetic_batch(cfg, seed=5, domain_idx=1)
    x_THW = x.reshape(2, cfg.T, 16, 16)
    with torch.no_grad():
        ref, _ = model.compute_logits(x_THW, action_ids=actions, domain=[DOMAINS[1]] * 2)
        mine = O.compute_logits(x_THW, actions, [DOMAINS[1]] * 2, sd, cfg)
        torch.testing.assert_close(mine, ref, rtol=1e-4, atol=1e-4)
        x2 = x_THW.clone()
        x2[:, -1] = torch.randint(0, 262144, x2[:, -1].shape)
        mine2 = O.compute_logits(x2, actions, [DOMAINS[1]] * 2, sd, cfg)
    assert torch.equal(mine2[:, :, :-1], mine[:, :, :-1])
