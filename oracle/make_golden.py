"""Generate tests/golden/*.pt by running the REAL reference (/root/reference) in this container.

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY. The reference cannot travel to the GPU box, so its outputs on
deterministic weights/inputs are committed as small fixtures; tests/test_oracle.py pins the CPU
oracle (oracle/stmaskgit_oracle.py) against them, and the GPU tests pin the CUDA path against
both. Weights are NOT stored: oracle.stmaskgit_oracle.make_state_dict(seed) regenerates them, and
this script proves that state_dict loads strictly into the reference model (key-layout parity,
SURVEY.md Appendix A).
"""
from __future__ import annotations

import contextlib
import io
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import reference_loader  # noqa: E402
from oracle import stmaskgit_oracle as O  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

VARIANTS = {
    # the shipped MagVit architecture (hma/configs/magvit_n32_h8_d256_action.json) at 2 layers
    "tiny_magvit": dict(num_layers=2, num_heads=8, d_model=256, T=4, S=256, use_mup=False, qk_norm=False,
                        qkv_bias=False, action_network="concat+modulate"),
    # the flags train_multi.py actually forces / MAR-style attention: muP scale + readout, qk-norm, qkv bias
    "tiny_mup_qknorm": dict(num_layers=2, num_heads=8, d_model=256, T=4, S=256, use_mup=True, qk_norm=True,
                            qkv_bias=True, action_network="concat+modulate"),
}
DOMAINS = ["dom00", "dom01"]
D_ACTIONS = [14, 10]      # action_dim * stride (data.py:207-210)
ACTION_DIMS = [7, 10]     # ActionStat buffers have action_dim entries (st_mask_git.py:131-136)
B = 2


def synthetic_batch(cfg: O.OracleConfig, seed: int, domain_idx: int):
    """Collator-like masking (data.py:42-83): per (sample, frame>=1) mask rate cos(pi/2 * U)."""
    g = torch.Generator().manual_seed(seed)
    T, S = cfg.T, cfg.S
    labels = torch.randint(0, cfg.image_vocab_size, (B, T * S), generator=g)
    x = labels.clone().reshape(B, T, S)
    for b in range(B):
        for t in range(1, T):
            rate = math.cos(math.pi / 2 * torch.rand((), generator=g).item())
            m = torch.rand(S, generator=g) < rate
            x[b, t][m] = cfg.image_vocab_size
    actions = torch.randn(B, T, D_ACTIONS[domain_idx], generator=g)
    return x.reshape(B, T * S), labels, actions


def build_reference(name: str):
    STMaskGIT, GenieConfig = reference_loader.load()
    kw = dict(VARIANTS[name])
    rcfg = GenieConfig(num_factored_vocabs=2, **kw)
    with contextlib.redirect_stdout(io.StringIO()):
        model = STMaskGIT(rcfg)
        stats = [[[0.0] * a, [1.0] * a] for a in ACTION_DIMS]
        model.init_action_projectors(DOMAINS, D_ACTIONS, stats, rcfg.action_network)
    ocfg = O.OracleConfig(num_factored_vocabs=2, **kw)
    sd = O.make_state_dict(ocfg, DOMAINS, D_ACTIONS, seed=0, action_dims=ACTION_DIMS)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return model.eval(), ocfg, sd


def run_variant(name: str) -> dict:
    model, cfg, sd = build_reference(name)
    out = {"variant": name, "domains": DOMAINS, "d_actions": D_ACTIONS, "action_dims": ACTION_DIMS, "seed": 0}
    h = w = math.isqrt(cfg.S)
    for di, dom in enumerate(DOMAINS):
        x, labels, actions = synthetic_batch(cfg, seed=100 + di, domain_idx=di)
        x_THW = x.reshape(B, cfg.T, h, w)
        model.zero_grad()
        logits, _ = model.compute_logits(x_THW, action_ids=actions, domain=[dom] * B)
        relevant = x_THW[:, 1:] == cfg.mask_token_id
        loss, acc = model.compute_video_loss_and_acc(logits, labels, relevant)
        loss.backward()
        grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        rec = {
            "input_ids": x, "labels": labels, "actions": actions,
            "loss": loss.detach(), "acc": acc.detach(),
            # logits [B, 1024, T, H, W]: keep every 4th row/col (0.5 MB) + full last frame of sample 0
            "logits_sub": logits.detach()[:, :, :, ::4, ::4].clone(),
            "logits_b0_last": logits.detach()[0, :, -1].clone(),
            "grad_norms": {k: g.norm().item() for k, g in grads.items()},
            "grad_slices": {k: g.reshape(-1)[:: max(1, g.numel() // 64)][:64].clone() for k, g in grads.items()},
        }
        # MaskGIT decode of the last frame from the first T-1 frames (st_mask_git.py:337-467)
        for tag, steps, temp, mode in (("greedy1", 1, 0.0, "random"), ("greedy3", 3, 0.0, "greedy"),
                                       ("sample2", 2, 1.0, "random")):
            prompt = labels.reshape(B, cfg.T, h, w).clone()
            prompt[:, -1] = cfg.mask_token_id
            torch.manual_seed(777)
            with torch.no_grad():
                s, fl, _ = model.maskgit_generate(prompt, cfg.T - 1, maskgit_steps=steps, temperature=temp,
                                                  unmask_mode=mode, action_ids=actions, domain=[dom] * B)
            rec[f"gen_{tag}_samples"] = s.clone()
            rec[f"gen_{tag}_prompt_after"] = prompt[:, -1].clone()
            rec[f"gen_{tag}_logits_sub"] = fl[:, :, :, ::4, ::4].clone()
        out[dom] = rec
    # AR generate: 2 prompt frames -> 2 new frames, 2 steps, greedy tokens + random unmask order
    x, labels, actions = synthetic_batch(cfg, seed=100, domain_idx=0)
    torch.manual_seed(4242)
    with torch.no_grad():
        toks = model.generate(labels[:, : 2 * cfg.S], None, 2 * cfg.S, maskgit_steps=2, temperature=0.0,
                              action_ids=actions, domain=[DOMAINS[0]] * B, h=[h], w=[w])
    out["generate_tokens"] = toks.clone()
    return out


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    for name in VARIANTS:
        rec = run_variant(name)
        path = GOLDEN / f"{name}.pt"
        torch.save(rec, path)
        print(f"wrote {path} ({path.stat().st_size / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
