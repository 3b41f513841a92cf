"""Shared helpers for the parity tests (oracle side is test infrastructure only)."""
import math
from pathlib import Path

import torch

from oracle import stmaskgit_oracle as O

GOLDEN = Path(__file__).parent / "golden"
MAGVIT_TINY = dict(num_layers=2, num_heads=8, d_model=256, T=4, S=256, use_mup=False, qk_norm=False, qkv_bias=False,
                   action_network="concat+modulate")


def golden(name="tiny_magvit"):
    rec = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    cfg = O.OracleConfig(num_factored_vocabs=2, **MAGVIT_TINY)
    sd = O.make_state_dict(cfg, rec["domains"], rec["d_actions"], seed=rec["seed"], action_dims=rec["action_dims"])
    return rec, cfg, sd


def build_cuda_model(rec, sd, device="cuda", **overrides):
    from hma_b200 import GenieConfig, STMaskGIT

    kw = dict(MAGVIT_TINY)
    kw.update(overrides)
    cfg = GenieConfig(num_factored_vocabs=2, **kw)
    model = STMaskGIT(cfg)
    stats = [[[0.0] * a, [1.0] * a] for a in rec["action_dims"]]
    model.init_action_projectors(rec["domains"], rec["d_actions"], stats, cfg.action_network)
    model.load_state_dict(sd, strict=True)
    return model.to(device)
