"""Pins the STMAR CPU oracle (oracle/stmar_oracle.py) against outputs of the real reference recorded in
tests/golden/tiny_mar.pt (made by oracle/make_mar_golden.py). CPU only."""
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import stmar_oracle as M

GOLDEN = Path(__file__).parent / "golden"
H = W = 16


def mar_golden():
    rec = torch.load(GOLDEN / "tiny_mar.pt", weights_only=False)
    cfg = M.MarConfig(**rec["kw"])
    sd = M.make_state_dict(cfg, rec["domains"], rec["d_actions"], seed=rec["seed"], action_dims=rec["action_dims"])
    return rec, cfg, sd


def test_tables_match_published_schedule():
    """Known answers of the cosine schedule (gaussian_diffusion.py:112-138) and of '100'-step respacing
    (respace.py:12-62): beta_0, the 0.999 cap, and the kept timesteps."""
    tb = M.Tables()
    assert tb.num_timesteps == 1000
    ab = lambda u: math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
    assert math.isclose(tb.betas[0], 1 - ab(0.001) / ab(0.0), rel_tol=1e-12)
    assert tb.betas[-1] == 0.999
    assert np.all(np.diff(tb.sqrt_acp) < 0)
    ts = M.space_timesteps(1000, "100")
    assert len(ts) == 100 and ts[0] == 0 and ts[-1] == 999
    sp = M.Tables("100")
    assert sp.num_timesteps == 100 and sp.timestep_map == ts
    # respaced cumulative products coincide with the base ones at the kept steps
    np.testing.assert_allclose(np.cumprod(1 - sp.betas), np.cumprod(1 - tb.betas)[ts], rtol=1e-9)


def test_oracle_forward_backward_matches_reference_fixture():
    rec, cfg, sd = mar_golden()
    torch.set_num_threads(8)
    for dom in rec["domains"]:
        r = rec[dom]
        params = {k: v.clone().requires_grad_("action_preprocessor" not in k) for k, v in sd.items()}
        loss, z = M.forward(r["latents"], r["latents"], r["mask"], r["actions"], [dom, dom], params, cfg, r["t"], r["noise"], H, W)
        assert torch.allclose(loss, r["loss"].reshape(()), rtol=2e-5, atol=1e-6), (loss.item(), r["loss"].item())
        z_ref = r["z"].permute(0, 2, 3, 4, 1).reshape(z.shape)
        torch.testing.assert_close(z, z_ref, rtol=1e-4, atol=1e-4)
        loss.backward()
        for k, gn in r["grad_norms"].items():
            g = params[k].grad
            assert g is not None, k
            assert math.isclose(g.norm().item(), gn, rel_tol=5e-4, abs_tol=1e-7), (k, g.norm().item(), gn)
            sl = g.reshape(-1)[:: max(1, g.numel() // 64)][:64]
            torch.testing.assert_close(sl, r["grad_slices"][k], rtol=3e-3, atol=1e-6)
        for k, p in params.items():
            if p.requires_grad and k not in r["grad_norms"]:
                assert p.grad is None or p.grad.abs().max() == 0, k


def test_oracle_row_losses_with_decoder_nll_rows():
    """Rows with t == 0 take the discretised-Gaussian NLL branch (gaussian_diffusion.py:663-672)."""
    rec, cfg, sd = mar_golden()
    dom = rec["domains"][0]
    r = rec[dom]
    B, T = 2, cfg.T
    N = B * T * cfg.seq_len
    tb = M.Tables()
    tgt = M.patchify(r["latents"].reshape(B, T, H, W, -1), 2).reshape(N, -1)
    zf = r["z"].permute(0, 2, 3, 4, 1).reshape(N, -1)
    t0 = r["t_forced0"]
    x_t = M._ex(tb.sqrt_acp, t0) * tgt + M._ex(tb.sqrt_1m_acp, t0) * r["noise"]
    out = M.mlp_adaln(x_t, t0, zf, sd, "diffloss.net.", cfg.diffloss_d)
    rows = M.diffusion_row_losses(out, tgt, r["noise"], t0, tb)
    assert (t0 == 0).sum() >= N // 5
    torch.testing.assert_close(rows, r["rows_forced0"], rtol=2e-4, atol=1e-5)


def test_oracle_generation_matches_reference_fixture():
    rec, cfg, sd = mar_golden()
    for dom in rec["domains"]:
        r = rec[dom]
        torch.manual_seed(r["gen_seed"])
        frame, z0 = M.maskgit_generate(r["gen_prompt"], cfg.T - 1, r["gen_orders"], lambda s: torch.randn(*s), sd, cfg,
                                       action_ids=r["actions"], domain=[dom, dom], maskgit_steps=r["gen_steps"],
                                       temperature=r["gen_temperature"])
        z0_ref = r["gen_z0"].permute(0, 2, 3, 1).reshape(z0.shape)
        torch.testing.assert_close(z0, z0_ref, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(frame, r["gen_frame"], rtol=2e-3, atol=2e-3)


def test_mask_schedule_known_answers():
    # st_mar.py:393-400 with seq_len 64: floor(64*cos(pi/2*(k+1)/K)) clamped to [1, 63]
    assert M.mask_schedule(64, 1) == [1]
    assert M.mask_schedule(64, 3) == [55, 32, 1]
    assert M.mask_schedule(64, 16)[0] == 63 and M.mask_schedule(64, 16)[-1] == 1


def test_oracle_ar_generate_matches_reference_fixture():
    rec, cfg, sd = mar_golden()
    g = rec["generate"]
    cfg.maskgit_steps = g["maskgit_steps"]
    np.random.seed(g["np_seed"])
    torch.manual_seed(g["torch_seed"])
    dom = rec["domains"][0]
    out = M.generate(g["latents"][:, : 2 * H * W], 2 * H * W, lambda s: torch.randn(*s), sd, cfg, H, W, action_ids=g["actions"],
                     domain=[dom, dom], temperature=1.0)
    torch.testing.assert_close(out, g["out"], rtol=2e-3, atol=2e-3)
