"""CPU-side checks of the STMAR drop-in (no GPU): host logic against the oracle (which is pinned on the live reference),
state_dict layout, loud failures for what is not implemented and for CPU tensors."""
import numpy as np
import pytest
import torch

from oracle import stmar_oracle as M
from tests.test_mar_oracle import mar_golden


def _model(rec, **over):
    from hma_b200 import STMAR, DiffusionGenieConfig

    kw = dict(rec["kw"])
    kw.update(over)
    m = STMAR(DiffusionGenieConfig(**kw))
    stats = [[[0.0] * a, [1.0] * a] for a in rec["action_dims"]]
    m.init_action_projectors(rec["domains"], rec["d_actions"], stats, kw["action_network"])
    return m


@pytest.mark.parametrize("respacing", [None, "100", "20", "250", "10,20,30"])
def test_diffusion_tables_match_oracle(respacing):
    """gaussian_diffusion.py:149-186 / respace.py:72-93 tables, as uploaded to the device."""
    from hma_b200.mar import diffusion_tables, space_timesteps

    tb, tmap = diffusion_tables(respacing)
    o = M.Tables(respacing)
    assert tmap == o.timestep_map
    assert torch.equal(tb, o.packed())
    if respacing:
        assert space_timesteps(1000, respacing) == M.space_timesteps(1000, respacing)


def test_space_timesteps_errors_like_the_reference():
    from hma_b200.mar import space_timesteps

    with pytest.raises(ValueError):  # respace.py:46-47
        space_timesteps(10, "20")
    with pytest.raises(NotImplementedError):
        space_timesteps(1000, "ddim50")


def test_mask_schedule_and_orders_match_oracle():
    from hma_b200 import STMAR

    for k in (1, 2, 3, 8, 16, 64):
        assert STMAR.mask_schedule(64, k) == M.mask_schedule(64, k)
    rec, cfg, sd = mar_golden()
    m = _model(rec)
    np.random.seed(5)
    got = m.sample_orders(2)
    np.random.seed(5)
    assert torch.equal(got, M.sample_orders(2, cfg.seq_len))
    assert torch.equal(got, rec[rec["domains"][0]]["gen_orders"])  # the draws the reference made under the same seed


def test_state_dict_layout_and_patchify():
    rec, cfg, sd = mar_golden()
    m = _model(rec)
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    assert res.missing_keys and all(k.startswith("action_diff_losses.") for k in res.missing_keys)
    own = m.state_dict()
    for k, v in sd.items():
        assert own[k].shape == v.shape, k
    # the per-domain action heads the reference builds (st_mar.py:92-104): d_action-wide input / 2*d_action-wide output
    assert own["action_diff_losses.dom00.net.input_proj.weight"].shape == (1024, 14)
    assert own["action_diff_losses.dom01.net.final_layer.linear.weight"].shape == (20, 1024)
    x = torch.randn(2, 3, 16, 16, 4)
    assert torch.equal(m.patchify(x), M.patchify(x, 2))
    assert torch.equal(m.unpatchify(m.patchify(x)), x)
    assert torch.equal(m.unpatchify(M.patchify(x, 2)), M.unpatchify(M.patchify(x, 2), 2, 4))


def test_unsupported_configurations_fail_loudly():
    rec, _, _ = mar_golden()
    for over in (dict(diffloss_w=512), dict(jointly_predict_actions=True), dict(diffusion_batch_mul=4), dict(d_model=512),
                 dict(action_network="cross_attention"), dict(vae_embed_dim=32)):
        with pytest.raises(NotImplementedError):
            _model(rec, **over)


def test_no_cpu_fallback():
    rec, cfg, sd = mar_golden()
    m = _model(rec)
    r = rec[rec["domains"][0]]
    with pytest.raises(RuntimeError, match="CUDA device only"):
        m(r["latents"].clone(), r["latents"], action_ids=r["actions"], domain=["dom00", "dom00"],
          masked_tokens_indicator=r["mask"], h=[16], w=[16])
    with pytest.raises(RuntimeError, match="CUDA device only"):
        m.maskgit_generate(r["gen_prompt"], cfg.T - 1, action_ids=r["actions"], domain=["dom00", "dom00"], maskgit_steps=1)
    with pytest.raises(NotImplementedError):
        m.maskgit_generate(r["gen_prompt"], cfg.T - 1, cfg=2.0)


def test_param_arena_layout_for_stmar():
    """MarTrainStep's flat parameter arena (train.ParamArena) on CPU: every parameter becomes a view of the arena with its
    values intact, the shared block holds trunk + front end + latent head + diffusion MLP, each action domain has its own
    contiguous block, and the never-executed per-domain action heads sit behind them."""
    from hma_b200.train import ParamArena

    rec, cfg, sd = mar_golden()
    m = _model(rec)
    m.load_state_dict(sd, strict=False)
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    arena = ParamArena(m)
    named = dict(m.named_parameters())
    flat = arena.flat
    for k, p in named.items():
        off = arena.offsets[k]
        assert p.data_ptr() == flat.data_ptr() + 4 * off and off % 4 == 0, k
        assert torch.equal(p.detach(), before[k]), k
    assert set(arena.dom_range) == set(rec["domains"])
    shared = [k for k, o in arena.offsets.items() if o < arena.shared_size]
    assert any(k.startswith("diffloss.net.res_blocks.1.") for k in shared) and "mask_token" in shared
    assert "diffusion_pos_embed_learned" in shared and "decoder_norm.weight" in shared and "z_proj_ln.bias" in shared
    assert not any("action_projectors" in k or k.startswith("action_mlp.") or k.startswith("action_diff_losses.") for k in shared)
    end_of_domains = max(lo + n for lo, n in arena.dom_range.values())
    for dom, (lo, n) in arena.dom_range.items():
        mine = [k for k, o in arena.offsets.items() if lo <= o < lo + n]
        assert mine and all(dom in k for k in mine), dom
        assert any("adaLN_modulation" in k for k in mine) and any(k.startswith(f"action_mlp.{dom}.") for k in mine)
    left = [k for k, o in arena.offsets.items() if o >= end_of_domains]
    assert left and all(k.startswith("action_diff_losses.") or k == "action_mask_tokens" for k in left), left[:5]
    # the engine's gradient buffer uses the same intra-block order as the arena (optimizer = streaming kernels over ranges)
    eng = m._engine
    d = eng.mar_dims(2, cfg.T, 16, 16, True)
    names = eng.active_param_names(named, d, rec["domains"][1], True)
    offs = [arena.offsets[k] for k in names]
    lo1 = arena.dom_range[rec["domains"][1]][0]
    rel = [o if o < arena.shared_size else o - lo1 + arena.shared_size for o in offs]
    assert rel == sorted(rel) and rel[0] == 0
    g = eng.alloc_grads(named, d, rec["domains"][1], True, torch.device("cpu"))
    base = g[names[0]].data_ptr()
    assert [(g[k].data_ptr() - base) // 4 for k in names] == rel
