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
