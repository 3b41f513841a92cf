"""Generate tests/golden/tiny_mar.pt by running the REAL reference STMAR (/root/reference) in this container.

    python -m oracle.make_mar_golden

TEST INFRASTRUCTURE ONLY (see oracle/make_golden.py). The reference model runs in eval mode (mlp_drop inactive);
its hard-coded `.cuda()` calls in the sampling path (st_mar.py:20-22,354,400; diffloss.py:40,45;
gaussian_diffusion.py:467,477) are neutralised by making Tensor.cuda the identity for the duration of the call.
Every random draw is recorded in the fixture: the diffusion timesteps and noise of the training loss, the numpy
generation orders and the sampling noise — so that the oracle and the CUDA path can be fed the same tensors.
"""
from __future__ import annotations

import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import reference_loader  # noqa: E402
from oracle import stmar_oracle as M  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
KW = dict(num_layers=2, num_heads=8, d_model=256, T=4, S=256, use_mup=False, qk_norm=False, qkv_bias=True, proj_bias=True,
          mlp_bias=False, action_network="concat+modulate", patch_size=2, diffloss_d=2, diffloss_w=1024,
          num_sampling_steps="20", num_factored_vocabs=2)
DOMAINS = ["dom00", "dom01"]
D_ACTIONS = [14, 10]
ACTION_DIMS = [7, 10]
B = 2
H = W = 16


def build_reference():
    reference_loader.load()
    with contextlib.redirect_stdout(io.StringIO()):
        from hma.config import DiffusionGenieConfig
        from hma.model.st_mar import STMAR

        rcfg = DiffusionGenieConfig(mlp_drop=0.05, attn_drop=0.1, **KW)
        model = STMAR(rcfg)
        stats = [[[0.0] * a, [1.0] * a] for a in ACTION_DIMS]
        model.init_action_projectors(DOMAINS, D_ACTIONS, stats, rcfg.action_network)
    cfg = M.MarConfig(**KW)
    sd = M.make_state_dict(cfg, DOMAINS, D_ACTIONS, seed=0, action_dims=ACTION_DIMS)
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith("action_diff_losses.") for k in res.missing_keys), res.missing_keys
    return model.eval(), cfg, sd


def synthetic_batch(cfg, seed: int, di: int):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(B, cfg.T * H * W, cfg.vae_embed_dim, generator=g) * 0.8
    mask = torch.zeros(B, cfg.T, H, W, dtype=torch.bool)
    for b in range(B):
        for t in range(1, cfg.T):
            mask[b, t] = torch.rand(H, W, generator=g) < 0.25 * t
    actions = torch.randn(B, cfg.T, D_ACTIONS[di], generator=g)
    return lat, mask, actions


@contextlib.contextmanager
def cuda_is_identity():
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def main():
    torch.set_num_threads(8)
    model, cfg, sd = build_reference()
    out = {"kw": KW, "domains": DOMAINS, "d_actions": D_ACTIONS, "action_dims": ACTION_DIMS, "seed": 0}
    N = B * cfg.T * cfg.seq_len
    for di, dom in enumerate(DOMAINS):
        lat, mask, actions = synthetic_batch(cfg, 200 + di, di)
        model.zero_grad()
        torch.manual_seed(31 + di)
        res = model(lat.clone(), lat.clone(), action_ids=actions, domain=[dom] * B, masked_tokens_indicator=mask,
                    h=[H], w=[W])
        res.loss.backward()
        torch.manual_seed(31 + di)  # the two draws of DiffLoss.forward / training_losses, in order
        t = torch.randint(0, 1000, (N,))
        noise = torch.randn(N, cfg.token_dim)
        # the t == 0 branch (decoder NLL) is rare under randint: a second loss evaluation forces it on some rows
        t0 = t.clone()
        t0[::5] = 0
        tgt = M.patchify(lat.reshape(B, cfg.T, H, W, -1), 2).reshape(N, -1)
        m = (M.patchify(mask[..., None], 2).sum(-1) > 0).reshape(-1).float()
        zf = res.logits.detach().permute(0, 2, 3, 4, 1).reshape(N, -1)
        with torch.no_grad():
            ld = model.diffloss.train_diffusion.training_losses(model.diffloss.net, tgt, t0, dict(c=zf), noise=noise)["loss"]
        grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        rec = {"latents": lat, "mask": mask, "actions": actions, "t": t, "noise": noise, "loss": res.loss.detach(),
               "z": res.logits.detach().clone(),  # [B, d, T, 8, 8]
               "t_forced0": t0, "rows_forced0": ld.clone(),
               "grad_norms": {k: g.norm().item() for k, g in grads.items()},
               "grad_slices": {k: g.reshape(-1)[:: max(1, g.numel() // 64)][:64].clone() for k, g in grads.items()}}
        # generation of the last frame from the first T-1 (st_mar.py:357-454)
        prompt = lat.reshape(B, cfg.T, H, W, -1).clone()
        prompt[:, -1] = sd["mask_token"].reshape(-1)
        np.random.seed(5 + di)
        torch.manual_seed(77 + di)
        with cuda_is_identity(), torch.no_grad():
            frame, z0, _ = model.maskgit_generate(prompt.clone(), cfg.T - 1, action_ids=actions, domain=[dom] * B,
                                                  maskgit_steps=3, temperature=0.9)
        np.random.seed(5 + di)
        orders = []
        for _ in range(B):
            o = np.array(list(range(cfg.seq_len)))
            np.random.shuffle(o)
            orders.append(o)
        rec.update(gen_prompt=prompt, gen_orders=torch.tensor(np.array(orders)).long(), gen_frame=frame.clone(),
                   gen_z0=z0.clone(), gen_seed=77 + di, gen_steps=3, gen_temperature=0.9)
        out[dom] = rec
    # AR generate: 2 prompt frames -> 2 new frames (st_mar.py:273-345)
    lat, mask, actions = synthetic_batch(cfg, 200, 0)
    model.maskgit_steps = 2
    np.random.seed(11)
    torch.manual_seed(12)
    with cuda_is_identity(), torch.no_grad():
        toks = model.generate(lat[:, : 2 * H * W].clone(), None, 2 * H * W, temperature=1.0, action_ids=actions,
                              domain=[DOMAINS[0]] * B, h=[H], w=[W])
    out["generate"] = {"latents": lat, "actions": actions, "np_seed": 11, "torch_seed": 12, "maskgit_steps": 2,
                       "out": toks.clone()}
    GOLDEN.mkdir(parents=True, exist_ok=True)
    path = GOLDEN / "tiny_mar.pt"
    torch.save(out, path)
    print(f"wrote {path} ({path.stat().st_size / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
