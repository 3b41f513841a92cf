"""Training collator (hma/data.py:28-98): the CPU oracle against fixtures produced by the real reference
(oracle/make_collator_golden.py, which also asserts oracle == live reference when it writes them), and the
on-device collator (hma_b200/data.py, csrc/collate.cu) against the oracle, bit-exact given the same random draws."""
import math
import random
from pathlib import Path

import pytest
import torch

from oracle import collator_oracle as C
from oracle import reference_loader

GOLDEN = Path(__file__).parent / "golden" / "collator.pt"


def _cfg(kw):
    from hma_b200 import GenieConfig
    return GenieConfig(**kw)


def test_collator_oracle_matches_reference_fixtures():
    rec = torch.load(GOLDEN, weights_only=False)
    assert set(rec) == {"mlm_corrupt", "non_mlm", "no_corruption", "one_vocab"}
    for name, r in rec.items():
        cfg = _cfg(r["cfg"])
        ids, labels = C.apply(r["tokens"], r["draws"], cfg, 16, 16)
        assert torch.equal(ids, r["input_ids"]), name
        assert torch.equal(labels, r["labels"]) and torch.equal(labels, r["tokens"]), name
        # replaying the seeds reproduces the draws (same consumption order of torch's generator and `random`)
        torch.manual_seed(r["seed"]); random.seed(r["seed"])
        d = C.draw(cfg, r["tokens"].shape[0], 16, 16)
        assert d["first_masked_frame"] == r["draws"]["first_masked_frame"]
        for k in ("corrupt_r", "rand_vals", "mask_r", "mask_prob", "frame_r"):
            assert (k in d) == (k in r["draws"])
            if k in d:
                assert torch.equal(d[k], r["draws"][k]), (name, k)


@pytest.mark.skipif(not reference_loader.available(), reason="live reference not present")
def test_collator_oracle_matches_live_reference_other_seed():
    from oracle.make_collator_golden import reference_collator
    GenieConfig, get_maskgit_collator = reference_collator()
    cfg = GenieConfig(num_layers=1, num_heads=8, d_model=256, T=6, S=256, num_factored_vocabs=2, non_mlm_ratio=0.5)
    for seed in (21, 22, 23, 24):
        g = torch.Generator().manual_seed(seed)
        tokens = torch.randint(0, 262144, (2, cfg.T * 256), generator=g)
        feats = [dict(input_ids=tokens[b].clone(), h=16, w=16, domain="d") for b in range(2)]
        torch.manual_seed(seed); random.seed(seed)
        ref = get_maskgit_collator(cfg)(feats)
        torch.manual_seed(seed); random.seed(seed)
        ids, labels = C.apply(tokens, C.draw(cfg, 2, 16, 16), cfg, 16, 16)
        assert torch.equal(ids, ref["input_ids"]) and torch.equal(labels, ref["labels"])


@pytest.mark.skipif(not reference_loader.available(), reason="live reference not present")
def test_collator_without_masking_returns_the_original_tokens_like_the_reference():
    """dataloader_apply_mask=False: the reference folds the corrupted factors back only inside `if apply_mask` (data.py:69-83),
    so input_ids stay the original tokens even with corruption on."""
    from oracle.make_collator_golden import reference_collator
    GenieConfig, get_maskgit_collator = reference_collator()
    cfg = GenieConfig(num_layers=1, num_heads=8, d_model=256, T=6, S=256, num_factored_vocabs=2, non_mlm_ratio=0.5,
                      dataloader_apply_mask=False, dataloader_apply_corruption=True)
    for seed in (31, 32):
        g = torch.Generator().manual_seed(seed)
        tokens = torch.randint(0, 262144, (2, cfg.T * 256), generator=g)
        feats = [dict(input_ids=tokens[b].clone(), h=16, w=16, domain="d") for b in range(2)]
        torch.manual_seed(seed); random.seed(seed)
        ref = get_maskgit_collator(cfg)(feats)
        torch.manual_seed(seed); random.seed(seed)
        ids, labels = C.apply(tokens, C.draw(cfg, 2, 16, 16), cfg, 16, 16)
        assert torch.equal(ref["input_ids"], tokens) and torch.equal(ids, tokens) and torch.equal(labels, ref["labels"])


@pytest.mark.gpu
def test_device_collator_without_masking_returns_the_original_tokens():
    from hma_b200 import data
    cfg = _cfg(dict(num_layers=1, num_heads=8, d_model=256, T=6, S=256, num_factored_vocabs=2, non_mlm_ratio=0.5,
                    dataloader_apply_mask=False, dataloader_apply_corruption=True))
    tokens = torch.randint(0, 262144, (2, cfg.T * 256)).cuda()
    random.seed(3)
    ids, labels = data.collate_from_draws(tokens, data.draw_on_device(cfg, 2, 16, 16, tokens.device), cfg, 16, 16)
    assert torch.equal(ids, tokens) and torch.equal(labels, tokens)


@pytest.mark.gpu
def test_device_collator_bit_exact_given_reference_draws():
    from hma_b200 import data
    rec = torch.load(GOLDEN, weights_only=False)
    for name, r in rec.items():
        cfg = _cfg(r["cfg"])
        ids, labels = data.collate_from_draws(r["tokens"].cuda(), r["draws"], cfg, 16, 16)
        assert torch.equal(ids.cpu(), r["input_ids"]), name
        assert torch.equal(labels.cpu(), r["labels"]), name


@pytest.mark.gpu
def test_device_collator_api_and_distribution_full_size():
    """get_maskgit_collator(config) drop-in at the config-2 batch shape: same dict keys, labels untouched, frame 0 never
    masked, masked fraction ~ (2/pi)(T-1)/T in the MLM branch, oracle agrees on the very draws used."""
    from hma_b200 import data
    cfg = _cfg(dict(num_layers=1, num_heads=8, d_model=256, T=16, S=256, num_factored_vocabs=2, non_mlm_ratio=0.0))
    g = torch.Generator().manual_seed(0)
    feats = [dict(input_ids=torch.randint(0, 262144, (16 * 256,), generator=g), h=16, w=16, domain="dom", action_ids=torch.randn(16, 7, generator=g))
             for _ in range(8)]
    collate = data.get_maskgit_collator(cfg)
    torch.manual_seed(1); random.seed(1)
    fracs = []
    for _ in range(20):
        out = collate(feats)
        assert set(out) == {"input_ids", "labels", "action_ids", "domain", "h", "w"}
        assert out["input_ids"].shape == (8, 4096) and out["input_ids"].is_cuda and out["action_ids"].shape == (8, 16, 7)
        assert torch.equal(out["labels"].cpu(), torch.stack([f["input_ids"] for f in feats]))
        x = out["input_ids"].view(8, 16, 256)
        assert (x[:, 0] != 262144).all()
        fracs.append((x == 262144).float().mean().item())
    mean = sum(fracs) / len(fracs)
    assert abs(mean - (2 / math.pi) * 15 / 16) < 0.06, mean
    # the device path and the oracle agree on whatever draws the device made
    tokens = torch.stack([f["input_ids"] for f in feats])
    d = data.draw_on_device(cfg, 8, 16, 16, torch.device("cuda"))
    ids, _ = data.collate_from_draws(tokens.cuda(), d, cfg, 16, 16)
    d_cpu = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in d.items()}
    ref_ids, _ = C.apply(tokens, d_cpu, cfg, 16, 16)
    assert torch.equal(ids.cpu(), ref_ids)
