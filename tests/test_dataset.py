"""hma_b200.dataset.RawTokenDataset against outputs of the reference class (tests/golden/rawtoken.pt, made by
oracle/make_rawtoken_golden.py) on the regenerated synthetic directory; the device gather against the host path."""
from pathlib import Path

import numpy as np
import pytest
import torch

from tests import _rawdata

GOLDEN = Path(__file__).parent / "golden" / "rawtoken.pt"


@pytest.fixture(scope="module")
def root(tmp_path_factory):
    return _rawdata.write(tmp_path_factory.mktemp("rawtoken") / "ds", seed=0)


@pytest.mark.parametrize("case", list(_rawdata.CASES))
def test_matches_reference_fixture(root, case):
    from hma_b200.dataset import RawTokenDataset

    ref = torch.load(GOLDEN, weights_only=False)[case]
    ds = RawTokenDataset(root, freq_table=_rawdata.FREQ, **_rawdata.CASES[case])
    assert ds.stride == ref["stride"] and ds.n_action == ref["n_action"] and ds.num_videos == ref["num_videos"]
    assert len(ds) == ref["len"] and list(ds.valid_start_inds) == ref["valid_start_inds"]
    np.random.seed(0)
    items = [ds[i] for i in ref["idx"]]
    assert torch.equal(torch.stack([it["input_ids"] for it in items]), ref["input_ids"])
    assert items[0]["domain"] == ref["domain"] and items[0]["h"] == 16 and items[0]["labels"] is items[0]["input_ids"]
    if ref["action_ids"] is not None:
        assert torch.equal(torch.stack([it["action_ids"] for it in items]), ref["action_ids"])
        assert ds.action_stat == ref["action_stat"]
    else:
        assert "action_ids" not in items[0]


def test_edge_cases(tmp_path):
    from hma_b200.dataset import RawTokenDataset

    # a table shorter than one window: empty dataset, no error (the reference's range() is empty too)
    r = _rawdata.write(tmp_path / "short", seed=1, num_images=10)
    assert len(RawTokenDataset(r, window_size=16)) == 0
    # no segment ids: only usable without interrupt filtering (data.py:225-227)
    (r / "segment_ids.bin").unlink()
    with pytest.raises(NotImplementedError):
        RawTokenDataset(r, window_size=2)
    assert len(RawTokenDataset(r, window_size=2, filter_interrupts=False, freq_table=_rawdata.FREQ)) == 10 - 3 - 3
    with pytest.raises(RuntimeError, match="to_device"):
        RawTokenDataset(r, window_size=2, filter_interrupts=False).gather([0])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["uint32", "uint16"])
def test_device_gather_equals_host_items(tmp_path, dtype):
    from hma_b200.dataset import RawTokenDataset

    r = _rawdata.write(tmp_path / dtype, seed=2, num_images=300, token_dtype=dtype)
    ds = RawTokenDataset(r, window_size=4, use_actions=True, freq_table=_rawdata.FREQ).to_device("cuda")
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, len(ds), (33,), generator=g)
    out = ds.gather(idx)
    items = [ds[int(i)] for i in idx]
    assert torch.equal(out["input_ids"].cpu(), torch.stack([it["input_ids"] for it in items]))
    assert torch.equal(out["action_ids"].cpu(), torch.stack([it["action_ids"] for it in items]))
    assert out["labels"] is out["input_ids"] and out["domain"] == ["synthetic_robot"] * 33 and out["h"] == [16] * 33
    # feeds the on-device collator unchanged
    from hma_b200 import GenieConfig
    from hma_b200.data import collate_from_draws, draw_on_device

    # non_mlm_ratio=0: with T = 4 and the default num_prompt_frames = 4 the non-MLM branch would call random.randint(4, 3)
    cfg = GenieConfig(num_layers=1, num_heads=8, d_model=256, T=4, S=256, num_factored_vocabs=2, non_mlm_ratio=0.0)
    if dtype == "uint32":
        ids, labels = collate_from_draws(out["input_ids"], draw_on_device(cfg, 33, 16, 16, torch.device("cuda")), cfg, 16, 16)
        assert torch.equal(labels, out["input_ids"]) and (ids == cfg.image_vocab_size).any()


@pytest.fixture(scope="module")
def feature_root(tmp_path_factory):
    return _rawdata.write(tmp_path_factory.mktemp("rawfeat") / "feat", seed=3, token_dtype="float16", latent_channels=4, h=8, w=8)


@pytest.mark.parametrize("case", list(_rawdata.FEATURE_CASES))
def test_feature_dataset_and_collator_match_reference_fixture(feature_root, case):
    """RawFeatureDataset (hma/data.py:297-435) and get_maskgit_collator_feature (:100-157), the STMAR data path: same windows,
    same items, and — seeding torch and `random` as the fixture did — the same masked_tokens_indicator."""
    import random

    from hma_b200.data import get_maskgit_collator_feature
    from hma_b200.dataset import RawFeatureDataset
    from hma_b200.mar import DiffusionGenieConfig

    ref = torch.load(GOLDEN, weights_only=False)["feature_" + case]
    ds = RawFeatureDataset(feature_root, freq_table=_rawdata.FREQ, **_rawdata.FEATURE_CASES[case])
    assert ds.stride == ref["stride"] and ds.n_action == ref["n_action"]
    assert len(ds) == ref["len"] and list(ds.valid_start_inds) == ref["valid_start_inds"]
    items = [ds[i] for i in ref["idx"]]
    assert torch.equal(torch.stack([it["input_ids"] for it in items]), ref["input_ids"])
    assert items[0]["domain"] == ref["domain"] and items[0]["c"] == ref["c"] and "_noquant" not in items[0]["domain"]
    if ref["action_ids"] is not None:
        assert torch.equal(torch.stack([it["action_ids"] for it in items]), ref["action_ids"])
    for tag in ("mlm", "non_mlm"):
        c = ref[f"collate_{tag}"]
        cfg = DiffusionGenieConfig(num_layers=1, num_heads=8, d_model=256, **c["cfg"])
        torch.manual_seed(c["seed"])
        random.seed(c["seed"])
        b = get_maskgit_collator_feature(cfg)(items)
        assert torch.equal(b["masked_tokens_indicator"], c["masked_tokens_indicator"])
        assert tuple(b["input_ids"].shape) == c["input_ids_shape"]
        assert torch.equal(b["input_ids"], torch.stack([it["input_ids"] for it in items]))
        assert torch.equal(b["labels"], b["input_ids"]) and b["labels"].data_ptr() != b["input_ids"].data_ptr()
        assert b["domain"] == [ref["domain"]] * len(items) and b["h"] == [8] * len(items)


@pytest.mark.gpu
def test_device_batch_pipeline_end_to_end(tmp_path):
    """MultiTaskBatchSampler -> device gather -> on-device collator (DeviceBatchPipeline), two datasets of different action
    widths: every batch comes from ONE dataset (= one action domain, external/data_sampler.py:177-303), its labels are
    exactly the host-path windows of the sampled indices, frame 0 and the prompt frames of the non-MLM branch are never
    masked, and a 2-layer model trains on the stream through TrainStep."""
    import random

    from hma_b200 import GenieConfig, STMaskGIT
    from hma_b200.dataset import RawTokenDataset
    from hma_b200.sampler import DeviceBatchPipeline
    from hma_b200.train import TrainStep

    T = 6
    roots = [_rawdata.write(tmp_path / "a", seed=5, num_images=400, action_dim=7),
             _rawdata.write(tmp_path / "b", seed=6, num_images=300, action_dim=3)]
    names = ["robot_a", "robot_b"]
    dsets = [RawTokenDataset(r, window_size=T, use_actions=True, name=n, freq_table={}).to_device("cuda") for r, n in zip(roots, names)]
    # num_prompt_frames < T so that the non-MLM branch's random.randint(num_prompt_frames, T - 1) is valid (data.py:55)
    cfg = GenieConfig(num_layers=2, num_heads=8, d_model=256, T=T, S=256, num_factored_vocabs=2, non_mlm_ratio=0.5,
                      num_prompt_frames=2, dataloader_apply_corruption=True, action_network="concat+modulate")
    pipe = DeviceBatchPipeline(dsets, cfg, batch_size=4, seed=3)
    random.seed(0)
    torch.manual_seed(0)
    plan = list(pipe.sampler.iter_tasks())
    seen = set()
    batches = []
    for (task, local), batch in zip(plan, pipe):
        ds = dsets[task]
        seen.add(task)
        assert batch["domain"] == [names[task]] * 4 and batch["input_ids"].is_cuda and batch["input_ids"].shape == (4, T * 256)
        want = torch.stack([ds[int(i)]["input_ids"] for i in local])
        assert torch.equal(batch["labels"].cpu(), want)
        assert batch["action_ids"].shape == (4, T, ds.n_action)
        x = batch["input_ids"].view(4, T, 256)
        assert (x[:, 0] != cfg.image_vocab_size).all() and (x == cfg.image_vocab_size).any()
        batches.append(batch)
        if len(batches) == 12:
            break
    assert seen == {0, 1}
    torch.manual_seed(0)
    with torch.device("cuda"):
        model = STMaskGIT(cfg)
        model.init_action_projectors(names, [d.n_action for d in dsets], [d.action_stat for d in dsets], "concat+modulate")
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() >= 2:
                p.normal_(0.0, 0.05)
    step = TrainStep(model, lr=1e-3, weight_decay=0.0)
    passes = []
    for rep in range(4):  # the same 12 batches every pass, so the pass means are comparable (single losses are not:
        tot = 0.0         # the masked fraction differs from batch to batch)
        for b in batches:
            loss = step(b["input_ids"], b["labels"], b["action_ids"], b["domain"])[0].item()
            assert loss == loss
            tot += loss
        passes.append(tot / len(batches))
    assert passes[-1] < passes[0], passes
