"""CPU-side checks of the drop-in boundary: state_dict layout, config round trip, C-ABI exports."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from tests._util import build_cuda_model, golden

ROOT = Path(__file__).resolve().parent.parent


def test_state_dict_layout_matches_reference_layout():
    """oracle.make_state_dict is proven to load strictly into the reference (oracle/make_golden.py);
    the CUDA model must expose exactly the same keys and shapes."""
    rec, cfg, sd = golden()
    model = build_cuda_model(rec, sd, device="cpu")
    mine = model.state_dict()
    assert set(mine) == set(sd)
    for k, v in sd.items():
        assert tuple(mine[k].shape) == tuple(v.shape), k
        assert torch.equal(mine[k], v), k


def test_config_roundtrip(tmp_path):
    from hma_b200 import GenieConfig

    c = GenieConfig(num_layers=32, num_heads=8, d_model=256, num_factored_vocabs=2, action_network="concat+modulate")
    assert c.factored_vocab_size == 512
    c.save_pretrained(tmp_path / "c.json")
    c2 = GenieConfig.from_pretrained(tmp_path / "c.json")
    assert vars(c) == vars(c2)


def test_unsupported_configs_fail_loudly():
    from hma_b200 import GenieConfig, STMaskGIT

    with pytest.raises(NotImplementedError):
        STMaskGIT(GenieConfig(num_layers=1, num_heads=8, d_model=256, num_factored_vocabs=2, action_network="cross_attention"))
    with pytest.raises(NotImplementedError):
        STMaskGIT(GenieConfig(num_layers=1, num_heads=8, d_model=512, num_factored_vocabs=2, qk_norm=False))


def test_no_cpu_fallback():
    rec, cfg, sd = golden()
    model = build_cuda_model(rec, sd, device="cpu")
    r = rec[rec["domains"][0]]
    with pytest.raises(RuntimeError, match="CUDA device only"):
        model(r["input_ids"], r["labels"], action_ids=r["actions"], domain=[rec["domains"][0]] * 2)


def test_c_abi_exports_every_declared_symbol():
    from hma_b200 import build

    lib_path = build.build()
    lib = ctypes.CDLL(str(lib_path))
    header = (ROOT / "include" / "hma_b200.h").read_text()
    names = sorted(set(re.findall(r"\b(hma_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hma_b200.h but not exported"
    assert lib.hma_abi_version() == 3


def test_forward_consumes_cpu_rng_like_the_reference(monkeypatch):
    """st_mask_git.py:707-708: two CPU torch.rand draws per forward with actions. Checked on
    the CPU-side bookkeeping only (the CUDA path is patched out)."""
    import torch
    from hma_b200 import GenieConfig, STMaskGIT
    from hma_b200 import model as M

    cfg = GenieConfig(num_layers=1, num_heads=8, d_model=256, T=3, S=256, num_factored_vocabs=2, qk_norm=False,
                      action_network="concat+modulate")
    m = STMaskGIT(cfg)
    m.init_action_projectors(["a"], [4], [[[0.0] * 4, [1.0] * 4]], "concat+modulate")
    m._require_cuda = lambda t: None
    m._logits_nograd = lambda *a, **k: (torch.zeros(2 * 3 * 256, 1024), None)
    monkeypatch.setattr(M.ops, "ce_fwd", lambda *a, **k: (torch.zeros(2), None, None))
    ids = torch.zeros(2, 3 * 256, dtype=torch.long)
    acts = torch.zeros(2, 3, 4)
    torch.manual_seed(5)
    with torch.no_grad():
        m(ids, ids, action_ids=acts, domain=["a", "a"])
    after = torch.rand(1)
    torch.manual_seed(5)
    torch.rand(2, 1, 1)
    torch.rand(2, 3, 1)
    assert torch.equal(after, torch.rand(1))
    assert m.relevant_action_mask.shape == (2, 3, 1, 1)


def test_config_fields_and_defaults_equal_the_reference():
    """GenieConfig / DiffusionGenieConfig carry exactly the reference's fields and defaults (hma/config.py:8-117); checked
    against the live reference when it is present (authoring container), otherwise skipped."""
    import dataclasses

    from oracle import reference_loader

    if not reference_loader.available():
        pytest.skip("reference not present")
    reference_loader.load()
    from hma.config import DiffusionGenieConfig as RD
    from hma.config import GenieConfig as R

    from hma_b200 import DiffusionGenieConfig, GenieConfig

    for ref, ours in ((R, GenieConfig), (RD, DiffusionGenieConfig)):
        a = {f.name: f.default for f in dataclasses.fields(ref)}
        b = {f.name: f.default for f in dataclasses.fields(ours)}
        assert a == b
        assert [f.name for f in dataclasses.fields(ref)][:5] == [f.name for f in dataclasses.fields(ours)][:5]
    kw = dict(num_layers=2, num_heads=8, d_model=256, num_factored_vocabs=2)
    assert vars(R(**kw)) == vars(GenieConfig(**kw))
    assert vars(RD(patch_size=2, **kw)) == vars(DiffusionGenieConfig(patch_size=2, **kw))


def test_save_pretrained_from_pretrained_roundtrip(tmp_path):
    """The reference's checkpoint path (evaluate.py:141, generate.py:110, train_multi.py:310-321): config + weights through
    PyTorchModelHubMixin, for both model classes, and again after a TrainStep has re-homed the parameters into its arena."""
    import torch
    from hma_b200 import DiffusionGenieConfig, GenieConfig, STMaskGIT
    from hma_b200.mar import STMAR
    from hma_b200.train import TrainStep

    stats = [[[0.0] * 7, [1.0] * 7], [[0.0] * 14, [1.0] * 14]]
    for cls, cfg in ((STMaskGIT, GenieConfig(num_layers=2, num_heads=8, d_model=256, T=4, S=256, num_factored_vocabs=2,
                                             action_network="concat+modulate")),
                     (STMAR, DiffusionGenieConfig(num_layers=2, num_heads=8, d_model=256, T=4, S=256, num_factored_vocabs=2,
                                                  patch_size=2, action_network="concat+modulate"))):
        torch.manual_seed(0)
        m = cls(cfg)
        m.init_action_projectors(["a", "b"], [7, 14], stats, "concat+modulate")
        with torch.no_grad():
            for p in m.parameters():
                p.normal_(0.0, 0.1)
        m.save_pretrained(tmp_path / cls.__name__)
        m2 = cls.from_pretrained(tmp_path / cls.__name__)
        assert type(m2.config) is type(cfg) and m2.config.action_domains == ["a", "b"]
        sd1, sd2 = m.state_dict(), m2.state_dict()
        assert list(sd1) == list(sd2) and all(torch.equal(sd1[k], sd2[k]) for k in sd1)
    # parameters re-homed into the TrainStep arena are views of one storage: save_pretrained must still work
    m = STMaskGIT(GenieConfig(num_layers=2, num_heads=8, d_model=256, T=4, S=256, num_factored_vocabs=2,
                              action_network="concat+modulate"))
    m.init_action_projectors(["a", "b"], [7, 14], stats, "concat+modulate")
    before = {k: v.clone() for k, v in m.state_dict().items()}
    TrainStep(m)
    m.save_pretrained(tmp_path / "arena")
    m3 = STMaskGIT.from_pretrained(tmp_path / "arena")
    assert all(torch.equal(before[k], v) for k, v in m3.state_dict().items())
