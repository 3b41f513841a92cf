"""Pin hma_b200.dataset.RawTokenDataset against the REAL reference class (hma/data.py:159-294) on a synthetic dataset
directory and write tests/golden/rawtoken.pt (reference outputs only; the directory is regenerated from its seed).

    python -m oracle.make_rawtoken_golden

TEST INFRASTRUCTURE ONLY. The reference resolves the stride through DATA_FREQ_TABLE[name]; the stub table installed here
holds the synthetic dataset's hz, which is what the writer also stores in metadata.json."""
import contextlib
import io
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import reference_loader  # noqa: E402
from tests import _rawdata  # noqa: E402


def reference_dataset_class(freq):
    reference_loader.load()
    stub = types.ModuleType("datasets.encode_openx_dataset")
    stub.DATA_FREQ_TABLE = freq
    pkg = types.ModuleType("datasets")
    pkg.encode_openx_dataset = stub
    sys.modules["datasets"] = pkg
    sys.modules["datasets.encode_openx_dataset"] = stub
    sys.modules.pop("hma.data", None)
    import hma.data as D
    return D.RawTokenDataset, D


def main():
    Ref, D = reference_dataset_class({"synthetic_robot": 6})
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        root = _rawdata.write(Path(tmp) / "ds", seed=0)
        for name, kw in _rawdata.CASES.items():
            with contextlib.redirect_stdout(io.StringIO()):
                ds = Ref(root, **kw)
            idx = sorted(set([0, 1, len(ds) // 2, len(ds) - 1]))
            np.random.seed(0)
            items = [ds[i] for i in idx]
            out[name] = {"valid_start_inds": list(map(int, ds.valid_start_inds)), "len": len(ds), "stride": ds.stride,
                         "n_action": ds.n_action, "num_videos": ds.num_videos, "idx": idx,
                         "input_ids": torch.stack([it["input_ids"] for it in items]),
                         "action_ids": torch.stack([it["action_ids"] for it in items]) if "action_ids" in items[0] else None,
                         "action_stat": getattr(ds, "action_stat", None), "domain": items[0]["domain"]}
            print(name, len(ds), ds.stride, ds.n_action)
    # continuous-latent dataset + its collator (hma/data.py:297-435, 100-157)
    import random

    from hma.config import DiffusionGenieConfig
    with tempfile.TemporaryDirectory() as tmp:
        root = _rawdata.write(Path(tmp) / "feat", seed=3, token_dtype="float16", latent_channels=4, h=8, w=8)
        for name, kw in _rawdata.FEATURE_CASES.items():
            with contextlib.redirect_stdout(io.StringIO()):
                ds = D.RawFeatureDataset(root, **kw)
            idx = sorted(set([0, 1, len(ds) // 2, len(ds) - 1]))
            items = [ds[i] for i in idx]
            rec = {"valid_start_inds": list(map(int, ds.valid_start_inds)), "len": len(ds), "stride": ds.stride, "n_action": ds.n_action,
                   "idx": idx, "input_ids": torch.stack([it["input_ids"] for it in items]),
                   "action_ids": torch.stack([it["action_ids"] for it in items]) if "action_ids" in items[0] else None,
                   "domain": items[0]["domain"], "c": items[0]["c"]}
            for tag, cfg_kw, seed in (("mlm", dict(non_mlm_ratio=0.0), 5), ("non_mlm", dict(non_mlm_ratio=1.0, num_prompt_frames=1), 9)):
                cfg = DiffusionGenieConfig(num_layers=1, num_heads=8, d_model=256, T=kw["window_size"], **cfg_kw)
                torch.manual_seed(seed); random.seed(seed)
                with contextlib.redirect_stdout(io.StringIO()):
                    b = D.get_maskgit_collator_feature(cfg)(items)
                rec[f"collate_{tag}"] = {"cfg": dict(T=kw["window_size"], **cfg_kw), "seed": seed,
                                         "masked_tokens_indicator": b["masked_tokens_indicator"].clone(),
                                         "input_ids_shape": tuple(b["input_ids"].shape)}
            out["feature_" + name] = rec
            print("feature", name, len(ds), ds.stride, ds.n_action, rec["collate_mlm"]["masked_tokens_indicator"].float().mean().item())
    path = ROOT / "tests" / "golden" / "rawtoken.pt"
    torch.save(out, path)
    print("wrote", path, path.stat().st_size)


if __name__ == "__main__":
    main()
