"""Pin oracle/collator_oracle.py against the REAL reference collator (hma/data.py:28-98) and write
tests/golden/collator.pt. Run in the authoring container (needs /root/reference):
    python -m oracle.make_collator_golden
Seeds torch's CPU generator and Python's `random`, calls the reference collate_fn, re-seeds, replays the draws with
oracle.collator_oracle.draw() and checks apply() reproduces the reference bit for bit; stores tokens, draws and outputs."""
import random
import sys
import types
from pathlib import Path

import torch

from oracle import collator_oracle as C
from oracle import reference_loader

CASES = [
    dict(name="mlm_corrupt", seed=3, T=4, B=3, cfg={}),
    dict(name="non_mlm", seed=11, T=8, B=2, cfg=dict(non_mlm_ratio=1.0, num_prompt_frames=3)),
    dict(name="no_corruption", seed=5, T=4, B=2, cfg=dict(dataloader_apply_corruption=False, non_mlm_ratio=0.0)),
    dict(name="one_vocab", seed=7, T=3, B=2, cfg=dict(num_factored_vocabs=1, image_vocab_size=64)),
]


def reference_collator():
    reference_loader.load()
    stub = types.ModuleType("datasets.encode_openx_dataset")  # hma/data.py:12 otherwise drags in tensorflow_datasets
    stub.DATA_FREQ_TABLE = {}
    pkg = types.ModuleType("datasets")
    pkg.encode_openx_dataset = stub
    sys.modules.setdefault("datasets", pkg)
    sys.modules["datasets.encode_openx_dataset"] = stub
    from hma.config import GenieConfig
    from hma.data import get_maskgit_collator
    return GenieConfig, get_maskgit_collator


def main():
    GenieConfig, get_maskgit_collator = reference_collator()
    out = {}
    for case in CASES:
        kw = dict(num_layers=1, num_heads=8, d_model=256, T=case["T"], S=256, num_factored_vocabs=2)
        kw.update(case["cfg"])
        cfg = GenieConfig(**kw)
        h = w = 16
        g = torch.Generator().manual_seed(case["seed"])
        tokens = torch.randint(0, cfg.image_vocab_size, (case["B"], cfg.T * h * w), generator=g)
        feats = [dict(input_ids=tokens[b].clone(), h=h, w=w, domain="d", action_ids=torch.zeros(cfg.T, 2)) for b in range(case["B"])]
        torch.manual_seed(case["seed"]); random.seed(case["seed"])
        ref = get_maskgit_collator(cfg)(feats)
        torch.manual_seed(case["seed"]); random.seed(case["seed"])
        d = C.draw(cfg, case["B"], h, w)
        ids, labels = C.apply(tokens, d, cfg, h, w)
        assert torch.equal(ids, ref["input_ids"]), case["name"]
        assert torch.equal(labels, ref["labels"]), case["name"]
        out[case["name"]] = dict(cfg=kw, seed=case["seed"], tokens=tokens, draws=d, input_ids=ref["input_ids"], labels=ref["labels"])
        print(case["name"], "ok; masked fraction", (ref["input_ids"] == cfg.image_vocab_size).float().mean().item(),
              "first_masked_frame", d["first_masked_frame"])
    path = Path(__file__).resolve().parent.parent / "tests" / "golden" / "collator.pt"
    torch.save(out, path)
    print("wrote", path)


if __name__ == "__main__":
    main()
