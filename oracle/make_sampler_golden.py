"""Pin hma_b200.sampler.MultiTaskBatchSampler against the REAL reference class (external/data_sampler.py:177-303) and write
tests/golden/sampler.pt (the reference's index lists for a few configurations).

    python -m oracle.make_sampler_golden

TEST INFRASTRUCTURE ONLY. external/data_sampler.py imports matplotlib / PIL at module level for its pie plot; they are
stubbed here (the sampler does not use them)."""
import contextlib
import io
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import reference_loader  # noqa: E402

CASES = {
    "one_rank": dict(dataset_sizes=[120, 37, 900, 64], batch_size=8, temperature=3.0),
    "rank1_of_4": dict(dataset_sizes=[120, 37, 900, 64], batch_size=4, temperature=3.0, num_replicas=4, rank=1, seed=7),
    "same_task_all_ranks": dict(dataset_sizes=[50, 500], batch_size=6, temperature=1.0, num_replicas=2, rank=1, shuffle_task=False),
    "no_shuffle_groups": dict(dataset_sizes=[30, 60, 90, 20], batch_size=5, temperature=4.0, shuffle=False,
                              dataset_groups=[(0, 2), (2, 4)]),
}


def reference_class():
    for name in ("matplotlib", "matplotlib.pyplot", "PIL", "PIL.Image"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["PIL"], "Image"):
        sys.modules["PIL"].Image = sys.modules["PIL.Image"]
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, str(reference_loader.REFERENCE_ROOT / "external"))
    import data_sampler
    return data_sampler.MultiTaskBatchSampler


def main():
    Ref = reference_class()
    out = {}
    for name, kw in CASES.items():
        with contextlib.redirect_stdout(io.StringIO()):
            s = Ref(**kw)
        rec = {"len": len(s), "weights": s.generate_tasks_distribution()}
        for epoch in (0, 3):
            s.set_epoch(epoch)
            rec[f"epoch{epoch}"] = torch.tensor(list(iter(s)))
        out[name] = rec
        print(name, rec["len"], rec["weights"].tolist())
    path = ROOT / "tests" / "golden" / "sampler.pt"
    torch.save(out, path)
    print("wrote", path)


if __name__ == "__main__":
    main()
