"""hma_b200.sampler against the reference MultiTaskBatchSampler (tests/golden/sampler.pt, from oracle/make_sampler_golden.py):
identical index lists given the same seed / epoch / rank; and the device pipeline's host logic with stand-in datasets."""
from pathlib import Path

import pytest
import torch

from oracle.make_sampler_golden import CASES

GOLDEN = Path(__file__).parent / "golden" / "sampler.pt"


@pytest.mark.parametrize("case", list(CASES))
def test_index_lists_equal_the_reference(case):
    from hma_b200.sampler import MultiTaskBatchSampler

    ref = torch.load(GOLDEN, weights_only=False)[case]
    s = MultiTaskBatchSampler(**CASES[case])
    assert len(s) == ref["len"]
    torch.testing.assert_close(s.generate_tasks_distribution(), ref["weights"], rtol=1e-12, atol=0)
    for epoch in (0, 3):
        s.set_epoch(epoch)
        assert torch.equal(torch.tensor(list(iter(s))), ref[f"epoch{epoch}"])


def test_one_dataset_per_batch_and_rank_shards_are_disjoint():
    from hma_b200.sampler import MultiTaskBatchSampler

    sizes = [40, 100, 12]
    seen = []
    for rank in range(2):
        s = MultiTaskBatchSampler(sizes, batch_size=4, temperature=2.0, num_replicas=2, rank=rank, seed=1, shuffle_task=False)
        picked = [set() for _ in sizes]
        tasks = []
        for task, local in s.iter_tasks():
            assert local.numel() == 4 and int(local.max()) < sizes[task]
            picked[task].update(local.tolist())
            tasks.append(task)
        seen.append((tasks, picked))
    assert seen[0][0] == seen[1][0]  # shuffle_task=False: every rank trains the same dataset at every step
    for a, b in zip(seen[0][1], seen[1][1]):
        assert not (a & b)           # the two ranks draw from disjoint shards of the epoch's permutation
    with pytest.raises(ValueError):
        MultiTaskBatchSampler(sizes, 4, 1.0, num_replicas=2, rank=2)


class _FakeDataset:
    """Stands in for a RawTokenDataset that has been moved to the device."""

    def __init__(self, name, n):
        self.name, self.n = name, n
        self.calls = []

    def __len__(self):
        return self.n

    def gather(self, idx):
        self.calls.append(idx.clone())
        B = idx.numel()
        return {"input_ids": idx.view(B, 1).repeat(1, 4), "labels": None, "domain": [self.name] * B, "h": [2] * B, "w": [2] * B}


def test_device_pipeline_routes_batches_to_their_dataset():
    from hma_b200.sampler import DeviceBatchPipeline

    data = [_FakeDataset("a", 30), _FakeDataset("b", 200)]
    pipe = DeviceBatchPipeline(data, config=None, batch_size=5, temperature=3.0, seed=2, collate=False)
    batches = list(pipe)
    assert len(batches) == len(pipe) == (230 + 4) // 5
    for b in batches:
        assert len(set(b["domain"])) == 1 and b["input_ids"].shape == (5, 4)
    assert sum(len(d.calls) for d in data) == len(batches) and all(d.calls for d in data)
    assert all(int(c.max()) < d.n for d in data for c in d.calls)
    pipe.set_epoch(1)
    again = list(pipe)
    assert [b["domain"][0] for b in again] != [b["domain"][0] for b in batches] or \
        not all(torch.equal(x["input_ids"], y["input_ids"]) for x, y in zip(again, batches))

