"""MultiTaskBatchSampler + the device-side batch pipeline: the caller side of the training path.

`MultiTaskBatchSampler` mirrors the sampler the reference trainer builds (hma/train_multi.py:928-932, class in
external/data_sampler.py:177-303): every batch comes from ONE dataset (= one action domain, which is why the model only
reads `domain[0]`), chosen by temperature sampling over the dataset sizes; indices inside the dataset are drawn with
replacement from this rank's shard of a per-epoch permutation. The constructor, `generate_tasks_distribution`,
`set_epoch`, `__len__` and — given the same seed / epoch / rank — the exact index lists are the reference's (it consumes
the torch generator in the same order), so it can drive the reference's DataLoader unchanged.

`DeviceBatchPipeline` is the B200 data path: datasets resident in HBM (`RawTokenDataset.to_device`), a batch = one gather
kernel + the on-device collator, no dataloader workers and no host copies of token data.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch


class MultiTaskBatchSampler(torch.utils.data.Sampler):
    def __init__(self, dataset_sizes: List[int], batch_size: int, temperature: float, dataset_groups: Sequence = (),
                 num_replicas: Optional[int] = 1, rank: Optional[int] = 0, seed: int = 0, shuffle: bool = True,
                 shuffle_task: bool = True) -> None:
        if num_replicas is None or rank is None:
            import torch.distributed as dist

            if not dist.is_available():
                raise RuntimeError("Requires distributed package to be available")
            num_replicas = dist.get_world_size() if num_replicas is None else num_replicas
            rank = dist.get_rank() if rank is None else rank
        if rank >= num_replicas or rank < 0:
            raise ValueError("Invalid rank {}, rank should be in the interval [0, {}]".format(rank, num_replicas - 1))
        self.dataset_groups = list(dataset_groups)
        self.num_replicas, self.rank = num_replicas, rank
        self.shuffle, self.shuffle_task = shuffle, shuffle_task
        self.batch_size = batch_size
        self.dataset_sizes = list(dataset_sizes)
        self.rank_dataset_sizes = [s // num_replicas for s in self.dataset_sizes]  # the remainder is dropped
        self.total_sizes = [r * num_replicas for r in self.rank_dataset_sizes]
        self.dataset_offsets = torch.cumsum(torch.LongTensor([0] + self.dataset_sizes), 0)
        self.temperature = temperature
        self.seed, self.epoch = seed, 0
        self.num_batches_per_epoch = (int(np.sum(self.dataset_sizes)) + batch_size - 1) // batch_size // num_replicas

    def generate_tasks_distribution(self) -> torch.Tensor:
        """(size / total) ** (1 / temperature), normalised — per group first when dataset_groups = [(lo, hi), ...] is given."""
        def temper(sizes):
            w = np.array([(s / sum(sizes)) ** (1.0 / self.temperature) for s in sizes])
            return w / np.sum(w)

        if self.dataset_groups:
            parts = [temper(self.dataset_sizes[lo:hi]) / len(self.dataset_groups) for lo, hi in self.dataset_groups]
            weights = np.concatenate(parts)
        else:
            weights = temper(self.dataset_sizes)
        return torch.as_tensor(weights, dtype=torch.double)

    def iter_tasks(self) -> Iterator[Tuple[int, torch.Tensor]]:
        """(dataset index, i64 indices INSIDE that dataset) per batch of the current epoch."""
        gen = torch.Generator()
        gen.manual_seed(self.seed + self.epoch)
        shards = []
        for i, size in enumerate(self.dataset_sizes):
            order = torch.randperm(size, generator=gen) if self.shuffle else torch.arange(size)
            shards.append(order[self.rank: self.total_sizes[i]: self.num_replicas])
        if self.shuffle_task:  # ranks then draw different tasks; with False every rank trains the same task per step
            gen.manual_seed(self.seed + self.epoch + self.rank)
        tasks = torch.multinomial(self.generate_tasks_distribution(), self.num_batches_per_epoch, replacement=True, generator=gen)
        for task in tasks.tolist():
            pick = torch.randint(low=0, high=self.rank_dataset_sizes[task], size=(self.batch_size,), generator=gen)
            yield task, shards[task][pick]

    def __iter__(self):
        for task, local in self.iter_tasks():
            yield (self.dataset_offsets[task] + local).tolist()

    def __len__(self):
        return self.num_batches_per_epoch

    def set_epoch(self, epoch):
        self.epoch = epoch


class DeviceBatchPipeline:
    """datasets (already `.to_device()`-ed) + sampler + on-device collator -> batches for `TrainStep` / `model(**batch)`.

    Each item is the dict the reference's DataLoader + get_maskgit_collator yield (input_ids, labels, action_ids, domain,
    h, w), with every tensor already on the device."""

    def __init__(self, datasets: Sequence, config, batch_size: int, temperature: float = 3.0, num_replicas: int = 1, rank: int = 0,
                 seed: int = 0, collate: bool = True):
        self.datasets = list(datasets)
        self.config = config
        self.collate = collate
        self.sampler = MultiTaskBatchSampler([len(d) for d in self.datasets], batch_size, temperature, num_replicas=num_replicas,
                                             rank=rank, seed=seed)

    def __len__(self):
        return len(self.sampler)

    def set_epoch(self, epoch: int) -> None:
        self.sampler.set_epoch(epoch)

    def __iter__(self) -> Iterator[Dict[str, object]]:
        from .data import collate_from_draws, draw_on_device

        for task, local in self.sampler.iter_tasks():
            ds = self.datasets[task]
            batch = ds.gather(local)
            if self.collate:
                tokens = batch["input_ids"]
                h, w = batch["h"][0], batch["w"][0]
                draws = draw_on_device(self.config, tokens.shape[0], h, w, tokens.device)
                batch["input_ids"], batch["labels"] = collate_from_draws(tokens, draws, self.config, h, w)
            yield batch
