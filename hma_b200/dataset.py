"""RawTokenDataset: the reference's memmap token dataset (hma/data.py:159-294) with a device-resident fast path.

On-disk format (written by datasets/encode_openx_dataset.py:340-388): `metadata.json` (num_images, h, w, token_dtype,
action_dim, hz, name, ...), `video.bin` (token_dtype [num_images, h, w]), `segment_ids.bin` (int32 [num_images]),
`actions/*.bin` (float32 [num_images, action_dim] each, concatenated along the last axis).

Same constructor arguments, `valid_start_inds`, `__len__`, `__getitem__` dict (CPU tensors) and `action_stat` as the
reference, so it can stand behind a torch DataLoader unchanged. The B200 path is `to_device()` + `gather(indices)`: the
token and action tables are uploaded once (a 1 M-frame dataset is 1 GB of the 180 GB), and a batch is one index gather
on the device (csrc/dataset.cu) that feeds the on-device collator (hma_b200/data.py) — no per-sample host work, no
dataloader workers.

Stride: the reference looks the dataset name up in DATA_FREQ_TABLE (datasets/encode_openx_dataset.py:51-108) and uses
max(hz // natural_hz, 1); the same table ships here as data (data_freq_table.json) and is looked up the same way —
by `name`, default 1 — so a `name` that differs from the directory's own name resolves exactly as in the reference.
Pass `freq_table` to override it.
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops


SVD_SCALE = 0.18215  # hma/data.py:16: latent scaling of the continuous (SVD-VAE) features


def _load_freq_table() -> Dict[str, int]:
    """DATA_FREQ_TABLE of the reference (datasets/encode_openx_dataset.py:51-108; read at hma/data.py:207): control
    frequency per dataset name. Shipped as data (hma_b200/data_freq_table.json, extracted by oracle/make_freq_table.py)."""
    with open(Path(__file__).with_name("data_freq_table.json")) as f:
        return {k: int(v) for k, v in json.load(f).items()}


DATA_FREQ_TABLE: Dict[str, int] = _load_freq_table()


def normalize_actions(actions: np.ndarray):
    """data.py:18-24: statistics only; the normalisation itself happens inside the network (ActionStat)."""
    return actions, [np.mean(actions, axis=0).tolist(), np.std(actions, axis=0).tolist()]


class RawTokenDataset(torch.utils.data.Dataset):
    def __init__(self, data_dir, window_size, stride=1, filter_interrupts=True, filter_overlaps=False, use_actions=False, name="",
                 max_traj_num=1000000, compute_stride_from_freq_table=True, natural_hz=2, drop_action_ratio=0.0,
                 freq_table: Optional[Dict[str, int]] = None):
        data_dir = Path(data_dir)
        with open(data_dir / "metadata.json") as f:
            self.metadata = json.load(f)
        n_img = self.metadata["num_images"]
        shape = (n_img, self.metadata["h"], self.metadata["w"])
        token_dtype = np.dtype(self.metadata.get("token_dtype", "uint32"))
        self.data = np.memmap(data_dir / "video.bin", dtype=token_dtype, mode="r", shape=shape)
        self.window_size, self.stride = window_size, stride
        self.name = name if len(name) else self.metadata["name"]
        if compute_stride_from_freq_table:
            # data.py:207: DATA_FREQ_TABLE.get(self.name, 1) — the table, not the metadata's "hz" field
            table = DATA_FREQ_TABLE if freq_table is None else freq_table
            self.stride = max(table.get(self.name, 1) // natural_hz, 1)
        self.n_action = self.metadata.get("action_dim", 1) * self.stride
        self.drop_action_ratio = drop_action_ratio
        if use_actions:
            parts = [np.memmap(fn, dtype=np.float32, mode="r").reshape(len(self.data), -1)
                     for fn in sorted((data_dir / "actions").iterdir())]
            self.actions, self.action_stat = normalize_actions(np.concatenate(parts, axis=-1))
        seg_path = data_dir / "segment_ids.bin"
        if os.path.isfile(seg_path):
            self.segment_ids = np.memmap(seg_path, dtype=np.int32, mode="r", shape=(n_img,))
        else:
            self.segment_ids = None
            if filter_interrupts:
                raise NotImplementedError("Cannot filter interrupted sequences without segment ids.")
        self.video_len = (self.window_size - 1) * self.stride
        self.valid_start_inds = self._valid_starts(filter_interrupts, max_traj_num)
        if filter_overlaps:
            self.valid_start_inds = self._drop_overlaps(self.valid_start_inds)
        self.num_videos = len(np.unique(self.valid_start_inds))
        self._dev: Optional[dict] = None

    # ------------------------------------------------------------------ data.py:235-244, vectorised
    def _valid_starts(self, filter_interrupts: bool, max_traj_num: int) -> List[int]:
        n = max(len(self.data) - self.video_len - self.stride, 0)
        seg = None if self.segment_ids is None else np.asarray(self.segment_ids)
        if seg is not None and n:
            # the reference loop stops AFTER processing the first start whose segment id reaches max_traj_num
            over = np.nonzero(seg[:n] >= max_traj_num)[0]
            if len(over):
                n = int(over[0]) + 1
        starts = np.arange(n)
        if filter_interrupts and n:
            starts = starts[seg[:n] == seg[self.video_len: self.video_len + n]]
        return starts.tolist()

    def _drop_overlaps(self, starts: Sequence[int]) -> List[int]:
        """data.py:246-260: greedy in order; a start is dropped if an already kept start lies exactly i*stride before it
        (i < window_size), looking only at the last window_size*stride kept starts, as the reference does."""
        kept: List[int] = []
        for s in starts:
            clash = {s - i * self.stride for i in range(1, self.window_size)}
            if not any(k in clash for k in kept[-self.window_size * self.stride:]):
                kept.append(s)
        return kept

    def __len__(self):
        return len(self.valid_start_inds)

    # ------------------------------------------------------------------ data.py:265-294 (host path, CPU tensors)
    def __getitem__(self, idx):
        start = self.valid_start_inds[idx]
        x = torch.from_numpy(self.data[start: start + self.video_len + 1: self.stride].astype(np.int64)).flatten()
        out = {"input_ids": x, "labels": x, "attention_mask": torch.ones_like(x), "h": self.metadata["h"], "w": self.metadata["w"]}
        if hasattr(self, "actions") and np.random.uniform() > self.drop_action_ratio:
            a = self.actions[start: start + self.video_len + self.stride].reshape(self.window_size, -1)
            out["action_ids"] = torch.from_numpy(a.astype(np.float32))
        out["domain"] = self.name
        return out

    # ------------------------------------------------------------------ device path
    def to_device(self, device="cuda") -> "RawTokenDataset":
        """Upload the token table (in its stored dtype), the action table and the start indices to HBM."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("RawTokenDataset.to_device needs a CUDA device (the host path is __getitem__)")
        if self.data.dtype.itemsize not in (2, 4):
            raise NotImplementedError(f"token dtype {self.data.dtype} (uint16 / uint32 tables are supported on the device)")
        # one host copy of the read-only memmap (torch cannot wrap non-writable arrays), reinterpreted as a signed type of
        # the same width for torch; the kernel reads the bits as unsigned
        view = np.array(self.data).view(np.int16 if self.data.dtype.itemsize == 2 else np.int32)
        d = {"video": torch.from_numpy(view).to(dev), "starts": torch.tensor(self.valid_start_inds, dtype=torch.int64, device=dev)}
        if hasattr(self, "actions"):
            d["actions"] = torch.from_numpy(np.array(self.actions, dtype=np.float32)).to(dev)
        self._dev = d
        return self

    def gather(self, indices) -> Dict[str, object]:
        """The batch `[self[i] for i in indices]` stacked, on the device: input_ids / labels i64 [B, window*h*w] (one tensor,
        as in the reference where labels is input_ids), action_ids f32 [B, window, stride*action_dim], domain, h, w lists.
        Every sample carries its actions (drop_action_ratio is a host-path, per-sample decision)."""
        if self._dev is None:
            raise RuntimeError("call to_device() first (there is no CPU fallback for gather; use __getitem__ on the host)")
        d = self._dev
        dev = d["video"].device
        idx = torch.as_tensor(indices, dtype=torch.int64, device=dev)
        starts = d["starts"][idx].contiguous()
        B = starts.numel()
        h, w = self.metadata["h"], self.metadata["w"]
        tokens = torch.empty(B, self.window_size * h * w, device=dev, dtype=torch.int64)
        _lib.call("hma_gather_token_windows", d["video"].data_ptr(), self.data.dtype.itemsize, len(self.data), starts.data_ptr(), B,
                  self.window_size, self.stride, h * w, tokens.data_ptr(), ops._s())
        out: Dict[str, object] = {"input_ids": tokens, "labels": tokens}
        if "actions" in d:
            adim = d["actions"].shape[1]
            rows = self.video_len + self.stride
            act = torch.empty(B, self.window_size, rows * adim // self.window_size, device=dev, dtype=torch.float32)
            _lib.call("hma_gather_rows_f32", d["actions"].data_ptr(), d["actions"].shape[0], adim, starts.data_ptr(), B, rows,
                      act.data_ptr(), ops._s())
            out["action_ids"] = act
        out["domain"] = [self.name] * B
        out["h"] = [h] * B
        out["w"] = [w] * B
        return out


class RawFeatureDataset(torch.utils.data.Dataset):
    """The continuous-latent sibling (hma/data.py:297-435), read by STMAR training: `video.bin` holds
    token_dtype (float16 by default) [num_images, latent_channels, h, w]; an item is the window scaled by SVD_SCALE and laid
    out "(t h w) c". Host path only (CPU tensors, for a torch DataLoader); same constructor arguments and quirks as the
    reference: `max_traj_num` caps the NUMBER of valid windows here (data.py:383-384), not the segment id, and "_noquant" is
    stripped from the domain name."""

    def __init__(self, data_dir, window_size, stride=1, filter_interrupts=True, filter_overlaps=False, use_actions=False,
                 max_traj_num=1000000, compute_stride_from_freq_table=True, natural_hz=2, datio_noise_ratio=0.0,
                 use_raw_image_as_latent=False, domain=None, freq_table: Optional[Dict[str, int]] = None):
        data_dir = Path(data_dir)
        with open(data_dir / "metadata.json") as f:
            self.metadata = json.load(f)
        n_img = self.metadata["num_images"]
        shape = (n_img, self.metadata.get("latent_channels", 4), self.metadata["h"], self.metadata["w"])
        self.data = np.memmap(data_dir / "video.bin", mode="r", shape=shape, dtype=np.dtype(self.metadata.get("token_dtype", "float16")))
        self.window_size, self.stride = window_size, stride
        self.datio_noise_ratio = datio_noise_ratio
        self.name = (domain if domain is not None else self.metadata["name"]).replace("_noquant", "")
        if compute_stride_from_freq_table:
            table = DATA_FREQ_TABLE if freq_table is None else freq_table  # data.py:350
            self.stride = max(table.get(self.name, 1) // natural_hz, 1)
        self.n_action = self.metadata.get("action_dim", 1) * self.stride
        if use_actions:
            parts = [np.memmap(fn, dtype=np.float32, mode="r").reshape(len(self.data), -1)
                     for fn in sorted((data_dir / "actions").iterdir())]
            self.actions, self.action_stat = normalize_actions(np.concatenate(parts, axis=-1))
        seg_path = data_dir / "segment_ids.bin"
        if os.path.isfile(seg_path):
            self.segment_ids = np.memmap(seg_path, dtype=np.int32, mode="r", shape=(n_img,))
        else:
            self.segment_ids = None
            if filter_interrupts:
                raise NotImplementedError("Cannot filter interrupted sequences without segment ids.")
        self.video_len = (self.window_size - 1) * self.stride
        n = max(len(self.data) - self.video_len - self.stride, 0)
        starts = np.arange(n)
        if filter_interrupts and n:
            seg = np.asarray(self.segment_ids)
            starts = starts[seg[:n] == seg[self.video_len: self.video_len + n]]
        # the reference appends, then stops once it holds max_traj_num windows (max_traj_num <= 0 still keeps the first)
        self.valid_start_inds = starts[: max(int(max_traj_num), 1)].tolist() if len(starts) else []
        if filter_overlaps:
            self.valid_start_inds = RawTokenDataset._drop_overlaps(self, self.valid_start_inds)

    def __len__(self):
        return len(self.valid_start_inds)

    def __getitem__(self, idx):
        start = self.valid_start_inds[idx]
        x = torch.from_numpy(self.data[start: start + self.video_len + 1: self.stride].astype(np.float32)) * SVD_SCALE
        x = x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])  # "t c h w -> (t h w) c"
        out = {"input_ids": x, "labels": x, "attention_mask": torch.ones_like(x), "h": self.metadata["h"], "w": self.metadata["w"],
               "c": self.metadata["latent_channels"]}
        if hasattr(self, "actions"):
            a = self.actions[start: start + self.video_len + self.stride].reshape(self.window_size, -1)
            out["action_ids"] = torch.from_numpy(a.astype(np.float32))
        out["domain"] = self.name.replace("_noquant", "")
        return out
