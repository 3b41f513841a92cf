"""Synthetic dataset directory in the reference's on-disk format (datasets/encode_openx_dataset.py:340-388), shared by
oracle/make_rawtoken_golden.py and tests/test_dataset*.py. Deterministic given the seed."""
import json
from pathlib import Path

import numpy as np

# the stub DATA_FREQ_TABLE the golden generators install into the reference (oracle/make_rawtoken_golden.py): the synthetic
# dataset's name is not in the shipped table, so the tests pass the same stub through `freq_table`
FREQ = {"synthetic_robot": 6}

CASES = {
    "default": dict(window_size=4, use_actions=True),
    "overlaps": dict(window_size=3, use_actions=True, filter_overlaps=True),
    "max_traj": dict(window_size=4, use_actions=False, max_traj_num=3),
    "no_filter_fixed_stride": dict(window_size=5, stride=2, use_actions=True, filter_interrupts=False,
                                   compute_stride_from_freq_table=False),
}


FEATURE_CASES = {
    "default": dict(window_size=4, use_actions=True),
    "capped_overlaps": dict(window_size=3, use_actions=True, filter_overlaps=True, max_traj_num=40),
    "renamed_fixed_stride": dict(window_size=2, stride=3, use_actions=False, compute_stride_from_freq_table=False,
                                 domain="other_robot_noquant"),
}


def write(root: Path, seed: int = 0, num_images: int = 240, h: int = 16, w: int = 16, action_dim: int = 7, hz: int = 6,
          token_dtype: str = "uint32", latent_channels: int = 0) -> Path:
    """latent_channels > 0: the continuous-feature layout (token_dtype e.g. float16, [N, C, h, w], name "..._noquant")."""
    rng = np.random.default_rng(seed)
    root = Path(root)
    (root / "actions").mkdir(parents=True, exist_ok=True)
    if latent_channels:
        video = (rng.normal(size=(num_images, latent_channels, h, w)) * 4).astype(token_dtype)
    else:
        video = rng.integers(0, 2 ** 18 if token_dtype == "uint32" else 2 ** 16, size=(num_images, h, w)).astype(token_dtype)
    seg, lens = [], []
    while sum(lens) < num_images:
        lens.append(int(rng.integers(15, 45)))
    for i, n in enumerate(lens):
        seg += [i] * n
    seg = np.array(seg[:num_images], dtype=np.int32)
    actions = rng.normal(size=(num_images, action_dim)).astype(np.float32)
    for name, arr in (("video.bin", video), ("segment_ids.bin", seg), ("actions/actions.bin", actions)):
        fp = np.memmap(root / name, dtype=arr.dtype, mode="w+", shape=arr.shape)
        fp[:] = arr[:]
        fp.flush()
    with open(root / "metadata.json", "w") as f:
        json.dump({"token_dtype": token_dtype, "action_dim": action_dim, "s": 16, "h": h, "w": w, "vocab_size": 2 ** 18, "hz": hz,
                   "num_images": num_images, "name": "synthetic_robot" + ("_noquant" if latent_channels else ""),
                   "latent_channels": latent_channels or None, "quantized": not latent_channels}, f)
    return root
