"""hma_b200: B200-native (sm_100a) implementation of HMA's ST-MaskGIT hot path.

    from hma_b200 import STMaskGIT, GenieConfig                  # discrete tokens (hma/model/st_mask_git.py)
    from hma_b200 import STMAR, DiffusionGenieConfig             # continuous tokens + diffusion head (hma/model/st_mar.py)
    from hma_b200 import RawTokenDataset, get_maskgit_collator   # hma/data.py

The CUDA kernels live in hma_b200/csrc and are reached through the C ABI in include/hma_b200.h
(libhma_b200.so, built by `python -m hma_b200.build`). Nothing here falls back to the CPU.
"""
from .config import GenieConfig  # noqa: F401


_LAZY = {"STMaskGIT": "model", "STMAR": "mar", "DiffusionGenieConfig": "mar", "MarTrainStep": "mar", "TrainStep": "train",
         "RawTokenDataset": "dataset", "RawFeatureDataset": "dataset", "get_maskgit_collator_feature": "data", "get_maskgit_collator": "data", "MultiTaskBatchSampler": "sampler",
         "DeviceBatchPipeline": "sampler"}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(f".{_LAZY[name]}", __name__), name)
    raise AttributeError(name)
