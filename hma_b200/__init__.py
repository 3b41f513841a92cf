"""hma_b200: B200-native (sm_100a) implementation of HMA's ST-MaskGIT hot path.

    from hma_b200 import STMaskGIT, GenieConfig

The CUDA kernels live in hma_b200/csrc and are reached through the C ABI in include/hma_b200.h
(libhma_b200.so, built by `python -m hma_b200.build`). Nothing here falls back to the CPU.
"""
from .config import GenieConfig  # noqa: F401


def __getattr__(name):
    if name == "STMaskGIT":
        from .model import STMaskGIT

        return STMaskGIT
    raise AttributeError(name)
