"""Load the REAL reference (read-only, /root/reference) in the authoring container.

TEST INFRASTRUCTURE ONLY. /root/reference does not exist on the GPU box, so nothing under
tests/ -m gpu, smoke() or bench.py imports this module; it is used by oracle/make_golden.py and by
the CPU-only test that pins oracle/stmaskgit_oracle.py against the reference when it is present.
"""
import os
import sys
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get("HMA_REFERENCE_ROOT", "/root/reference"))
_SHIMS = Path(__file__).resolve().parent / "ref_shims"


def available() -> bool:
    return (REFERENCE_ROOT / "hma" / "model" / "st_mask_git.py").exists()


def load():
    """Returns (STMaskGIT, GenieConfig) classes of the reference."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    os.environ["XFORMERS_DISABLED"] = "true"  # attention.py:7 reads it at import time
    for p in (str(_SHIMS), str(REFERENCE_ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        from hma.config import GenieConfig
        from hma.model.st_mask_git import STMaskGIT
    return STMaskGIT, GenieConfig
