"""Build libhma_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hma_b200.build [--force] [--verbose]

The shared library is the C ABI declared in include/hma_b200.h. It links cudart statically and
resolves the one driver symbol it needs (cuTensorMapEncodeTiled) at run time, so it has no
link-time dependency on libcuda.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "libhma_b200.so"
ROOT = HERE.parent

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(ROOT / "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    flags = list(NVCC_FLAGS)
    if os.environ.get("HMA_B200_TIMELINE") == "1":  # development: clock64() event tables (tools/timeline.py)
        flags += ["-DHMA_TIMELINE"]
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "hma_b200.h"]
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *flags, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))

    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
               "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


PROBE_SRC = ROOT / "tests" / "csrc" / "umma_probe.cu"
PROBE_LIB = ROOT / "tests" / "libhma_b200_probe.so"


def build_probe(force: bool = False) -> Path:
    """Test-only tcgen05 descriptor probe (tests/test_umma_probe_gpu.py): its own shared library, linked against the
    product's host-side helpers (TMA descriptor encoding), so the probe is not an entry point of libhma_b200.so."""
    build(force=False)
    host_obj = OBJ / "host.o"
    headers = sorted(CSRC.glob("*.cuh"))
    if force or _stale(PROBE_LIB, [PROBE_SRC, host_obj] + headers):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", str(PROBE_SRC), str(host_obj), "-o", str(PROBE_LIB), "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"probe build failed:\n{r.stdout}\n{r.stderr}")
    return PROBE_LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
    print(build_probe(force="--force" in sys.argv))
