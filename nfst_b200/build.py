"""In-tree build of the sm_100a library (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = [os.path.join(PKG, "csrc", "nfst_kernels.cu"), os.path.join(PKG, "csrc", "nfst_sell.cu"),
       os.path.join(PKG, "csrc", "nfst_walk.cu")]
HDR = [os.path.join(ROOT, "include", "nfst_b200.h")]
LIB = os.path.join(PKG, "lib", "libnfst_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SRC + HDR)


def build_library(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """Compile nfst_b200/csrc/*.cu into nfst_b200/lib/libnfst_b200.so for sm_100a.
    `out` / `defines` build a variant elsewhere (e.g. -DNFST_TIMING for tools/phase_timing.py)."""
    if out is None:
        if not force and not is_stale():
            return LIB
        out = LIB
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", os.path.join(ROOT, "include"), "-o", out, *SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
