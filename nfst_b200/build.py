"""In-tree build of the sm_100a library (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = [os.path.join(PKG, "csrc", "nfst_kernels.cu"), os.path.join(PKG, "csrc", "nfst_sell.cu"),
       os.path.join(PKG, "csrc", "nfst_tiles.cu"), os.path.join(PKG, "csrc", "nfst_walk.cu"),
       os.path.join(PKG, "csrc", "nfst_pack.cu")]
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


def _compile_one(args):
    src, obj, flags = args
    res = subprocess.run([_nvcc(), *flags, "-c", "-o", obj, src], capture_output=True, text=True)
    return src, res


def build_library(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """Compile nfst_b200/csrc/*.cu into nfst_b200/lib/libnfst_b200.so for sm_100a: one object per source, the
    sources in parallel, only those that changed.  `out` / `defines` build a variant elsewhere (e.g. -DNFST_TIMING
    for tools/phase_timing.py)."""
    from concurrent.futures import ThreadPoolExecutor

    variant = out is not None or bool(defines)
    if out is None:
        if not force and not is_stale():
            return LIB
        out = LIB
    os.makedirs(os.path.dirname(out), exist_ok=True)
    obj_dir = os.path.join(os.path.dirname(out), "obj" + ("_" + os.path.basename(out) if variant else ""))
    os.makedirs(obj_dir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + [f"-D{d}" for d in defines] + ["-I", os.path.join(ROOT, "include")]
    if verbose:
        flags.insert(0, "-Xptxas=-v")
    hdr_t = max(os.path.getmtime(h) for h in HDR)
    jobs, objs = [], []
    for src in SRC:
        obj = os.path.join(obj_dir, os.path.splitext(os.path.basename(src))[0] + ".o")
        objs.append(obj)
        fresh = os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t)
        if force or variant or verbose or not fresh:
            jobs.append((src, obj, flags))
    with ThreadPoolExecutor(max_workers=max(len(jobs), 1)) as pool:
        for src, res in pool.map(_compile_one, jobs):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
    res = subprocess.run([_nvcc(), "-shared", "-o", out, *objs], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
