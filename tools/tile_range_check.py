"""Range-checked run of the tile-stream kernels (compute-sanitizer is closed on this pool): build the library with
-DNFST_TILE_DEBUG (every data-dependent ring slot, arc id, state id and tile header field is checked inside the
kernels; the first violation is recorded), run the shapes the tile tests use plus the config-4 sweep points through
both semirings, both score modes, float64 state and the 64-bit flow, and read the violation record back.

    python tools/tile_range_check.py build      # here (nvcc cross-compiles): nfst_b200/lib/libnfst_b200_dbg.so
    NFST_LIB=nfst_b200/lib/libnfst_b200_dbg.so python tools/tile_range_check.py      # on the GPU box
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DBG = os.path.join(ROOT, "nfst_b200", "lib", "libnfst_b200_dbg.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    from nfst_b200 import build

    print(build.build_library(out=DBG, defines=["NFST_TILE_DEBUG"]))
    sys.exit(0)

import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import _lib, synth  # noqa: E402
from nfst_b200 import tiles as T  # noqa: E402

assert os.environ.get("NFST_LIB", "").endswith("_dbg.so"), "run with NFST_LIB=<the -DNFST_TILE_DEBUG build>"
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.nfst_tile_debug_read.argtypes = [C.c_void_p]
rec = (C.c_int32 * 8)()
bad = 0


def check(name):
    global bad
    torch.cuda.synchronize()
    assert lib.nfst_tile_debug_read(C.cast(rec, C.c_void_p)) == 0
    v = list(rec)
    ok = v[0] == 0
    bad += not ok
    print(f"{name:70s} {'no violation' if ok else 'VIOLATION ' + str(v)}", flush=True)


def run(name, ab, **env):
    old = {k: getattr(T, k) for k in env}
    for k, v in env.items():
        setattr(T, k, v)
    try:
        p, sc = ab.to(dev).pack()
    finally:
        for k, v in old.items():
            setattr(T, k, v)
    if not p.has_tiles:
        print(f"{name:70s} not a tile-stream batch (the packer left it to the other kernels): skipped", flush=True)
        return
    theta = torch.randn(p.vocab, device=dev)
    nb.lattice_forward_backward(p, arc_scores=sc)
    nb.lattice_viterbi(p, arc_scores=sc)
    nb.lattice_forward_backward(p, theta=theta, want_dtheta=True)
    nb.lattice_forward_backward(p, arc_scores=sc, theta=theta, want_dtheta=True)
    nb.lattice_viterbi(p, theta=theta)
    try:
        nb.lattice_forward_backward(p, arc_scores=sc, state_dtype=torch.float64)
    except RuntimeError as e:  # a ring sized for float32 state (at most 96 levels) may not fit with 8-byte slots
        assert "shared memory" in str(e) and p.max_levels <= 96, e
        name += " (float64 state forced: ring does not fit, refused)"
    old_bits = nb.ops.FLOW_BITS
    nb.ops.FLOW_BITS = 64
    try:
        nb.lattice_forward_backward(p, arc_scores=sc)
    except RuntimeError as e:  # the 64-bit accumulator doubles the ring: refused next to a ring that fills the SM
        assert "shared memory" in str(e), e
        name += " (64-bit flow: ring does not fit, refused)"
    finally:
        nb.ops.FLOW_BITS = old_bits
    check(f"{name} [{', '.join(str(g.block_threads // 32) + ' warps' for g in p.groups)}; ring {max(g.tile_ring for g in p.groups)}, far {max(g.tile_far for g in p.groups)}]")


run("3k arcs x 5, 8 levels", synth.random_dag_batch(5, 3_000, levels=8, seed=7))
run("10k arcs x 64", synth.random_dag_batch(64, 10_000, seed=3))
run("30k arcs x 64", synth.random_dag_batch(64, 30_000, seed=3))
run("100k arcs x 64", synth.random_dag_batch(64, 100_000, seed=3))
run("100k arcs x 8, 1 warp", synth.random_dag_batch(8, 100_000, seed=4), TILE_WARPS=1)
run("300k arcs x 16", synth.random_dag_batch(16, 300_000, seed=3))
run("1M arcs x 4", synth.random_dag_batch(4, 1_000_000, seed=3))
run("400k arcs x 2, 16 warps", synth.random_dag_batch(2, 400_000, seed=7), TILE_WARPS=16)
run("12k arcs x 3, ring forced to 20 slices (far table)", synth.random_dag_batch(3, 12_000, levels=24, seed=5), FORCE_RING_SLICES=20, TILE_WARPS=2)
run("8k arcs x 3, 6 levels, 64-arc tiles (heavy states in pieces)", synth.random_dag_batch(3, 8_000, levels=6, seed=9), TILE_ARCS=64)
run("60k arcs x 3, 400 levels", synth.random_dag_batch(3, 60_000, levels=400, seed=9))
run("1M arcs x 2, 128 levels (float64 state by default: 8-byte ring slots)", synth.random_dag_batch(2, 1_000_000, levels=128, seed=2))
run("2M arcs x 2, 128 levels", synth.random_dag_batch(2, 2_000_000, levels=128, seed=2))
print("violations:", bad)
sys.exit(1 if bad else 0)
