"""CPU-only: random tile-stream packs (warps, tile sizes, forced short rings -> far table, heavy thresholds, integer scores)
walked in numpy with the kernels' addressing (tests/tile_replay.py) and held to the float64 oracle -- every arc visited
once, ring slots consistent, beta / posteriors / Viterbi as the oracle's.  python tests/fuzz/fuzz_pack_cpu.py [seconds] [seed]"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from nfst_b200 import synth, tiles as T
from tests.test_pack_tiles import check_lattices
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t0 = time.time(); n = fails = 0
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
while time.time() - t0 < budget:
    arcs = int(np.exp(rng.uniform(np.log(400), np.log(9000))))
    levels = int(rng.choice([3, 5, 9, 17, 33]))
    B = int(rng.integers(1, 4))
    warps = int(rng.choice([0, 1, 2, 4, 8]))
    tile_arcs = int(rng.choice([64, 128, 384]))
    force = int(rng.choice([0, 0, 6, 12, 24]))
    tailmax = int(rng.choice([32, 32, 8, 16]))
    T.TILE_WARPS, T.TILE_ARCS, T.FORCE_RING_SLICES, T.TAILMAX = warps, tile_arcs, force, tailmax
    seed = int(rng.integers(0, 10**6))
    n += 1
    try:
        ab = synth.random_dag_batch(B, arcs, levels=levels, seed=seed)
        if rng.integers(0, 3) == 0:
            ab.scores = -torch.from_numpy(rng.integers(0, 3, size=ab.src.numel())).float()
        try:
            p, w = ab.pack(tiles=True)
        except ValueError as e:
            if "DP ring" in str(e) or "level is wider" in str(e):
                continue  # a forced tiny ring that cannot hold a level: refused at pack time, fine
            raise
        if not p.has_tiles:
            continue
        check_lattices(ab, p, w); globals().__setitem__("checked", globals().get("checked", 0) + 1)
    except Exception:
        fails += 1
        print(f"FAIL arcs={arcs} levels={levels} B={B} warps={warps} tile_arcs={tile_arcs} force={force} tailmax={tailmax} seed={seed}\n{traceback.format_exc()}", flush=True)
print(f"{n} cases ({globals().get('checked', 0)} replayed), {fails} failures, {time.time()-t0:.0f} s")
