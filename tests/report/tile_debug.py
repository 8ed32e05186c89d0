"""Debug helper for the tile-stream kernels: per-level error of beta against the float64 oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nfst_b200 as nb
from nfst_b200 import synth
from oracle import c_oracle

arcs, levels, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ab = synth.random_dag_batch(B, arcs, levels=levels, seed=7)
p, sc = ab.to("cuda:0").pack()
print("groups", [(g.tiles, g.block_threads, g.n, g.tile_ring, g.tile_far, g.tile_cap_arcs, g.tile_cap_bytes) for g in p.groups])
r = nb.lattice_backward(p, arc_scores=sc, want_beta=True)
torch.cuda.synchronize()
ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(), ab.n_states.numpy())
o_logz, o_alpha, o_beta, o_post = c_oracle.forward_backward(ob)
so = np.concatenate([[0], np.cumsum(ab.n_states.numpy())])
state_off = p.state_off.cpu().numpy()
lat = np.repeat(np.arange(p.n_lattices), np.diff(state_off))
g2o = so[lat] + p.orig_state.cpu().numpy()
beta = r["beta"].cpu().numpy().astype(np.float64)
err = np.abs(beta - o_beta[g2o])
level_ptr, level_off = p.level_ptr.cpu().numpy(), p.level_off.cpu().numpy()
deg = np.diff(p.out_ptr.cpu().numpy())
for b in range(min(B, 2)):
    lp = level_ptr[level_off[b]:level_off[b + 1]]
    print("lattice", b, "logz", float(r["logz_bwd"][b]), "oracle", o_logz[b])
    for l in range(len(lp) - 1):
        e = err[lp[l]:lp[l + 1]]
        bad = np.nonzero(e > 1e-4)[0]
        print(f"  level {l}: states {lp[l + 1] - lp[l]} max err {e.max():.3e} bad {len(bad)} first bad {bad[:8]} deg of bad {deg[lp[l]:lp[l + 1]][bad[:8]]}")
