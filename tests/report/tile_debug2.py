"""Debug helper: full forward-backward of single lattices of a batch through the tile-stream kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nfst_b200 as nb
from nfst_b200 import synth
from oracle import c_oracle

arcs, levels, B, which = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mode = sys.argv[5] if len(sys.argv) > 5 else "fb"
ab = synth.random_dag_batch(B, arcs, levels=levels, seed=7)
if which >= 0:
    m = ab.arc_lattice == which
    ab = synth.ArcBatch(torch.zeros(int(m.sum()), dtype=torch.int64), ab.src[m], ab.dst[m], ab.label[m], ab.scores[m],
                        ab.n_states[which:which + 1], ab.vocab)
p, sc = ab.to("cuda:0").pack()
deg = np.diff(p.out_ptr.cpu().numpy())
print("groups", [(g.tiles, g.block_threads, g.n, g.tile_ring, g.tile_far, g.tile_cap_arcs, g.tile_cap_bytes) for g in p.groups],
      "A", p.n_arcs, "A%4", p.n_arcs % 4, "max deg", deg.max(), "deg>32", int((deg > 32).sum()), flush=True)
ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(), ab.n_states.numpy())
o_logz, o_alpha, o_beta, o_post = c_oracle.forward_backward(ob)
def dbg():
    import ctypes
    out = (ctypes.c_int32 * 8)()
    nb._lib.load().nfst_tile_debug_read(out)
    print("tile debug record", list(out), flush=True)


if mode == "pullcond":
    logz, alpha, cond = nb.ops.lattice_pull(p, arc_scores=sc)
    dbg()
    torch.cuda.synchronize()
    print("pull+cond ok", float(logz[0]), o_logz[0], "cond sum/state ~", float(cond.sum()) / p.n_states)
else:
    logz, alpha, beta, post = nb.lattice_forward_backward(p, arc_scores=sc)
    torch.cuda.synchronize()
    ref = o_post[p.arc_origin.cpu().numpy()]
    got = post.cpu().numpy().astype(np.float64)
    print("fb ok logz", float(logz[0]), o_logz[0], "post max rel err", float(np.max(np.abs(got - ref) / (ref + 1e-7))))
