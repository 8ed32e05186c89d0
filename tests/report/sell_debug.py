"""Locate mismatching states of the sliced-column pull pass against the float64 C oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nfst_b200 as nb  # noqa: E402
from nfst_b200 import synth  # noqa: E402
from oracle import c_oracle  # noqa: E402


def main():
    arcs, levels, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    ab = synth.random_dag_batch(B, arcs, levels=levels, seed=7)
    p, sc = ab.to("cuda:0").pack()
    print("groups", [(g.sell, g.block_threads, g.n, g.sell_window, g.sell_far) for g in p.groups])
    r = nb.ops.lattice_backward(p, sc, want_beta=True)
    torch.cuda.synchronize()
    ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(),
                        ab.n_states.numpy())
    o_logz, o_alpha, o_beta, o_post = c_oracle.forward_backward(ob)
    so = np.concatenate([[0], np.cumsum(ab.n_states.numpy())])
    state_off = p.state_off.cpu().numpy()
    lat = np.repeat(np.arange(p.n_lattices), np.diff(state_off))
    g2o = so[lat] + p.orig_state.cpu().numpy()
    beta = r["beta"].cpu().numpy().astype(np.float64)
    err = np.abs(beta - o_beta[g2o])
    bad = np.nonzero(err > 1e-4)[0]
    print("logz", r["logz_bwd"].cpu().numpy(), o_logz, "bad states", bad.size, "of", beta.size)
    if bad.size:
        level_ptr, level_off = p.level_ptr.cpu().numpy(), p.level_off.cpu().numpy()
        deg = np.diff(p.out_ptr.cpu().numpy())
        for s in list(bad[-12:]) + list(bad[:4]):
            b = lat[s]
            lp = level_ptr[level_off[b]:level_off[b + 1]]
            l = np.searchsorted(lp, s, side="right") - 1
            print(f"state {s} lattice {b} level {l} pos-in-level {s - lp[l]} (level size {lp[l + 1] - lp[l]}) deg {deg[s]} "
                  f"beta {beta[s]:.6f} oracle {o_beta[g2o[s]]:.6f}")
        lv_bad = {}
        for s in bad:
            b = lat[s]
            lp = level_ptr[level_off[b]:level_off[b + 1]]
            l = int(np.searchsorted(lp, s, side="right") - 1)
            lv_bad[(b, l)] = lv_bad.get((b, l), 0) + 1
        print("bad per (lattice, level):", sorted(lv_bad.items()))


if __name__ == "__main__":
    main()
