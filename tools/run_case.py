"""Run a few forward + fused-backward steps of one synthetic workload (profiling driver).

    python tools/run_case.py --workload dag --arcs 100000 --batch 296 --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import nfst_b200 as nb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="dag")
ap.add_argument("--arcs", type=int, default=100_000)
ap.add_argument("--batch", type=int, default=296)
ap.add_argument("--levels", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--viterbi", action="store_true")
ap.add_argument("--theta", action="store_true", help="scores = theta[label] (WFSTScorer mode) instead of per-arc")
a = ap.parse_args()
dev = torch.device("cuda", 0)
packed, scores = bench.build_packed(a, dev)
kw = dict(arc_scores=scores)
if a.theta:
    kw = dict(theta=-torch.rand(packed.vocab, device=dev))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for i in range(a.steps):
    ev[0].record()
    alpha, logz = nb.lattice_forward(packed, **kw)
    ev[1].record()
    r = nb.lattice_backward(packed, alpha=alpha, logz=logz, want_beta=True, want_post=True, want_dtheta=a.theta, **kw,
                            want_viterbi=a.viterbi)
    ev[2].record()
    torch.cuda.synchronize()
    f, b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(f"step {i}: A={packed.n_arcs} S={packed.n_states} fwd {f:.3f} ms ({packed.n_arcs / f / 1e6:.1f} Garcs/s) "
          f"bwd {b:.3f} ms ({packed.n_arcs / b / 1e6:.1f} Garcs/s) groups="
          f"{[(g.n, g.block_threads, g.window_states()) for g in packed.groups]}")
