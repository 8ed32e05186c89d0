"""Time the sliced-column passes over block sizes (one packed batch, block_threads overridden per run).
usage: python tools/sell_tune.py [arcs_per_lattice] [batch] [levels]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nfst_b200 as nb  # noqa: E402
from nfst_b200 import ops  # noqa: E402


class A:
    workload = "dag"
    levels = 64


def main():
    a = A()
    a.arcs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    a.batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    a.levels = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    dev = torch.device("cuda", 0)
    packed, scores = bench.build_packed(a, dev)
    Ar, S = packed.n_arcs, packed.n_states
    print(f"A={Ar} S={S} groups={[(g.sell, g.block_threads, g.n, g.sell_window) for g in packed.groups]}")
    beta = torch.empty(S, device=dev)
    default = [g.block_threads for g in packed.groups]
    variants = [int(x) for x in os.environ.get("THREADS", "0,32,64,128,256,512").split(",")]
    for th in variants:
        for g, d in zip(packed.groups, default):
            g.block_threads = th if th else d

        def step(ev=None):
            if ev:
                ev[0].record()
            logz, alpha, cond = ops.lattice_pull(packed, arc_scores=scores, beta_out=beta)
            if ev:
                ev[1].record()
            r = nb.lattice_backward(packed, arc_scores=scores, alpha=alpha, logz=logz, cond=cond, want_beta=False,
                                    want_post=True)
            if ev:
                ev[2].record()
            return logz, r

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        n = 10
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
        for e in evs:
            step(e)
        torch.cuda.synchronize()
        t1 = sum(e[0].elapsed_time(e[1]) for e in evs) / n
        t2 = sum(e[1].elapsed_time(e[2]) for e in evs) / n
        vit0, vit1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nb.ops.lattice_backward(packed, arc_scores=scores, want_beta=False, want_viterbi=True)
        vit0.record()
        for _ in range(n):
            nb.ops.lattice_backward(packed, arc_scores=scores, want_beta=False, want_viterbi=True)
        vit1.record()
        torch.cuda.synchronize()
        tv = vit0.elapsed_time(vit1) / n
        print(f"threads={th or default}: pull {t1:.3f} ms ({(12 * Ar + 4 * S) / t1 / 1e6:.0f} GB/s)  flow {t2:.3f} ms "
              f"({12 * Ar / t2 / 1e6:.0f} GB/s)  step {t1 + t2:.3f} ms = {Ar / (t1 + t2) / 1e6:.1f} Garc/s "
              f"({(20 * Ar + 20 * S) / (t1 + t2) / 1e6:.0f} GB/s)  viterbi {tv:.3f} ms = {Ar / tv / 1e6:.1f} Garc/s", flush=True)


if __name__ == "__main__":
    main()
