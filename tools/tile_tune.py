"""Time the tile-stream passes of one config-4 sweep point under the current NFST_TILE_* environment.
usage: [NFST_TILE_WARPS=.. NFST_TILE_ARCS=.. ...] python tools/tile_tune.py [arcs_per_lattice] [batch] [levels]
Prints the launch groups (warps, ring, far table, stage sizes, shared memory) and pull / flow / Viterbi times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import nfst_b200 as nb  # noqa: E402
from nfst_b200 import _lib, ops  # noqa: E402


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
    env = {k: v for k, v in os.environ.items() if k.startswith("NFST_")}
    desc = []
    for g in packed.groups:
        if g.tiles:
            desc.append(f"tiles n={g.n} warps={g.block_threads // 32} ring={g.tile_ring} cap_arcs={g.tile_cap_arcs} cap_bytes={g.tile_cap_bytes}")
        else:
            desc.append(f"{'sell' if g.sell else 'csr'} n={g.n} threads={g.block_threads}")
    print(f"arcs/lattice={a.arcs} A={Ar} S={S} env={env}\n  groups: {desc}\n  stream bytes/arc={packed.tile_stream.numel() / max(Ar, 1):.2f}", flush=True)
    beta = torch.empty(S, dtype=ops.resolve_state_dtype(packed), device=dev)
    sell_only = all(g.sell or g.tiles for g in packed.groups)

    def step(ev=None):
        if ev:
            ev[0].record()
        logz, alpha, cond = ops.lattice_pull(packed, arc_scores=scores, beta_out=beta)
        if ev:
            ev[1].record()
        nb.lattice_backward(packed, arc_scores=scores, alpha=alpha, logz=logz, cond=cond, want_beta=not sell_only, want_post=True)
        if ev:
            ev[2].record()

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
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb.ops.lattice_backward(packed, arc_scores=scores, want_beta=False, want_viterbi=True)
    v0.record()
    for _ in range(n):
        nb.ops.lattice_backward(packed, arc_scores=scores, want_beta=False, want_viterbi=True)
    v1.record()
    torch.cuda.synchronize()
    tv = v0.elapsed_time(v1) / n
    # theta mode (scores = theta[label]): pull with the label lookup, flow with the label histogram and no per-arc output
    theta = -torch.rand(packed.vocab, device=dev)
    tev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
    for e in [None] * 3 + tev:
        if e:
            e[0].record()
        lz, al, cd = ops.lattice_pull(packed, theta=theta, beta_out=beta)
        if e:
            e[1].record()
        nb.lattice_backward(packed, theta=theta, alpha=al, logz=lz, cond=cd, want_beta=not sell_only, want_post=False, want_dtheta=True)
        if e:
            e[2].record()
    torch.cuda.synchronize()
    tp = sum(e[0].elapsed_time(e[1]) for e in tev) / n
    tf = sum(e[1].elapsed_time(e[2]) for e in tev) / n
    print(f"  theta mode: pull {tp:.3f} ms  flow+dtheta {tf:.3f} ms  step {tp + tf:.3f} ms = {Ar / (tp + tf) / 1e6:.1f} Garc/s", flush=True)
    # the same step replayed from a CUDA graph: device time without the host's per-call work
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(3):
        g.replay()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(n):
        g.replay()
    g1.record()
    torch.cuda.synchronize()
    tg = g0.elapsed_time(g1) / n
    peak = bench.measured_peak_gbs()[0]
    gbs = (20 * Ar + 20 * S) / (t1 + t2) / 1e6
    print(f"  pull {t1:.3f} ms  flow {t2:.3f} ms  step {t1 + t2:.3f} ms = {Ar / (t1 + t2) / 1e6:.1f} Garc/s, {gbs:.0f} GB/s = {gbs / peak:.3f} of peak; "
          f"viterbi pass {tv:.3f} ms; graph replay of the step {tg:.3f} ms = {(20 * Ar + 20 * S) / tg / 1e6 / peak:.3f} of peak", flush=True)


if __name__ == "__main__":
    main()
