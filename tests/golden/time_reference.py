"""Time the UNMODIFIED Python reference on BASELINE configs 1 and 2 (SURVEY.md §8(d) "reference CPU timing",
BASELINE.md §4 item 1).  Run from the repo root, in the build container (it needs /root/reference):

    python tests/golden/time_reference.py [--configs 1 2] [--procs N] [--limit L] > profiles/<round>_reference_python_timing.txt

What is timed: ``FSAGRUScorer.compute_beta_per_sample(transition[S, V])`` (``scorers.py:692-751``) -- the reference's
own implementation of the backward recurrence, graph build from the dense table included, as shipped (float32,
hidden size 256 = ``conf/train/lstm.yaml``), imported through ``oracle/ref_harness.py`` (stubs for the absent
OpenFst wrappers only; none of this repo's kernels on the path).  ``Wh`` is zeroed so that the recurrence is the
log-semiring backward pass the CUDA path computes (the arithmetic per arc is unchanged: the ``Wh`` matvec still
runs).  The batch is spread over the host cores with ``multiprocessing``, one torch thread per process.  Every
``beta[start]`` is checked against the float64 oracle with theta = W.tanh(Wx e + b).

The Python reference cannot travel to the GPU box (``bench.py --impl reference`` times the C port there, a far
stronger baseline); this number is therefore a build-container figure, reported beside the others in DESIGN.md §6.
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

H, V = 256, 256
_m = None


def _init():
    global _m
    import torch

    torch.set_num_threads(1)
    from oracle import ref_harness as rh

    _m = rh.make_scorer(H, V, seed=7, zero_wh=True, double=False)


def _one(tr: np.ndarray):
    import torch

    t0 = time.perf_counter()
    with torch.no_grad():
        beta = _m.compute_beta_per_sample(torch.from_numpy(tr).int())
    return time.perf_counter() - t0, float(beta[0])


def dense_tables(ab):
    """one int64 [S_b, V] table per lattice, the reference's format (scorers.py:995-1035): cell = next state, 0 = no arc"""
    from nfst_b200 import synth

    tabs = []
    lat, src, dst, lab = (x.numpy() for x in (ab.arc_lattice, ab.src, ab.dst, ab.label))
    for b in range(int(ab.n_states.numel())):
        S = int(ab.n_states[b])
        tr = np.zeros((S, V), dtype=np.int64)
        sel = lat == b
        tr[src[sel], lab[sel]] = dst[sel]
        sink = np.setdiff1d(np.arange(S), src[sel])
        tr[sink, synth.PAD] = sink  # the sink's pad self-loop (scorers.py:1013-1016)
        tabs.append(tr)
    return tabs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, nargs="+", default=[1, 2])
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--limit", type=int, default=0, help="time only the first L lattices of a config (0 = the stated batch)")
    args = ap.parse_args()
    import torch

    from nfst_b200 import synth
    from oracle import lattice_oracle as lo
    from oracle import ref_harness as rh

    assert rh.available(), "the reference is not mounted: this script runs in the build container only"
    m = rh.make_scorer(H, V, seed=7, zero_wh=True, double=False)
    theta = rh.arc_theta(m).double().numpy()
    print(f"# unmodified reference, FSAGRUScorer.compute_beta_per_sample (scorers.py:692-751), float32, H={H}, V={V}, Wh=0; "
          f"torch {torch.__version__}, {args.procs} processes x 1 thread on {os.cpu_count()} host cores (build container, not the GPU box)")
    gens = {1: ("config1 transliteration B=32", lambda: synth.transliteration_batch(32, seed=0)),
            2: ("config2 SNIPS B=256", lambda: synth.snips_batch(256, seed=1))}
    for c in args.configs:
        name, gen = gens[c]
        ab = gen()
        tabs = dense_tables(ab)
        if args.limit:
            tabs = tabs[: args.limit]
        arcs = [int(np.count_nonzero((t != 0) & (t != np.arange(t.shape[0])[:, None]))) for t in tabs]
        with mp.get_context("spawn").Pool(args.procs, initializer=_init) as pool:
            pool.map(_one, tabs[: args.procs])  # warm-up: imports, first-call allocations
            t0 = time.perf_counter()
            res = pool.map(_one, tabs, chunksize=1)
            wall = time.perf_counter() - t0
        per = np.array([r[0] for r in res])
        worst = 0.0
        for tr, (_, b0) in zip(tabs, res):
            s, l, d, _ = lo.arcs_from_dense(tr)
            logz = lo.forward_backward(tr.shape[0], s, d, theta[l])[0]
            worst = max(worst, abs(np.log(b0) - logz) / max(1.0, abs(logz)))
        assert worst < 1e-3, worst
        print(f"{name}: {len(tabs)} lattices, {sum(arcs)} arcs, states/lattice {min(t.shape[0] for t in tabs)}..{max(t.shape[0] for t in tabs)}; "
              f"wall {wall:.2f} s on {args.procs} processes = {sum(arcs) / wall:.0f} arcs/s; "
              f"per lattice on one core {per.min():.3f}..{per.max():.3f} s (median {np.median(per):.3f}) = {sum(arcs) / per.sum():.0f} arcs/s per core; "
              f"log beta[start] vs the float64 oracle: max rel. diff {worst:.1e}", flush=True)


def parallel_leg(B: int = 4, hid: int = 8):
    """``compute_beta()`` -> ``compute_beta_parallel`` (scorers.py:753-875), the batched form the sampler calls, on the
    first ``B`` lattices of config 1 collate-padded.  At the shipped H = 256 its ``[B, S, S, H]`` work tensors need tens of
    GB, so this leg runs at H = 8 and says so; one process, torch's own threads."""
    import torch

    from nfst_b200 import synth
    from oracle import lattice_oracle as lo
    from oracle import ref_harness as rh

    tabs = dense_tables(synth.transliteration_batch(32, seed=0))[:B]
    tr = lo.collate_pad(tabs, synth.PAD)
    m = rh.make_scorer(hid, V, seed=7, zero_wh=True, double=False)
    theta = rh.arc_theta(m).double().numpy()
    arcs = sum(int(np.count_nonzero((t != 0) & (t != np.arange(t.shape[0])[:, None]))) for t in tabs)
    with torch.no_grad():
        m.set_masks(emission=torch.from_numpy(tr != 0), transition=torch.from_numpy(tr))
        m.set_k(1)
        t0 = time.perf_counter()
        beta = m.compute_beta()
        wall = time.perf_counter() - t0
    worst = 0.0
    for b, t in enumerate(tabs):
        s, l, d, _ = lo.arcs_from_dense(t)
        logz = lo.forward_backward(t.shape[0], s, d, theta[l])[0]
        worst = max(worst, abs(float(beta[b, 0].log()) - logz) / max(1.0, abs(logz)))
    assert worst < 1e-3, worst
    print(f"config1, first {B} lattices through compute_beta() = compute_beta_parallel (H={hid}, not the shipped 256: memory), "
          f"tables [{B}, {tr.shape[1]}, {V}], {arcs} arcs: {wall:.2f} s = {arcs / wall:.0f} arcs/s "
          f"({torch.get_num_threads()} torch threads); log beta[start] vs the float64 oracle: max rel. diff {worst:.1e}", flush=True)


if __name__ == "__main__":
    main()
    parallel_leg()
