"""Kernel-level timing of the small-lattice configs (1, 2, 5): forward, fused backward and
Viterbi calls, each timed with CUDA events over `--reps` back-to-back calls (launch overhead
included; run under `ncu --metrics gpu__time_duration.sum -k regex:nfst_` for pure kernel time).

    python tools/small_cases.py [--reps 20]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import synth  # noqa: E402
from nfst_b200.pack import concat_packed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--only", type=int, default=-1, help="run just this case (0, 1, 2)")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps


CASES = [
    ("config1 translit B=32", lambda n, o: synth.transliteration_batch(n, seed=o), 32),
    ("config2 SNIPS B=256", lambda n, o: synth.snips_batch(n, seed=1 + o), 256),
    ("config5 translit B=4096", lambda n, o: synth.transliteration_batch(n, seed=4 + o), 4096),
]
for name, gen, B in (CASES if a.only < 0 else CASES[a.only:a.only + 1]):
    parts, scs = [], []
    for o in range(0, B, 512):
        p, sc = gen(min(512, B - o), o).to(dev).pack()
        parts.append(p)
        scs.append(sc)
    packed = concat_packed(parts) if len(parts) > 1 else parts[0]
    sc = torch.cat(scs)
    al, lz = nb.lattice_forward(packed, arc_scores=sc)
    t_f = timed(lambda: nb.lattice_forward(packed, arc_scores=sc))
    t_b = timed(lambda: nb.lattice_backward(packed, arc_scores=sc, alpha=al, logz=lz, want_beta=True, want_post=True))
    t_fb = timed(lambda: nb.lattice_forward_backward(packed, arc_scores=sc))
    t_v = timed(lambda: nb.lattice_viterbi(packed, arc_scores=sc))
    g = [(x.n, x.block_threads, x.small_max_arcs, x.small_max_states) for x in packed.groups]
    print(f"{name:26s} A={packed.n_arcs:8d} S={packed.n_states:8d} L={packed.max_levels:4d} state={al.dtype} "
          f"fwd {t_f * 1e3:7.1f} us  bwd {t_b * 1e3:7.1f} us  fused f+b {t_fb * 1e3:7.1f} us  viterbi {t_v * 1e3:7.1f} us  groups {g}",
          flush=True)
