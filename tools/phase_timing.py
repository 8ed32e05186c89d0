"""Per-phase cycle counts of the forward kernel's pipeline (needs the -DNFST_TIMING variant):

    python -c "from nfst_b200 import build; build.build_library(out='nfst_b200/lib/libnfst_b200_timing.so', defines=['NFST_TIMING'])"
    NFST_LIB=nfst_b200/lib/libnfst_b200_timing.so python tools/phase_timing.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import nfst_b200 as nb  # noqa: E402
from nfst_b200 import _lib  # noqa: E402


class A:
    workload, arcs, batch, levels = "dag", 100_000, int(os.environ.get("BATCH", "592")), 64


packed, scores = bench.build_packed(A, torch.device("cuda", 0))
lib = _lib.load()
buf = (C.c_ulonglong * 32)()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(3):
    torch.cuda.synchronize()
    lib.nfst_debug_read(buf, 1)
    ev[0].record()
    alpha, logz = nb.lattice_forward(packed, arc_scores=scores)
    ev[1].record()
    torch.cuda.synchronize()
    lib.nfst_debug_read(buf, 0)
    v = list(buf)
    n = max(v[7], 1)
    print(f"fwd {ev[0].elapsed_time(ev[1]):.3f} ms, {n} chunk iterations; cycles per iteration: "
          f"lane0 issue+prefetch {v[0] / n:.0f} | consumer warp: wait arrays(it+2) {v[1] / n:.0f}, +gather issue {v[2] / n:.0f}, "
          f"wait arrays/scores(it) {v[3] / n:.0f}, reduce {v[4] / n:.0f} (last warp {v[5] / n:.0f}), whole iteration {v[6] / n:.0f}")
