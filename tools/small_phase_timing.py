"""Per-phase cycle counts of nfst_small_kernel (needs the -DNFST_TIMING build, see phase_timing.py)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import _lib, synth  # noqa: E402
from nfst_b200.pack import concat_packed  # noqa: E402

dev = torch.device("cuda", 0)
lib = _lib.load()
buf = (C.c_ulonglong * 32)()
for name, gen, B in [("translit B=32", lambda n, o: synth.transliteration_batch(n, seed=o), 32),
                     ("SNIPS B=256", lambda n, o: synth.snips_batch(n, seed=1 + o), 256),
                     ("translit B=4096", lambda n, o: synth.transliteration_batch(n, seed=4 + o), 4096)]:
    parts, scs = [], []
    for o in range(0, B, 512):
        p, sc = gen(min(512, B - o), o).to(dev).pack()
        parts.append(p)
        scs.append(sc)
    packed = concat_packed(parts) if len(parts) > 1 else parts[0]
    sc = torch.cat(scs)
    for it in range(2):
        torch.cuda.synchronize()
        lib.nfst_debug_read(buf, 1)
        al, lz = nb.lattice_forward(packed, arc_scores=sc)
        torch.cuda.synchronize()
        lib.nfst_debug_read(buf, 1)
        vf = list(buf)
        nb.lattice_backward(packed, arc_scores=sc, alpha=al, logz=lz, want_beta=True, want_post=True)
        torch.cuda.synchronize()
        lib.nfst_debug_read(buf, 1)
        vb = list(buf)
    nf, nbk = max(vf[20], 1), max(vb[20], 1)
    print(f"{name}: fwd blocks {nf} levels/block {vf[21] / nf:.0f}: load {vf[16] / nf:.0f} cyc, levels {vf[17] / nf:.0f} cyc "
          f"({vf[17] / max(vf[21], 1):.0f}/level) | bwd blocks {nbk}: load {vb[16] / nbk:.0f}, levels {vb[18] / nbk:.0f} "
          f"({vb[18] / max(vb[21], 1):.0f}/level)")
