"""Host-side cost of getting lattices to the kernels: pack (once per example / batch) and
concat_packed (per step, when packs are cached per example).  `--profile` prints a cProfile of one
warm pack call."""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import synth  # noqa: E402
from nfst_b200.pack import concat_packed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def wall(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


one = synth.transliteration_batch(1, seed=0).to(dev)
b32 = synth.transliteration_batch(32, seed=0).to(dev)
b4k = synth.transliteration_batch(512, seed=0).to(dev)
print(f"pack 1 lattice: {wall(lambda: one.pack()):.2f} ms   32 lattices: {wall(lambda: b32.pack()):.2f} ms   "
      f"512 lattices: {wall(lambda: b4k.pack()):.2f} ms")
parts = [synth.transliteration_batch(1, seed=i).to(dev).pack()[0] for i in range(32)]
print(f"concat_packed of 32 cached single-lattice packs: {wall(lambda: concat_packed(parts)):.2f} ms")
big = concat_packed(parts)
th = -torch.rand(big.vocab, device=dev)
print(f"forward + backward on it: {wall(lambda: nb.lattice_forward_backward(big, theta=th), 20):.3f} ms")
if a.profile:
    pr = cProfile.Profile()
    pr.enable()
    b32.pack()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
