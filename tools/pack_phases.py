"""Milliseconds per phase of the tensor-op packer on one bench chunk (NFST_PACK_TIMING=1 python tools/pack_phases.py)."""
import os
import sys

os.environ["NFST_PACK_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nfst_b200 import pack, synth  # noqa: E402

dev = torch.device("cuda", 0)
ab = synth.random_dag_batch(600, 100_000, levels=64, seed=3, device=dev)
ab.pack()
pack.phase_ms.clear()
for _ in range(3):
    ab.pack()
tot = sum(pack.phase_ms.values())
for k, v in pack.phase_ms.items():
    print(f"{k:55s} {v / 3:8.1f} ms  {100 * v / tot:5.1f}%")
print(f"{'total':55s} {tot / 3:8.1f} ms for {ab.src.numel()} arcs")
