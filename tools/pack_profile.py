"""Where the packer's time goes: wall time, kernel launches and the heaviest ops of pack_arcs on two batches.
usage: python tools/pack_profile.py"""
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nfst_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
def dense_config1():
    """config 1 as the reference holds it: collate()-padded dense [B, S, V] tables (pad rows carry V arcs each)"""
    ab = synth.transliteration_batch(32, seed=0)
    B, S, V = 32, int(ab.n_states.max()), ab.vocab
    tr = torch.full((B, S, V), synth.PAD, dtype=torch.int64)
    for b in range(B):
        tr[b, : int(ab.n_states[b])] = 0
    tr[ab.arc_lattice, ab.src, ab.label] = ab.dst
    tr = tr.to(dev)

    class T:
        src = tr
        def pack(self):
            import nfst_b200 as nb
            return nb.pack_dense(tr != 0, tr), None
    return T()


cases = [("config1 32 dense collate()-padded tables [32, S, 256] (pack_dense)", dense_config1),
         ("config1 32 edit lattices", lambda: synth.transliteration_batch(32, seed=0).to(dev)),
         ("config5 4096 edit lattices", lambda: synth.transliteration_batch(4096, seed=4).to(dev)),
         ("config2 256 SNIPS grids", lambda: synth.snips_batch(256, seed=1).to(dev)),
         ("bench chunk 600 x 100k-arc DAGs", lambda: synth.random_dag_batch(600, 100_000, levels=64, seed=3, device=dev))]
for name, gen in cases:
    ab = gen()
    ab.pack()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ab.pack()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        ab.pack()
        torch.cuda.synchronize()
    ev = prof.key_averages()
    n_kernels = sum(e.count for e in ev if e.device_type == torch.autograd.DeviceType.CUDA)
    cuda_ms = sum(e.self_device_time_total for e in ev) / 1e3
    print(f"== {name}: {ab.src.numel()} arcs/cells, pack {dt * 1e3:.1f} ms wall, {n_kernels} kernel launches, {cuda_ms:.1f} ms of device time")
    print(ev.table(sort_by="self_cuda_time_total", row_limit=14, max_name_column_width=60))
