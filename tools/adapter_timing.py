"""Wall time of the exact JointProb adapter's forward(return_samples=True) on config-1 / config-5 shaped batches:
the padded read-out written on the device (ops.lattice_viterbi_padded, one host read) next to the ragged result
re-padded row by row on the host (what the adapter did before).  python tools/adapter_timing.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)


def wall(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for B in (1, 32, 512):
    p, _ = synth.transliteration_batch(B, seed=0).to(dev).pack()
    theta = torch.randn(p.vocab, device=dev)
    model = nb.ExactJointProb(theta, bos=synth.BOS, eos=synth.EOS, pad=synth.PAD)

    def old():
        num, _, _, _ = model.log_marginalize(p, None)
        _, off, _, labels = nb.lattice_viterbi(p, theta=theta)
        off_c = off.cpu()
        rows = []
        for b in range(p.n_lattices):
            lab = labels[int(off_c[b]): int(off_c[b + 1])].to(torch.int64)
            if lab.numel() and int(lab[0]) == synth.BOS:
                lab = lab[1:]
            rows.append(lab)
        T = max(r.numel() for r in rows)
        best = torch.full((p.n_lattices, T), synth.PAD, dtype=torch.int64, device=dev)
        for b, r in enumerate(rows):
            best[b, : r.numel()] = r
        return num, best

    new = lambda: model(p, None, return_samples=True)  # noqa: E731
    a, b = old(), new()
    best_new = b[2] if B > 1 else b[2].unsqueeze(0)
    assert torch.equal(a[1], best_new), "padded read-out differs from the re-padded ragged result"
    print(f"B={B:4d}: forward(return_samples=True) {wall(new):8.3f} ms (device read-out)   {wall(old):8.3f} ms (ragged result re-padded on the host)", flush=True)
