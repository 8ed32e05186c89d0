"""theta-mode gradient on the small-lattice / CSR kernels (the label histogram): CUDA-event time per call of
lattice_forward_backward(theta, want_dtheta=True) and of the autograd training step (-mean logZ -> d theta) on configs 1, 2, 5
and a cipher batch.  python tools/dtheta_timing.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import synth  # noqa: E402
from nfst_b200.pack import concat_packed  # noqa: E402

dev = torch.device("cuda", 0)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


cases = [("config1 translit B=32", lambda: synth.transliteration_batch(32, seed=0)),
         ("config2 SNIPS B=256", lambda: synth.snips_batch(256, seed=1)),
         ("config5 translit B=4096", None),
         ("cipher bigram T=200 B=16", lambda: synth.cipher_batch(16, T=200, bigram=True, seed=2))]
for name, gen in cases:
    if gen is None:
        parts = [synth.transliteration_batch(512, seed=4 + o).to(dev).pack()[0] for o in range(0, 4096, 512)]
        p = concat_packed(parts)
    else:
        p, _ = gen().to(dev).pack()
    theta = (-torch.rand(p.vocab, device=dev)).requires_grad_(True)

    def step():
        theta.grad = None
        (-nb.lattice_log_partition(p, theta=theta).mean()).backward()

    t_fb = timed(lambda: nb.lattice_forward_backward(p, theta=theta.detach(), want_dtheta=True))
    t_step = timed(step)
    print(f"{name:28s} A={p.n_arcs:9d}  forward_backward + dtheta {t_fb:8.1f} us   autograd training step {t_step:8.1f} us", flush=True)
