"""Random batches of the reference's dense [S, V] tables (tests/lattice_gen.py) through the drop-in calls against the
numpy oracle (test infrastructure: imports oracle/): compute_beta(emission, transition, theta, k) -- dense scan, device
packer, backward kernel, expansion to beta[B*k, S] in real space -- on collate()-padded batches, ExactJointProb.forward
with the best path, and the walker's exact samples.  python tests/fuzz/fuzz_dense.py [seconds] [seed]"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from oracle import lattice_oracle as lo  # noqa: E402
from tests.lattice_gen import BOS, EOS, PAD, random_mark_lattice  # noqa: E402

DEV = "cuda:0"
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def run_case():
    V = int(rng.choice([16, 40, 256]))
    B = int(rng.integers(1, 14))
    k = int(rng.choice([1, 2, 5]))
    par = bool(rng.integers(0, 2))
    tabs = [random_mark_lattice(rng, int(rng.integers(1, 60)), V, max_out=int(rng.integers(1, 4)), parallel_arcs=par)[1] for _ in range(B)]
    S = max(t.shape[0] for t in tabs)
    tr = np.full((B, S, V), PAD, dtype=np.int64)  # collate(): rows added by padding hold the pad id (quirk Q5)
    for b, t in enumerate(tabs):
        tr[b, : t.shape[0]] = t
    theta = (rng.normal(size=V) * float(rng.choice([0.3, 1.0, 3.0]))).astype(np.float32)
    trt, tht = torch.from_numpy(tr).to(DEV), torch.from_numpy(theta).to(DEV)
    beta = nb.compute_beta(trt != 0, trt, tht, k=k).cpu().numpy()
    assert beta.shape == (B * k, S) and beta.dtype == np.float32
    model = nb.ExactJointProb(tht, bos=BOS, eos=EOS, pad=PAD)
    num, den, best = model(trt != 0, trt, None, None, None, None, return_samples=True)
    best = best.cpu().numpy().reshape(B, -1)
    for b, t in enumerate(tabs):
        n = t.shape[0]
        s, l, d, _ = lo.arcs_from_dense(t)
        logz, _, o_beta, _ = lo.forward_backward(n, s, d, theta[l].astype(np.float64))
        with np.errstate(over="ignore"):
            want = np.exp(o_beta)
        fin = want < 3.0e38  # real-space float32, as scorers.py:854 returns it: larger values are +inf there too
        big = want > 3.5e38
        for j in range(k):
            # exp() of a float32 log-value of magnitude x carries ~6e-8 * x: 2e-5 up to |log beta| ~ 40 (the reference's
            # own float32 real-space values are no better), a few 1e-5 at 60-90 with theta three times the unit scale
            np.testing.assert_allclose(beta[b * k + j, :n][fin], want[fin], rtol=5e-5, atol=1e-30)
            assert np.all(np.isinf(beta[b * k + j, :n][big]))
            assert np.all(beta[b * k + j, n:] == 0.0)
        assert abs(float(num[b]) - (logz - float(theta[BOS]))) <= 1e-5 * max(1.0, abs(logz)), (float(num[b]), logz)
        _, _, vl, _, _ = lo.viterbi_f32(n, s, l, d, theta[l])
        row = best[b].tolist()
        assert row[: len(vl) - 1] == list(vl)[1:] and all(x == PAD for x in row[len(vl) - 1:]), (row, list(vl))
    return f"B={B} S={S} V={V} k={k} parallel_arcs={par}"


t0 = time.time()
n = fails = 0
while time.time() - t0 < budget:
    n += 1
    try:
        msg = run_case()
        if n % 50 == 1:
            print("ok  ", msg, flush=True)
    except Exception:  # noqa: BLE001
        fails += 1
        print(f"FAIL case {n}\n{traceback.format_exc()}", flush=True)
print(f"{n} dense batches, {fails} failures, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
