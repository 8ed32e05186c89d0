"""Random batches through the sampling-loop kernels and the beta-hat recurrence against their oracles (test
infrastructure: imports oracle/): sample_paths -- every sampled row is a start->sink path, log q = score - logZ (so the
k-sample IWAE estimate of estimatros.py:32-44 equals logZ with zero variance) on small-lattice, CSR and column-major
packs; lattice_beta_hat (scorers.py:732-747, Wh != 0) vs the float64 numpy port.  python tests/fuzz/fuzz_walk.py [seconds] [seed]"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import synth  # noqa: E402
from nfst_b200.sampler import sample_paths  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle import lattice_oracle as lo  # noqa: E402

DEV = "cuda:0"
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def batch():
    k = int(rng.integers(0, 4))
    s = int(rng.integers(0, 10**6))
    if k == 0:
        return synth.transliteration_batch(int(rng.integers(1, 12)), seed=s)
    if k == 1:
        return synth.snips_batch(int(rng.integers(1, 4)), seed=s)
    if k == 2:
        return synth.cipher_batch(int(rng.integers(1, 3)), T=int(rng.choice([20, 60])), bigram=bool(rng.integers(0, 2)), seed=s)
    return synth.random_dag_batch(int(rng.integers(1, 4)), int(np.exp(rng.uniform(np.log(300), np.log(40_000)))),
                                  levels=int(rng.choice([3, 8, 24, 64])), seed=s)


def case_samples():
    ab = batch()
    p, sc = ab.to(DEV).pack()
    k = int(rng.choice([4, 32]))
    T = max(p.max_levels - 1, 1)
    u = torch.from_numpy(rng.random((p.n_lattices * k, T)).astype(np.float32)).to(DEV)
    labels, length, log_q, arcs, logz = sample_paths(p, k, arc_scores=sc, uniform=u, pad_id=synth.PAD)
    ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(), ab.n_states.numpy())
    o_logz = c_oracle.forward_backward(ob, want_post=False, want_states=False)[0]
    np.testing.assert_allclose(logz.cpu().numpy(), o_logz, rtol=1e-5, atol=1e-5)
    arcs_c, len_c, lq, lab = arcs.cpu().numpy(), length.cpu().numpy(), log_q.cpu().numpy(), labels.cpu().numpy()
    w, dst, src = sc.cpu().numpy().astype(np.float64), p.dst_out.cpu().numpy(), p.src_out.cpu().numpy()
    deg = np.bincount(src, minlength=p.n_states)
    start, label_out = p.start_state.cpu().numpy(), p.label_out.cpu().numpy()
    for r in range(p.n_lattices * k):
        b, path = r // k, arcs_c[r, :len_c[r]]
        assert len(path) >= 1 and src[path[0]] == start[b] and np.array_equal(dst[path[:-1]], src[path[1:]])
        assert deg[dst[path[-1]]] == 0 and np.all(arcs_c[r, len_c[r]:] == -1)
        assert np.array_equal(lab[r, :len_c[r]], label_out[path]) and np.all(lab[r, len_c[r]:] == synth.PAD)
        assert abs(lq[r] - (w[path].sum() - o_logz[b])) < 1e-5 * max(1.0, abs(o_logz[b])) + 2e-4, (lq[r], w[path].sum() - o_logz[b])
    kinds = sorted({"tiles" if g.tiles else "sell" if g.sell else "small" if g.small_max_arcs > 0 else "csr" for g in p.groups})
    return f"samples {kinds} B={p.n_lattices} k={k} arcs={p.n_arcs} levels={p.max_levels}"


def case_beta_hat():
    ab = synth.transliteration_batch(int(rng.integers(1, 5)), seed=int(rng.integers(0, 10**6))) if rng.integers(0, 2) else \
        synth.random_dag_batch(2, int(rng.integers(200, 3000)), levels=int(rng.choice([4, 12, 30])), seed=int(rng.integers(0, 10**6)))
    p, _ = ab.to(DEV).pack()
    H, V = int(rng.choice([4, 16, 64])), ab.vocab
    g = torch.Generator().manual_seed(int(rng.integers(0, 10**6)))
    emb = torch.randn(V, H, generator=g, dtype=torch.float64)
    Wx, Wh = (torch.randn(H, H, generator=g, dtype=torch.float64) / H ** 0.5 for _ in range(2))
    W = torch.randn(1, H, generator=g, dtype=torch.float64) / H ** 0.5
    bias = 0.1 * torch.randn(H, generator=g, dtype=torch.float64)
    log_beta, beta_hat = nb.ops.lattice_beta_hat(p, (emb @ Wx.T + bias).to(DEV), Wh.to(DEV), W.to(DEV))
    log_beta, beta_hat = log_beta.cpu().numpy().astype(np.float64), beta_hat.cpu().numpy().astype(np.float64)
    lat, src, dst, lab = (t.numpy() for t in (ab.arc_lattice, ab.src, ab.dst, ab.label))
    so, orig = p.state_off.cpu().numpy(), p.orig_state.cpu().numpy()
    for b in range(p.n_lattices):
        m = lat == b
        be, bh = lo.beta_recurrent(int(ab.n_states[b]), src[m], lab[m], dst[m], emb.numpy(), Wx.numpy(), Wh.numpy(), W.numpy(), bias.numpy())
        o = orig[so[b]:so[b + 1]]
        np.testing.assert_allclose(log_beta[so[b]:so[b + 1]], np.log(be[o]), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(beta_hat[so[b]:so[b + 1]], bh[o], rtol=0, atol=1e-4)
    return f"beta-hat H={H} B={p.n_lattices} arcs={p.n_arcs} levels={p.max_levels}"


t0 = time.time()
n = fails = 0
while time.time() - t0 < budget:
    n += 1
    try:
        msg = case_samples() if n % 3 else case_beta_hat()
        if n % 25 == 1:
            print("ok  ", msg, flush=True)
    except Exception:  # noqa: BLE001
        fails += 1
        print(f"FAIL case {n}\n{traceback.format_exc()}", flush=True)
print(f"{n} cases (2/3 sample_paths, 1/3 beta-hat), {fails} failures, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
