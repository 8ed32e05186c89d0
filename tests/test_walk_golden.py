"""The numpy restatement of the reference's sampling-loop step (oracle/walk_oracle.py) against outputs of the
UNMODIFIED reference (tests/golden/walk_step.npz, made by tests/golden/make_golden.py: gen_walk)."""
import os

import numpy as np

from oracle import walk_oracle as wo

G = os.path.join(os.path.dirname(__file__), "golden", "walk_step.npz")


def test_walk_oracle_matches_reference_loop():
    g = np.load(G)
    k, T, pad = int(g["k"]), float(g["temperature"]), int(g["pad"])
    for t in range(int(g["steps"])):
        masked, logp, logz, nxt = wo.walk_step_dense(
            g["em"], g["tr"], k, g["beta"], g[f"state_old_{t}"], g[f"state_new_{t}"], g[f"prefix_{t}"],
            g[f"base_{t}"], pad, T, g[f"sym_{t}"])
        ref = g[f"masked_{t}"].astype(np.float64)
        assert np.array_equal(np.isneginf(masked), np.isneginf(ref))
        fin = np.isfinite(ref)
        np.testing.assert_allclose(masked[fin], ref[fin], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(logp, g[f"logp_{t}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(logz, g[f"zs_{t}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_array_equal(nxt, g[f"next_{t}"])
        if t + 1 < int(g["steps"]):  # the loop's own bookkeeping: the next step starts where this one ended
            np.testing.assert_array_equal(g[f"state_old_{t + 1}"], g[f"state_new_{t}"])
            np.testing.assert_array_equal(g[f"state_new_{t + 1}"], nxt)


def test_walk_oracle_replays_a_full_stateful_sample_call():
    """tests/golden/stateful_sample.npz: one full call of the reference's Sampler.stateful_sample on a real
    FSAGRUScorer(use_beta=True) (make_golden.py: gen_stateful).  The step oracle, chained from the start state over the
    recorded network outputs, masks and samples, reproduces the summed log-probabilities the call returned; the
    recorded beta is the reference recurrence with theta = W tanh(Wx e + b) (scorers.py:732-738, Wh = 0)."""
    from oracle import lattice_oracle as lo

    g = np.load(os.path.join(os.path.dirname(G), "stateful_sample.npz"))
    k, T, pad, bos, steps = int(g["k"]), float(g["temperature"]), int(g["pad"]), int(g["bos"]), int(g["steps"])
    tr, em, beta, seqs = g["tr"], g["em"], g["beta"], g["sequences"]
    B, S, V = tr.shape
    N = B * k
    # beta: the reference's value against the oracle recurrence, on the states the start reaches
    theta = g["theta"].astype(np.float64)
    for b in range(B):
        src, lab, dst, _ = lo.arcs_from_dense(tr[b])
        ref = np.exp(lo.beta_log(S, src, dst, theta[lab]))
        reach = np.zeros(S, dtype=bool)
        reach[0] = True
        for _ in range(S):
            reach[dst[reach[src]]] = True
        np.testing.assert_allclose(beta[b * k][reach], ref[reach], rtol=2e-5)
    # the loop: bos is consumed first (scorers.py:230-231), the look-ahead state lags one step (quirk Q9)
    rows = np.arange(N) // k
    state_old = np.zeros(N, dtype=np.int64)
    state_new = tr[rows, state_old, bos]
    total = np.zeros(N)
    for t in range(steps):
        sym = seqs[:, t] if t < seqs.shape[1] else np.full(N, pad)
        _, logp, _, nxt = wo.walk_step_dense(em, tr, k, beta, state_old, state_new, g[f"prefix_{t}"], g[f"base_{t}"], pad, T, sym)
        total += logp
        state_old, state_new = state_new, nxt
    np.testing.assert_allclose(total, g["summed_log_probs"], rtol=1e-5, atol=1e-5)
