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
