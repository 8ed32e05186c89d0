"""Pin the CPU oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz).

The fixtures were produced by tests/golden/make_golden.py calling the reference's own
compute_beta_per_sample (scorers.py:692-751), compute_beta_parallel (:753-856) and
Estimators.iwae (estimatros.py:32-44).
"""
import os

import numpy as np
import pytest

from oracle import lattice_oracle as lo
from tests.lattice_gen import BOS, PAD

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def test_beta_log_matches_reference_per_sample():
    g = _load("beta_per_sample.npz")
    theta = g["theta0"]
    for i in range(int(g["n_cases"])):
        tr = g[f"tr_{i}"]
        ref = g[f"beta0_{i}"]  # real space, float64
        src, lab, dst, _ = lo.arcs_from_dense(tr)
        beta = lo.beta_log(tr.shape[0], src, dst, theta[lab])
        np.testing.assert_allclose(np.exp(beta), ref, rtol=1e-12, atol=0)
        # log-space agreement wherever the reference did not underflow
        ok = ref > 0
        np.testing.assert_allclose(beta[ok], np.log(ref[ok]), rtol=0, atol=1e-10)


def test_beta_recurrent_matches_reference_per_sample():
    g = _load("beta_per_sample.npz")
    P = {k: g["p1_" + k] for k in ("emb", "Wx", "Wh", "W", "bias")}
    for i in range(int(g["n_cases"])):
        tr = g[f"tr_{i}"]
        ref = g[f"beta1_{i}"]
        src, lab, dst, _ = lo.arcs_from_dense(tr)
        beta, _ = lo.beta_recurrent(tr.shape[0], src, lab, dst, P["emb"], P["Wx"], P["Wh"], P["W"], P["bias"])
        np.testing.assert_allclose(beta, ref, rtol=1e-10, atol=0)


def test_beta_log_matches_reference_parallel_with_collate_padding():
    g = _load("beta_parallel.npz")
    theta = g["theta0"]
    k = int(g["k"])
    for bi in range(int(g["n_batches"])):
        tr = g[f"tr_{bi}"]  # [B, S, V] padded with pad id (quirk Q5)
        ref = g[f"beta_{bi}"]  # [B*k, S] float32
        n_states = g[f"n_states_{bi}"]
        B = tr.shape[0]
        assert ref.shape[0] == B * k
        for b in range(B):
            np.testing.assert_array_equal(ref[b * k], ref[b * k + 1])  # repeat_interleave(k)
            nb = int(n_states[b])
            # the real rows of the padded table; padded rows only add arcs *out of* rows
            # >= nb, which never fire in the batched reference (Q4/Q5) and keep beta = 0.
            src, lab, dst, _ = lo.arcs_from_dense(tr[b])
            real = src < nb
            beta = lo.beta_log(nb, src[real], dst[real], theta[lab[real]])
            np.testing.assert_allclose(np.exp(beta), ref[b * k, :nb], rtol=2e-5)
            pad_rows = ref[b * k, nb:]
            # padded rows: 0, except the row whose index is the pad id (a spurious sink, Q5)
            for r, v in enumerate(pad_rows, start=nb):
                assert v == 0.0 or (r == PAD and v == 1.0)


def test_logz_is_bracketed_by_reference_iwae():
    g = _load("iwae.npz")
    theta = g["theta"]
    for i in range(int(g["n_cases"])):
        tr = g[f"tr_{i}"]
        est = g[f"iwae_{i}"]
        src, lab, dst, _ = lo.arcs_from_dense(tr)
        logz = lo.beta_log(tr.shape[0], src, dst, theta[lab])[0]
        target = logz - theta[BOS]  # the sampler never emits bos (scorers.py:230-231)
        se = est.std(ddof=1) / np.sqrt(len(est))
        # IWAE is a lower bound in expectation that tightens with k; allow 5 s.e. + bias
        assert est.mean() <= target + 5 * se + 1e-3
        assert abs(est.mean() - target) < max(6 * se, 0.05), (est.mean(), target, se)


def test_alpha_beta_identities_and_brute_force():
    rng = np.random.default_rng(5)
    from tests.lattice_gen import random_mark_lattice

    for n_inner in [1, 3, 6, 9]:
        for par in (False, True):
            _, tr = random_mark_lattice(rng, n_inner, 24, parallel_arcs=par)
            S = tr.shape[0]
            src, lab, dst, _ = lo.arcs_from_dense(tr)
            w = rng.normal(size=len(src))
            logz, alpha, beta, post = lo.forward_backward(S, src, dst, w)
            bz, bpost, paths, scores = lo.brute_force(S, src, dst, w)
            assert abs(logz - bz) < 1e-10
            np.testing.assert_allclose(post, bpost, atol=1e-12)
            assert abs(alpha[S - 1] - logz) < 1e-10
            # flow conservation at every state
            gamma = np.exp(alpha + beta - logz)
            inflow = np.bincount(dst, weights=post, minlength=S)
            outflow = np.bincount(src, weights=post, minlength=S)
            np.testing.assert_allclose(outflow[:-1], gamma[:-1], atol=1e-12)
            np.testing.assert_allclose(inflow[1:], gamma[1:], atol=1e-12)
            # finite differences of logZ w.r.t. arc scores
            eps = 1e-6
            for a in range(0, len(w), max(1, len(w) // 5)):
                w2 = w.copy()
                w2[a] += eps
                lz2 = lo.beta_log(S, src, dst, w2)[0]
                assert abs((lz2 - logz) / eps - post[a]) < 1e-5


def test_viterbi_rule_against_brute_force_with_ties():
    rng = np.random.default_rng(6)
    from tests.lattice_gen import random_mark_lattice

    for trial in range(30):
        _, tr = random_mark_lattice(rng, int(rng.integers(1, 10)), 24, parallel_arcs=bool(trial % 2))
        S = tr.shape[0]
        src, lab, dst, _ = lo.arcs_from_dense(tr)
        w = rng.integers(-2, 1, size=len(src)).astype(np.float32)  # many exact ties
        score, path, labels, delta, bp = lo.viterbi_f32(S, src, lab, dst, w)
        paths = lo.enumerate_paths(S, src, dst)
        best = max(float(lo.path_score_f32_backward(w, p)) for p in paths)
        assert float(score) == best
        tied = [[int(lab[a]) for a in p] for p in paths if float(lo.path_score_f32_backward(w, p)) == best]
        # integer scores: fp32 sums are exact, so the rule must pick the lexicographically
        # smallest label sequence among the optimal paths
        assert list(labels) == min(tied)
        assert float(lo.path_score_f32_backward(w, path)) == best


def _batch_from_tables(tables, theta):
    from oracle import c_oracle

    lat, src, dst, lab, sc, ns = [], [], [], [], [], []
    for b, tr in enumerate(tables):
        s, l, d, _ = lo.arcs_from_dense(tr)
        lat.append(np.full(len(s), b)); src.append(s); dst.append(d); lab.append(l); sc.append(theta[l]); ns.append(tr.shape[0])
    cat = np.concatenate
    return c_oracle.Batch(cat(lat), cat(src), cat(dst), cat(lab), cat(sc), ns)


def test_c_oracle_matches_reference_golden():
    """The plain-C restatement (oracle/lattice_oracle.c) against the reference's own
    compute_beta_per_sample outputs."""
    from oracle import c_oracle

    g = _load("beta_per_sample.npz")
    theta = g["theta0"]
    tables = [g[f"tr_{i}"] for i in range(int(g["n_cases"]))]
    batch = _batch_from_tables(tables, theta)
    logz, alpha, beta, post = c_oracle.forward_backward(batch)
    for i, tr in enumerate(tables):
        ref = g[f"beta0_{i}"]
        sl = slice(batch.state_off[i], batch.state_off[i + 1])
        # scores reach the C oracle as float32 (what the GPU path is fed)
        np.testing.assert_allclose(np.exp(beta[sl]), ref, rtol=2e-6)
        assert abs(logz[i] - np.log(ref[0])) < 1e-5


def test_c_oracle_matches_numpy_oracle():
    from oracle import c_oracle
    from tests.lattice_gen import random_mark_lattice

    rng = np.random.default_rng(9)
    tables = [random_mark_lattice(rng, int(n), 24, parallel_arcs=bool(i % 2))[1] for i, n in enumerate(rng.integers(1, 30, size=12))]
    theta = rng.integers(-2, 1, size=24).astype(np.float64)  # exact in fp32; many Viterbi ties
    batch = _batch_from_tables(tables, theta)
    logz, alpha, beta, post = c_oracle.forward_backward(batch)
    score, paths, labels = c_oracle.viterbi(batch)
    for i, tr in enumerate(tables):
        s, l, d, _ = lo.arcs_from_dense(tr)
        lz, al, be, po = lo.forward_backward(tr.shape[0], s, d, theta[l])
        sl = slice(batch.state_off[i], batch.state_off[i + 1])
        al_c = slice(batch.arc_off[i], batch.arc_off[i + 1])
        assert abs(lz - logz[i]) < 1e-12
        np.testing.assert_allclose(alpha[sl], al, atol=1e-12)
        np.testing.assert_allclose(beta[sl], be, atol=1e-12)
        np.testing.assert_allclose(post[al_c], po, atol=1e-12)
        vs, vp, vl, _, _ = lo.viterbi_f32(tr.shape[0], s, l, d, theta[l])
        assert float(vs) == float(score[i])
        assert list(vl) == list(labels[i])
        np.testing.assert_array_equal(paths[i] - batch.arc_off[i], vp)
