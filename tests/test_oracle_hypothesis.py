"""Arbitrary small DAGs (hypothesis) through the three CPU pieces the GPU tests lean on -- the numpy oracle, the C
oracle and the host-side packer -- against explicit path enumeration ("brute-force path
enumeration on small lattices (hypothesis-generated) as a second check").

Unlike ``tests/lattice_gen.py``'s mark lattices these graphs have everything a general arc list may have: states the
start cannot reach (trimmed by the packer), dead ends (a state without outgoing arcs is a sink, beta = 0,
``scorers.py:719-727`` root rule), several sinks, parallel arcs, arcs that skip levels, state ids in no
topological order, an isolated start (logZ = 0: the empty path).
"""
import os

import numpy as np
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import nfst_b200 as nb
from oracle import c_oracle
from oracle import lattice_oracle as lo
from tests.test_pack import _np, check_structure, replay_alpha, replay_beta

V = 6
# NFST_HYP_EXAMPLES=n: n examples per test from a fresh random seed (exploration); default: the fixed 120 / 80
_N = int(os.environ.get("NFST_HYP_EXAMPLES", "0"))
_SET = dict(deadline=None, derandomize=not _N, suppress_health_check=list(HealthCheck))


@st.composite
def dags(draw):
    n = draw(st.integers(1, 7))
    pairs = [(s, d) for s in range(n) for d in range(s + 1, n)]
    arcs = set()
    if pairs:
        for s, d in draw(st.lists(st.sampled_from(pairs), max_size=14)):
            free = [l for l in range(V) if (s, l) not in {(a, b) for a, b, _ in arcs}]  # one arc per (state, label)
            if free:
                arcs.add((s, draw(st.sampled_from(free)), d))
    arcs = sorted(arcs)
    # state ids in no particular order (the start stays 0, as in the reference: scorers.py:1005)
    perm = [0] + draw(st.permutations(list(range(1, n))))
    src = np.array([perm[a[0]] for a in arcs], dtype=np.int64)
    lab = np.array([a[1] for a in arcs], dtype=np.int64)
    dst = np.array([perm[a[2]] for a in arcs], dtype=np.int64)
    ints = draw(st.booleans())  # integer scores force exact Viterbi ties
    elem = st.integers(-2, 0).map(float) if ints else st.floats(-3, 3, allow_nan=False, width=32)
    w = np.array(draw(st.lists(elem, min_size=len(arcs), max_size=len(arcs))), dtype=np.float32)
    return n, src, lab, dst, w


@settings(max_examples=_N or 120, **_SET)
@given(dags())
def test_oracles_agree_with_path_enumeration(g):
    n, src, lab, dst, w = g
    w64 = w.astype(np.float64)
    logz, alpha, beta, post = lo.forward_backward(n, src, dst, w64)
    bz, bpost, paths, scores = lo.brute_force(n, src, dst, w64)
    assert len(paths) >= 1 and abs(logz - bz) < 1e-9
    np.testing.assert_allclose(post, bpost, atol=1e-10)
    # C oracle == numpy oracle (states the start cannot reach carry alpha = -inf and posterior 0 in both)
    ob = c_oracle.Batch(np.zeros(len(src), dtype=np.int64), src, dst, lab, w, [n])
    c_logz, c_alpha, c_beta, c_post = c_oracle.forward_backward(ob)
    assert abs(c_logz[0] - logz) < 1e-9
    np.testing.assert_allclose(c_post, post, atol=1e-10)
    reach = np.isfinite(alpha)
    np.testing.assert_allclose(c_alpha[reach], alpha[reach], atol=1e-9)
    np.testing.assert_allclose(c_beta, beta, atol=1e-9)
    # Viterbi: best score by enumeration, bit-exact; among exactly tied optimal paths the smallest label sequence
    score, path, labels, _, _ = lo.viterbi_f32(n, src, lab, dst, w)
    best = max(float(lo.path_score_f32_backward(w, p)) for p in paths)
    assert float(score) == best
    c_score, c_paths, c_labels = c_oracle.viterbi(ob)
    assert np.float32(c_score[0]).view(np.uint32) == np.float32(score).view(np.uint32)
    assert list(c_labels[0]) == list(labels) and list(c_paths[0]) == list(path)
    if np.all(w == np.round(w)):
        tied = [[int(lab[a]) for a in p] for p in paths if float(lo.path_score_f32_backward(w, p)) == best]
        assert list(labels) == min(tied)


@settings(max_examples=_N or 80, **_SET)
@given(st.lists(dags(), min_size=1, max_size=3))
def test_packer_keeps_the_reachable_lattice_and_replays_to_the_oracle(gs):
    lat = np.concatenate([np.full(len(g[1]), b, dtype=np.int64) for b, g in enumerate(gs)])
    src, lab, dst = (np.concatenate([g[k] for g in gs]) for k in (1, 2, 3))
    w = np.concatenate([g[4] for g in gs]).astype(np.float64)
    ns = [g[0] for g in gs]
    p = nb.pack_arcs(*(torch.from_numpy(x) for x in (lat, src, dst, lab)), torch.tensor(ns), V)
    check_structure(p)
    origin = _np(p.arc_origin)
    w_out = w[origin]
    beta, alpha = replay_beta(p, w_out), replay_alpha(p, w_out)
    state_off, orig, arc_off = _np(p.state_off), _np(p.orig_state), _np(p.arc_off)
    a0 = 0
    for b, (n, s, l, d, wf) in enumerate(gs):
        logz, al, be, post = lo.forward_backward(n, s, d, wf.astype(np.float64))
        sl = slice(state_off[b], state_off[b + 1])
        # exactly the states the start reaches are kept, under their own ids
        assert sorted(orig[sl].tolist()) == np.nonzero(np.isfinite(al))[0].tolist()
        np.testing.assert_allclose(beta[sl], be[orig[sl]], atol=1e-10)
        np.testing.assert_allclose(alpha[sl], al[orig[sl]], atol=1e-10)
        assert abs(beta[_np(p.start_state)[b]] - logz) < 1e-10
        # exactly the arcs that leave a reachable state are kept, in canonical order: packed source state, then label
        kept = origin[arc_off[b]:arc_off[b + 1]] - a0
        new_id = np.full(n, -1)
        new_id[orig[sl]] = np.arange(state_off[b], state_off[b + 1])
        want = [int(a) for a in np.lexsort((l, new_id[s])) if np.isfinite(al[s[a]])]
        assert kept.tolist() == want
        a0 += len(s)
