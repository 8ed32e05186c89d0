"""Host-side packing (CPU tensors): structure invariants and equivalence with the oracle.

The DP is replayed over the *packed* arrays with plain numpy here, so that a packing bug
shows up without a GPU; the CUDA kernels are checked in the -m gpu tests.
"""
import numpy as np
import pytest
import torch

import nfst_b200 as nb
from oracle import lattice_oracle as lo
from tests.lattice_gen import PAD, chain_with_skips, random_mark_lattice


def _np(t):
    return t.cpu().numpy().astype(np.int64)


def replay_beta(p, w_out):
    """level-synchronous log-semiring backward over the packed arrays (float64)."""
    beta = np.full(p.n_states, -np.inf)
    out_ptr, dst = _np(p.out_ptr), _np(p.dst_out)
    level_off, level_ptr = _np(p.level_off), _np(p.level_ptr)
    for b in range(p.n_lattices):
        lp = level_ptr[level_off[b] : level_off[b + 1]]
        for l in range(len(lp) - 2, -1, -1):
            for s in range(lp[l], lp[l + 1]):
                a = np.arange(out_ptr[s], out_ptr[s + 1])
                beta[s] = 0.0 if len(a) == 0 else lo._lse(w_out[a] + beta[dst[a]])
    return beta


def replay_alpha(p, w_out):
    alpha = np.full(p.n_states, -np.inf)
    in_ptr, src, in2out = _np(p.in_ptr), _np(p.src_in), _np(p.in2out)
    level_off, level_ptr, start = _np(p.level_off), _np(p.level_ptr), _np(p.start_state)
    for b in range(p.n_lattices):
        lp = level_ptr[level_off[b] : level_off[b + 1]]
        alpha[start[b]] = 0.0
        for l in range(1, len(lp) - 1):
            for s in range(lp[l], lp[l + 1]):
                a = np.arange(in_ptr[s], in_ptr[s + 1])
                alpha[s] = lo._lse(w_out[in2out[a]] + alpha[src[a]])
    return alpha


def check_structure(p):
    out_ptr, in_ptr = _np(p.out_ptr), _np(p.in_ptr)
    dst, src_in, in2out = _np(p.dst_out), _np(p.src_in), _np(p.in2out)
    lab_out, lab_in = _np(p.label_out), _np(p.label_in)
    state_off, level_off, level_ptr = _np(p.state_off), _np(p.level_off), _np(p.level_ptr)
    S, A = p.n_states, p.n_arcs
    assert out_ptr[0] == 0 and out_ptr[-1] == A and in_ptr[-1] == A
    assert sorted(in2out.tolist()) == list(range(A))
    src_out = np.repeat(np.arange(S), np.diff(out_ptr))
    np.testing.assert_array_equal(src_in, src_out[in2out])
    np.testing.assert_array_equal(lab_in, lab_out[in2out])
    dst_in = np.repeat(np.arange(S), np.diff(in_ptr))
    np.testing.assert_array_equal(dst_in, dst[in2out])
    # level of each state; arcs go strictly upward; levels are contiguous ranges
    level = np.full(S, -1)
    for b in range(p.n_lattices):
        lp = level_ptr[level_off[b] : level_off[b + 1]]
        assert lp[0] == state_off[b] and lp[-1] == state_off[b + 1]
        assert np.all(np.diff(lp) > 0)  # no empty level
        for l in range(len(lp) - 1):
            level[lp[l] : lp[l + 1]] = l
        assert _np(p.start_state)[b] == lp[0] and lp[1] - lp[0] == 1
    assert np.all(level[dst] > level[src_out])
    # canonical order: by source state, then label
    key = src_out * p.vocab + lab_out
    assert np.all(np.diff(key) >= 0)
    # sinks = states without outgoing arcs
    sinks = _np(p.sinks)
    np.testing.assert_array_equal(sinks, np.nonzero(np.diff(out_ptr) == 0)[0])
    # launch groups partition the batch
    ids = np.concatenate([_np(g.ids) for g in p.groups])
    assert sorted(ids.tolist()) == list(range(p.n_lattices))
    capb = np.zeros(p.n_lattices, dtype=np.int64)
    for g in p.groups:
        capb[_np(g.ids)] = g.chunk_cap
        assert g.block_threads in (32, 64, 128, 256) and g.chunk_cap % 8 == 0
    # chunks: each direction tiles the states of a lattice exactly once, level by level
    for off, chunks, ptr, desc in ((p.fwd_chunk_off, p.fwd_chunks, in_ptr, False),
                                   (p.bwd_chunk_off, p.bwd_chunks, out_ptr, True)):
        off, chunks = _np(off), _np(chunks)
        assert off[0] == 0 and off[-1] == len(chunks)
        for b in range(p.n_lattices):
            ck = chunks[off[b] : off[b + 1]]
            if desc:
                ck = ck[::-1]
            assert ck[0, 2] == state_off[b] and ck[-1, 3] == state_off[b + 1]
            np.testing.assert_array_equal(ck[1:, 2], ck[:-1, 3])  # contiguous in state space
            np.testing.assert_array_equal(ck[:, 0], ptr[ck[:, 2]])
            np.testing.assert_array_equal(ck[:, 1], ptr[ck[:, 3]])
            assert np.all(level[ck[:, 2]] == level[ck[:, 3] - 1])  # never crosses a level
            big = (ck[:, 1] - ck[:, 0]) > capb[b]
            assert np.all((ck[big, 3] - ck[big, 2]) == 1)  # only a single heavy state may exceed a stage
            assert np.all((ck[:, 3] - ck[:, 2]) <= capb[b])  # states of a chunk fit the row-pointer stage


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pack_dense_matches_oracle(seed):
    rng = np.random.default_rng(seed)
    tabs = [random_mark_lattice(rng, int(n), 24, parallel_arcs=bool(i % 2))[1] for i, n in enumerate(rng.integers(1, 20, size=5))]
    tr = lo.collate_pad(tabs, PAD)  # padded with the pad id, quirk Q5
    theta = rng.normal(size=24)
    p = nb.pack_dense(None, torch.from_numpy(tr))
    check_structure(p)
    assert p.dense_shape == tr.shape
    w_out = theta[_np(p.label_out)]
    beta = replay_beta(p, w_out)
    alpha = replay_alpha(p, w_out)
    state_off, orig = _np(p.state_off), _np(p.orig_state)
    B, S, V = tr.shape
    for b in range(B):
        nb_states = tabs[b].shape[0]
        src, lab, dst, _ = lo.arcs_from_dense(tabs[b])
        logz, al, be, post = lo.forward_backward(nb_states, src, dst, theta[lab])
        sl = slice(state_off[b], state_off[b + 1])
        # every real state of a well-formed lattice is reachable, padded rows are trimmed
        assert sorted(orig[sl].tolist()) == list(range(nb_states))
        np.testing.assert_allclose(beta[sl], be[orig[sl]], atol=1e-12)
        np.testing.assert_allclose(alpha[sl], al[orig[sl]], atol=1e-12)
        # arc_origin points at the dense cell of every canonical arc
        a0, a1 = _np(p.arc_off)[b], _np(p.arc_off)[b + 1]
        cells = _np(p.arc_origin)[a0:a1]
        bb, ss, ll = np.unravel_index(cells, (B, S, V))
        assert np.all(bb == b)
        np.testing.assert_array_equal(ll, _np(p.label_out)[a0:a1])
        np.testing.assert_array_equal(tr[bb, ss, ll], orig[_np(p.dst_out)[a0:a1]])


def test_pack_weighted_tables_carry_static_scores():
    rng = np.random.default_rng(3)
    _, tr = random_mark_lattice(rng, 7, 24)
    em = np.where(tr != 0, rng.normal(size=tr.shape), -np.inf)
    p = nb.pack_dense(torch.from_numpy(em)[None], torch.from_numpy(tr)[None])
    assert p.static_scores is not None
    cells = _np(p.arc_origin)
    np.testing.assert_allclose(p.static_scores.numpy(), em.reshape(-1)[cells].astype(np.float32))


def test_pack_rejects_cycles_and_bad_shapes():
    tr = np.zeros((1, 4, 8), dtype=np.int64)
    tr[0, 0, 1] = 1
    tr[0, 1, 5] = 2
    tr[0, 2, 6] = 1  # 1 -> 2 -> 1
    tr[0, 2, 2] = 3
    with pytest.raises(ValueError, match="cyclic"):
        nb.pack_dense(None, torch.from_numpy(tr))
    with pytest.raises(ValueError):
        nb.pack_dense(None, torch.zeros(4, 8, dtype=torch.int64))  # scorers.py:878-879: tables are 3-D


def test_pack_edge_rule_drops_arcs_into_state0_and_self_loops():
    # quirks Q2/Q3: arcs into state 0 and self-loops are not edges
    tr = np.zeros((1, 4, 8), dtype=np.int64)
    tr[0, 0, 1] = 1
    tr[0, 1, 4] = 0  # "arc" into state 0 == no arc
    tr[0, 1, 5] = 1  # self loop
    tr[0, 1, 6] = 2
    tr[0, 2, 2] = 3
    tr[0, 3, 3] = 3  # sink pad loop
    p = nb.pack_dense(None, torch.from_numpy(tr))
    assert p.n_arcs == 3 and p.n_states == 4
    check_structure(p)


def test_pack_deep_chain_and_isolated_start():
    _, tr = chain_with_skips(300, 24)
    p = nb.pack_dense(None, torch.from_numpy(tr)[None])
    check_structure(p)
    assert p.max_levels == 302
    # a lattice whose start state has no arc at all: one state, no arcs, logZ = 0
    tr2 = np.zeros((1, 3, 8), dtype=np.int64)
    p2 = nb.pack_dense(None, torch.from_numpy(tr2))
    assert p2.n_states == 1 and p2.n_arcs == 0


def test_pack_arcs_coo_roundtrip():
    rng = np.random.default_rng(4)
    B = 4
    lat, src, dst, lab, ns = [], [], [], [], []
    for b in range(B):
        _, tr = random_mark_lattice(rng, int(rng.integers(2, 12)), 24)
        s, l, d, _ = lo.arcs_from_dense(tr)
        perm = rng.permutation(len(s))  # arbitrary input order
        lat += [b] * len(s)
        src += s[perm].tolist(); dst += d[perm].tolist(); lab += l[perm].tolist()
        ns.append(tr.shape[0])
    t = lambda x: torch.tensor(x, dtype=torch.int64)
    p = nb.pack_arcs(t(lat), t(src), t(dst), t(lab), t(ns), 24)
    check_structure(p)
    o = _np(p.arc_origin)
    np.testing.assert_array_equal(np.array(lab)[o], _np(p.label_out))


def test_concat_packed_equals_joint_pack():
    from nfst_b200 import synth
    from nfst_b200.pack import concat_packed

    ab = synth.transliteration_batch(6, seed=3)
    joint, _ = ab.pack()
    parts = []
    for b in range(6):
        sel = ab.arc_lattice == b
        one = synth.ArcBatch(torch.zeros(int(sel.sum()), dtype=torch.int64), ab.src[sel], ab.dst[sel], ab.label[sel],
                             ab.scores[sel], ab.n_states[b : b + 1], ab.vocab)
        parts.append(one.pack()[0])
    cat = concat_packed(parts)
    check_structure(cat)
    for f in nb.PackedLattices._INT_FIELDS + ("lanes_in_log2", "lanes_out_log2", "orig_state", "arc_off", "n_levels"):
        assert torch.equal(getattr(cat, f), getattr(joint, f)), f
    assert (cat.n_lattices, cat.n_states, cat.n_arcs) == (joint.n_lattices, joint.n_states, joint.n_arcs)
