"""``nfst_b200.data.get_state_mask_pynini`` / ``pack_fsts`` against the UNMODIFIED reference's
``FSAGRUScorer.get_state_mask_pynini`` (``scorers.py:995-1035``, SURVEY.md section 8 row a1: the function that defines the
lattice tables).  ``tests/golden/state_mask.npz`` holds the reference's own output on eight OpenFst-shaped stand-ins
(``tests/golden/make_golden.py:gen_state_mask``); the machines are rebuilt here from the arrays stored beside it."""
import os

import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import data as nd
from oracle import lattice_oracle as lo
from tests.lattice_gen import PAD, FakeFst, FakeWeight, random_fst_arrays
from tests.test_pack import _np, check_structure, replay_beta

G = os.path.join(os.path.dirname(__file__), "golden", "state_mask.npz")
ZERO = FakeWeight.zero("tropical")


def _cases():
    g = np.load(G)
    V = int(g["vocab"])
    for i in range(int(g["n_cases"])):
        arrs = tuple(g[f"{k}_{i}"] for k in ("n", "src", "ilabel", "nextstate", "weight", "is_final"))
        yield V, arrs, g[f"emission_{i}"], g[f"transition_{i}"]


def test_tables_equal_the_reference_output_bit_for_bit():
    for V, arrs, em_ref, tr_ref in _cases():
        em, tr = nd.get_state_mask_pynini(FakeFst(*arrs), V, PAD, to_numpy=True, final_zero=ZERO)
        assert em.dtype == em_ref.dtype == np.bool_ and tr.dtype == tr_ref.dtype == np.int64
        assert em.shape == em_ref.shape and np.array_equal(em, em_ref) and np.array_equal(tr, tr_ref)
        # default return type: torch tensors (scorers.py:1035)
        em_t, tr_t = nd.get_state_mask_pynini(FakeFst(*arrs), V, PAD, final_zero=ZERO)
        assert isinstance(em_t, torch.Tensor) and em_t.dtype == torch.bool and torch.equal(tr_t, torch.from_numpy(tr_ref))


def test_weighted_tables_hold_minus_the_arc_weight():
    for V, arrs, em_ref, tr_ref in _cases():
        n, src, lab, nxt, w, fin = arrs
        em, tr = nd.get_state_mask_pynini(FakeFst(*arrs), V, PAD, to_numpy=True, weighted=True, final_zero=ZERO)
        assert em.dtype == np.float64 and np.array_equal(tr, tr_ref)
        assert np.array_equal(np.isfinite(em), em_ref)  # -inf exactly where the bool table is False (scorers.py:1011-1013)
        assert em[int(n), PAD] == 0.0
        np.testing.assert_array_equal(em[src, lab], -w)  # scorers.py:1026-1027


def test_reference_assertions_are_kept():
    rng = np.random.default_rng(1)
    n, src, lab, nxt, w, fin = random_fst_arrays(rng, 6, 24)

    class NotAtZero(FakeFst):
        def start(self):
            return 1

    with pytest.raises(AssertionError):  # scorers.py:1005
        nd.get_state_mask_pynini(NotAtZero(n, src, lab, nxt, w, fin), 24, PAD, final_zero=ZERO)
    dup = (n, np.append(src, src[1]), np.append(lab, lab[1]), np.append(nxt, nxt[1]), np.append(w, 0.0), fin)
    with pytest.raises(AssertionError):  # scorers.py:1030: one arc per (state, label)
        nd.get_state_mask_pynini(FakeFst(*dup), 24, PAD, final_zero=ZERO)
    with pytest.raises(ValueError, match="label out of range"):
        nd.pack_fsts([FakeFst(n, src, lab, nxt, w, fin)], 8, final_zero=ZERO)


def test_pack_fsts_equals_pack_dense_of_the_collated_reference_tables():
    cases = list(_cases())
    V = cases[0][0]
    machines = [FakeFst(*arrs) for _, arrs, _, _ in cases]
    tabs = [tr for _, _, _, tr in cases]
    direct = nd.pack_fsts(machines, V, final_zero=ZERO)
    dense = nb.pack_dense(None, torch.from_numpy(lo.collate_pad(tabs, PAD)))  # what collate() + the DP's scan would see
    check_structure(direct)
    assert (direct.n_lattices, direct.n_states, direct.n_arcs) == (dense.n_lattices, dense.n_states, dense.n_arcs)
    for f in ("state_off", "level_off", "level_ptr", "out_ptr", "dst_out", "label_out", "orig_state", "start_state", "sinks"):
        assert torch.equal(getattr(direct, f), getattr(dense, f)), f
    # weighted: the static arc scores are -arc.weight of the kept arcs, and the DP over them matches the oracle
    wdirect = nd.pack_fsts(machines, V, weighted=True, final_zero=ZERO)
    em_w = [nd.get_state_mask_pynini(m, V, PAD, to_numpy=True, weighted=True, final_zero=ZERO)[0] for m in machines]
    w_out = wdirect.static_scores.numpy().astype(np.float64)
    beta = replay_beta(wdirect, w_out)
    for b, (tr, em) in enumerate(zip(tabs, em_w)):
        s, l, d, sc = lo.arcs_from_dense(tr, em)
        logz = lo.forward_backward(tr.shape[0], s, d, sc.astype(np.float32).astype(np.float64))[0]
        assert abs(beta[_np(wdirect.start_state)[b]] - logz) < 1e-9


def test_packed_to_dense_is_the_inverse_of_pack_dense():
    """``data.packed_to_dense``: a packed batch back to the tables ``collate`` hands the reference's modules -- the same
    lattices (arc for arc under the DP's edge rule), sinks with their pad self-loop, padding rows full of the pad id."""
    cases = list(_cases())
    V = cases[0][0]
    machines = [FakeFst(*arrs) for _, arrs, _, _ in cases]
    for weighted in (False, True):
        p = nd.pack_fsts(machines, V, weighted=weighted, final_zero=ZERO)
        em, tr = nd.packed_to_dense(p, PAD, weighted=weighted)
        B, S, _ = tr.shape
        state_off = _np(p.state_off)
        assert B == p.n_lattices and S == int(np.diff(state_off).max())
        assert em.dtype == (torch.float64 if weighted else torch.bool)
        for b in range(B):
            n = state_off[b + 1] - state_off[b]
            t = tr[b, :n].numpy()
            s, l, d, sc = lo.arcs_from_dense(t, em[b, :n].numpy())
            a0, a1 = _np(p.arc_off)[b], _np(p.arc_off)[b + 1]
            out_ptr = _np(p.out_ptr)
            want_src = np.repeat(np.arange(p.n_states), np.diff(out_ptr[: p.n_states + 1]))[a0:a1] - state_off[b]
            assert np.array_equal(s, want_src) and np.array_equal(l, _np(p.label_out)[a0:a1])
            assert np.array_equal(d, _np(p.dst_out)[a0:a1] - state_off[b])
            if weighted:
                np.testing.assert_array_equal(sc.astype(np.float32), p.static_scores.numpy()[a0:a1])
            assert not np.any(d == 0)  # nothing enters the start row
            sink_rows = np.nonzero(np.diff(out_ptr[state_off[b]: state_off[b + 1] + 1]) == 0)[0]
            assert np.array_equal(t[sink_rows, PAD], sink_rows)  # the pad self-loop (scorers.py:1013-1016)
            assert bool((tr[b, n:] == PAD).all()) and bool((em[b, n:] == (float(PAD) if weighted else True)).all())
        # and back again: the same packed batch
        q = nb.pack_dense(em, tr, weighted=weighted)
        for f in ("state_off", "level_off", "level_ptr", "out_ptr", "dst_out", "label_out", "start_state", "sinks"):
            assert torch.equal(getattr(p, f), getattr(q, f)), f
        if weighted:
            assert torch.equal(p.static_scores, q.static_scores)
    dup = nb.pack_arcs(torch.zeros(2, dtype=torch.int64), torch.tensor([0, 0]), torch.tensor([1, 2]), torch.tensor([5, 5]),
                       torch.tensor([3]), V)
    with pytest.raises(ValueError, match="share a label"):
        nd.packed_to_dense(dup, PAD)
