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
