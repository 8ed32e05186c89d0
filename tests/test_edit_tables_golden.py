"""The unmodified reference's DP run on lattices THIS repo constructs: ``tests/golden/edit_tables.npz`` holds, for ten
edit lattices (``make_golden.py:gen_edit_tables``), the dense tables the product writes for them (host construction ->
``pack_arcs`` -> ``data.packed_to_dense``) and the ``compute_beta_per_sample`` outputs of the reference on exactly those
tables (float64, ``Wh = 0`` and ``Wh != 0``, ``scorers.py:692-751``).  Here: the product still writes the same tables, and
the oracles reproduce the reference's beta on them -- construction (row f-4), table format (a1) and the recurrence
(a5) pinned end to end."""
import os

import numpy as np
import torch

import nfst_b200 as nb
from nfst_b200 import data as nd
from oracle import edit_lattice_oracle as elo
from oracle import lattice_oracle as lo
from tests.lattice_gen import PAD

G = os.path.join(os.path.dirname(__file__), "golden", "edit_tables.npz")
# the pairs of make_golden.py:EDIT_PAIRS, without and then with the substitution arcs
EDIT_PAIRS = [([7, 8, 9], [10, 11]), ([7], [12, 12, 13]), ([8, 9, 7, 7], [10, 13, 11, 12]), ([], [10]), ([9, 8], [])]


def _tables(x, y, sub, V):
    arcs, n = elo.edit_lattice(x, y, bos=1, eos=2, input_mark=4, output_mark=5, sub_mark=6 if sub else None)
    src, lab, dst = (torch.tensor([a[i] for a in arcs]) for i in (0, 1, 2))
    p = nb.pack_arcs(torch.zeros(len(arcs), dtype=torch.int64), src, dst, lab, torch.tensor([n]), V)
    return nd.packed_to_dense(p, PAD), p


def test_reference_beta_on_product_built_tables():
    g = np.load(G)
    V = int(g["vocab"])
    P0 = {k: g["p0_" + k] for k in ("emb", "Wx", "Wh", "W", "bias")}
    P1 = {k: g["p1_" + k] for k in ("emb", "Wx", "Wh", "W", "bias")}
    assert not P0["Wh"].any() and P1["Wh"].any()
    theta = (P0["W"] @ np.tanh(P0["Wx"] @ P0["emb"].T + P0["bias"][:, None]))[0]  # scorers.py:732-738 with Wh = 0
    i = 0
    for sub in (False, True):
        for x, y in EDIT_PAIRS:
            (em, tr), p = _tables(x, y, sub, V)
            t = tr[0].numpy()
            assert np.array_equal(t, g[f"tr_{i}"]), (x, y, sub)  # the product writes today what the reference was run on
            src, lab, dst, _ = lo.arcs_from_dense(t)
            beta = lo.beta_log(t.shape[0], src, dst, theta[lab])
            np.testing.assert_allclose(np.exp(beta), g[f"beta0_{i}"], rtol=1e-12, atol=0)
            bh, _ = lo.beta_recurrent(t.shape[0], src, lab, dst, P1["emb"], P1["Wx"], P1["Wh"], P1["W"], P1["bias"])
            np.testing.assert_allclose(bh, g[f"beta1_{i}"], rtol=1e-10, atol=0)
            # logZ counts the alignments when every arc scores 0 (Delannoy / binomial numbers)
            z0 = lo.beta_log(t.shape[0], src, dst, np.zeros(len(src)))[0]
            n_paths = len(elo.mark_strings(x, y, bos=1, eos=2, input_mark=4, output_mark=5, sub_mark=6 if sub else None))
            assert abs(np.exp(z0) - n_paths) < 1e-6 * n_paths
            i += 1
    assert i == int(g["n_cases"])


def test_label_scores_is_the_reference_message_weight_at_wh_zero():
    """``scorer.label_scores`` (what ``patch_compute_beta`` feeds the kernels when ``Wh = 0``): theta[l] = W.tanh(Wx e_l + b),
    ``scorers.py:732-738`` -- with it the oracle reproduces the reference's beta of the golden above."""
    from nfst_b200.scorer import label_scores

    g = np.load(G)
    P0 = {k: torch.from_numpy(g["p0_" + k]) for k in ("emb", "Wx", "W", "bias")}
    theta = label_scores(P0["emb"], P0["Wx"], P0["W"], P0["bias"])
    assert theta.dtype == torch.float32 and theta.shape == (int(g["vocab"]),)
    t = g["tr_0"]
    src, lab, dst, _ = lo.arcs_from_dense(t)
    beta = lo.beta_log(t.shape[0], src, dst, theta.double().numpy()[lab])
    np.testing.assert_allclose(np.exp(beta), g["beta0_0"], rtol=2e-6, atol=0)  # float32 theta
