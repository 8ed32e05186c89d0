"""Sliced-column layout (nfst_sell.cu), host side: the packer's column-major slices are replayed
in numpy with the SAME addressing the kernels use (slice base from out_ptr, one degree byte per
state, column offsets from the degrees) and compared with the float64 oracle -- beta / logZ,
posteriors through the conditional-probability flow, Viterbi with first-label ties."""
import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import synth
from nfst_b200.pack import concat_packed, pack_arcs
from oracle import lattice_oracle as lo


@pytest.fixture(autouse=True)
def narrow_lattices_too(monkeypatch):
    """The packer only sends lattices of >= 96 states per level down the sliced-column path (below that the
    CSR kernels are faster); the tests want small lattices there too."""
    monkeypatch.setattr(nb.pack, "SELL_MIN_WIDTH", 32)
    monkeypatch.setattr(nb.tiles, "TILES", 0)  # these tests are about the sliced-column layout (tiles win by default)


def _np(t):
    return t.cpu().numpy().astype(np.int64)


def slices_of(p, b):
    """[(level, first_state, n_states_in_slice)] of lattice b, in level order."""
    level_off, level_ptr = _np(p.level_off), _np(p.level_ptr)
    lp = level_ptr[level_off[b]: level_off[b + 1]]
    out = []
    for l in range(len(lp) - 1):
        for first in range(lp[l], lp[l + 1], 32):
            out.append((l, first, min(32, lp[l + 1] - first)))
    return out


def slice_columns(p, first, n):
    """arc ids of a slice as the kernels address them: {state: [arc of column 0, column 1, ...]}"""
    deg = _np(p.out_deg8)[first:first + n]
    true_deg = np.diff(_np(p.out_ptr)[first:first + n + 1])
    deg = np.where(deg == 255, true_deg, deg)  # the byte saturates; 255 = look at out_ptr
    off = int(_np(p.out_ptr)[first])
    arcs = {first + i: [] for i in range(n)}
    k = 0
    while (deg > k).any():
        on = np.nonzero(deg > k)[0]
        assert np.array_equal(on, np.arange(len(on))), "degrees must descend inside a slice"
        for rank, i in enumerate(on):
            arcs[first + i].append(off + rank)
        off += len(on)
        k += 1
    return arcs


def replay_sell(p, w, b):
    """(beta, cond, post, delta, backptr) of lattice b replayed over the sliced-column arrays."""
    dst, S = _np(p.dst_out), p.n_states
    beta = np.full(S, np.nan)
    delta = np.full(S, np.nan, dtype=np.float32)
    bp = np.full(S, -2)
    cond = np.zeros(p.n_arcs)
    sl = slices_of(p, b)
    for (_, first, n) in reversed(sl):
        for s, arcs in slice_columns(p, first, n).items():
            if not arcs:
                beta[s], delta[s], bp[s] = 0.0, 0.0, -1
                continue
            t = w[arcs].astype(np.float64) + beta[dst[arcs]]
            beta[s] = lo._lse(t)
            cond[arcs] = np.exp(t - beta[s])
            c = (w[arcs].astype(np.float32) + delta[dst[arcs]]).astype(np.float32)
            j = int(np.argmax(c))  # first maximum = smallest label
            delta[s], bp[s] = c[j], arcs[j]
    gamma = np.zeros(S)
    gamma[int(p.start_state[b])] = 1.0
    post = np.zeros(p.n_arcs)
    for (_, first, n) in sl:
        for s, arcs in slice_columns(p, first, n).items():
            for a in arcs:
                post[a] = gamma[s] * cond[a]
                gamma[dst[a]] += post[a]
    return beta, cond, post, delta, bp


def layered(B, arcs, levels, seed):
    return synth.random_dag_batch(B, arcs, levels=levels, seed=seed)


@pytest.mark.parametrize("arcs,levels,seed", [(1500, 6, 0), (4000, 10, 1), (900, 3, 2)])
def test_sell_layout_replays_to_the_oracle(arcs, levels, seed):
    ab = layered(3, arcs, levels, seed)
    p, w = ab.pack(sell=True)
    assert p.has_sell and all(g.sell for g in p.groups)
    w = w.numpy()
    src_out, dst, lab = _np(p.src_out), _np(p.dst_out), _np(p.label_out)
    # canonical arrays are a permutation of the input arcs
    origin = _np(p.arc_origin)
    assert sorted(origin.tolist()) == list(range(ab.src.numel()))
    np.testing.assert_array_equal(lab, ab.label.numpy()[origin])
    out_ptr = _np(p.out_ptr)
    np.testing.assert_array_equal(np.bincount(src_out, minlength=p.n_states), np.diff(out_ptr))
    np.testing.assert_array_equal(_np(p.out_deg8), np.minimum(np.diff(out_ptr), 255))
    state_off, arc_off = _np(p.state_off), _np(p.arc_off)
    for b in range(p.n_lattices):
        a0, a1 = arc_off[b], arc_off[b + 1]
        for (_, first, n) in slices_of(p, b):
            cols = slice_columns(p, first, n)
            for s, arcs in cols.items():
                assert all(src_out[a] == s for a in arcs)
                labs = [lab[a] for a in arcs]
                assert labs == sorted(labs), "columns follow label order (Viterbi tie-break)"
        # ring bound: every arc's destination level ends within the window of its source level
        g = [g for g in p.groups if b in g.ids.tolist()][0]
        level_off, level_ptr = _np(p.level_off), _np(p.level_ptr)
        lp = level_ptr[level_off[b]: level_off[b + 1]]
        lev = np.searchsorted(lp, np.arange(state_off[b], state_off[b + 1]), side="right") - 1
        la = np.arange(a0, a1)
        span = lp[lev[dst[la] - state_off[b]] + 1] - lp[lev[src_out[la] - state_off[b]]]
        # the ring covers nearly all arcs; sell_far announces the rest (they go through global memory)
        assert g.sell_window & (g.sell_window - 1) == 0 and (span <= g.sell_window).mean() >= 0.99
        assert g.sell_far or span.max() <= g.sell_window
        assert np.diff(lp).max() <= g.sell_window, "two states of a level must not share a ring slot"

        beta, cond, post, delta, bp = replay_sell(p, w, b)
        ns = state_off[b + 1] - state_off[b]
        ls, ld = src_out[la] - state_off[b], dst[la] - state_off[b]
        start = int(p.start_state[b]) - state_off[b]
        o_logz, o_alpha, o_beta, o_post = lo.forward_backward(ns, ls, ld, w[la].astype(np.float64), start=start)
        np.testing.assert_allclose(beta[state_off[b]:state_off[b + 1]], o_beta, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(post[la], o_post, rtol=1e-9, atol=1e-300)
        v_score, v_arcs, _, _, _ = lo.viterbi_f32(ns, ls, lab[la], ld, w[la], start=start)
        assert np.float32(delta[int(p.start_state[b])]) == np.float32(v_score)
        path, s = [], int(p.start_state[b])
        while bp[s] >= 0:
            path.append(bp[s] - a0)
            s = dst[bp[s]]
        assert path == list(v_arcs)


def test_sell_is_per_lattice_and_survives_concat():
    wide = layered(2, 2000, 5, 5)
    narrow = synth.transliteration_batch(3, seed=1)
    pw, _ = wide.pack()
    pn, _ = narrow.pack()
    assert pw.has_sell and not pn.has_sell
    both = concat_packed([pn, pw])
    kinds = sorted((g.sell, g.n) for g in both.groups)
    assert (True, 2) in kinds and sum(n for s, n in kinds if not s) == 3
    # the wide lattices keep their column-major arcs, shifted by the narrow part
    np.testing.assert_array_equal(_np(both.dst_out)[pn.n_arcs:], _np(pw.dst_out) + pn.n_states)
    np.testing.assert_array_equal(_np(both.src_out)[pn.n_arcs:], _np(pw.src_out) + pn.n_states)
    np.testing.assert_array_equal(_np(both.out_deg8), np.concatenate([_np(pn.out_deg8), _np(pw.out_deg8)]))
    # opting out gives plain CSR for everything
    pc, _ = wide.pack(sell=False)
    assert not pc.has_sell
    src_csr = np.repeat(np.arange(pc.n_states), np.diff(_np(pc.out_ptr)))
    np.testing.assert_array_equal(_np(pc.src_out), src_csr)


def test_sell_heavy_state_and_narrow_lattices():
    # a star: one state with 300 arcs (saturates the degree byte) over a wide level -> still sliced-column
    n = 300
    src = torch.cat([torch.zeros(n, dtype=torch.int64), torch.arange(1, n + 1)])
    dst = torch.cat([torch.arange(1, n + 1), torch.full((n,), n + 1)])
    lab = torch.cat([torch.arange(n) % 200 + 4, torch.full((n,), 3)])
    p = pack_arcs(torch.zeros(2 * n, dtype=torch.int64), src, dst, lab, torch.tensor([n + 2]), 256, sell=True)
    assert p.has_sell and int(p.out_deg8[int(p.start_state[0])]) == 255
    cols = slice_columns(p, int(p.start_state[0]), 1)
    assert cols[int(p.start_state[0])] == list(range(n))  # a lone state's arcs are contiguous
    # narrow lattices (the reference's own shapes) stay CSR
    pn, _ = synth.transliteration_batch(2, seed=3).pack(sell=True)
    assert not pn.has_sell


def test_sell_default_thresholds(monkeypatch):
    monkeypatch.undo()  # the packer's own defaults ...
    monkeypatch.setattr(nb.tiles, "TILES", 0)  # ... of the sliced-column path
    assert not layered(2, 2000, 10, 1).pack()[0].has_sell  # 50 states per level: CSR kernels
    p, _ = layered(1, 40_000, 16, 1).pack()  # 625 states per level
    assert p.has_sell and p.groups[0].block_threads == 128
    monkeypatch.setattr(nb.pack, "SELL_WINDOW_MAX", 1024)  # ring would not fit: stays CSR
    assert not layered(1, 40_000, 16, 1).pack()[0].has_sell
