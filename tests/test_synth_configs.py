"""The synthetic workloads are the ones SURVEY.md section 8(d) / BASELINE.json name -- shapes, seeds and determinism -- so
that the bench point and the config tests cannot drift without a test noticing (small batches of each generator; the
full sizes run on the GPU in tests/test_gpu_configs.py)."""
import numpy as np
import torch

from nfst_b200 import synth
from oracle import lattice_oracle as lo


def _per_lattice(ab, b):
    sel = (ab.arc_lattice == b).numpy()
    return int(ab.n_states[b]), ab.src.numpy()[sel], ab.dst.numpy()[sel], ab.label.numpy()[sel], ab.scores.numpy()[sel]


def _levels(n, s, d):
    """longest distance from state 0 (-1 = unreachable)"""
    lvl = np.full(n, -1)
    lvl[0] = 0
    for _ in range(n):
        cand = np.where(lvl[s] >= 0, lvl[s] + 1, -1)
        new = lvl.copy()
        np.maximum.at(new, d, cand)
        if np.array_equal(new, lvl):
            break
        lvl = new
    return lvl


def test_config4_random_dag_point():
    """A arcs per lattice, S = A/4 states in 64 levels, in-degree 1 + Poisson(3) from the previous 1-3 levels, single
    source and sink, scores U(-1, 0), seed 3."""
    for A in (10_000, 30_000):
        ab = synth.random_dag_batch(3, A, levels=64, seed=3)
        again = synth.random_dag_batch(3, A, levels=64, seed=3)
        assert all(torch.equal(getattr(ab, f), getattr(again, f)) for f in ("arc_lattice", "src", "dst", "label", "scores"))
        for b in range(3):
            n, s, d, l, w = _per_lattice(ab, b)
            assert abs(n - A / 4) <= 64 and abs(len(s) - A) < 0.06 * A, (n, len(s))
            assert lo.is_acyclic(n, s, d)
            lvl = _levels(n, s, d)
            assert lvl.min() >= 0 and lvl.max() == 63  # every state reachable, 64 levels
            assert np.count_nonzero(lvl == 0) == 1 and np.count_nonzero(lvl == 63) == 1
            out_deg = np.bincount(s, minlength=n)
            assert np.array_equal(np.nonzero(out_deg == 0)[0], [n - 1])  # one sink, no dead end left
            width = (n - 2) // 62
            layer = np.concatenate([[0], 1 + np.arange(n - 2) // width, [63]])  # the layout's own levels
            span = layer[d] - layer[s]
            inner = d != n - 1
            assert span[inner].min() >= 1 and span[inner].max() <= 3  # predecessors from the previous 1-3 levels
            in_deg = np.bincount(d[inner], minlength=n)[1:n - 1]
            deep = layer[1:n - 1] >= 2
            assert abs(in_deg[deep].mean() - 4.0) < 0.15 and in_deg.min() >= 1  # 1 + Poisson(3)
            assert w.dtype == np.float32 and -1.0 <= w.min() and w.max() <= 0.0 and abs(w.mean() + 0.5) < 0.02
            assert l.min() >= 8 and l.max() < 256


def test_config1_and_config5_transliteration_lattices():
    """|x|, |y| ~ U{4..12}, per grid cell delete = 2 arcs, insert = 2 arcs, substitute = 5 arcs, bos / eos, V = 256;
    config 5's second score set is integer-valued in {-2, -1, 0}."""
    ab = synth.transliteration_batch(32, seed=0)
    assert ab.vocab == 256 and int(ab.n_states.numel()) == 32
    sizes = set()
    for b in range(32):
        n, s, d, l, w = _per_lattice(ab, b)
        # invert S = 1 + (n+1)(m+1) + n(m+1) + (n+1)m + 4nm + 1 and A = 2 + 2n(m+1) + 2(n+1)m + 5nm over 4..12
        hit = [(x, y) for x in range(4, 13) for y in range(4, 13)
               if 2 + (x + 1) * (y + 1) + x * (y + 1) + (x + 1) * y + 4 * x * y == n
               and 2 + 2 * x * (y + 1) + 2 * (x + 1) * y + 5 * x * y == len(s)]
        assert hit, (n, len(s))
        sizes.add(hit[0])
        assert lo.is_acyclic(n, s, d) and len(set(zip(s.tolist(), l.tolist()))) == len(s)  # deterministic
    assert len(sizes) > 10
    ints = synth.transliteration_batch(64, seed=4, integer_scores=True)
    assert set(np.unique(ints.scores.numpy()).tolist()) <= {-2.0, -1.0, 0.0}


def test_config2_snips_and_config3_cipher_shapes():
    ab = synth.snips_batch(8, seed=1)
    for b in range(8):
        n, s, d, l, w = _per_lattice(ab, b)
        assert lo.is_acyclic(n, s, d) and len(set(zip(s.tolist(), l.tolist()))) == len(s)
        assert np.count_nonzero((l >= 140) & (l < 147)) % 7 == 0  # the 7-way intent fan-out
    T = 50
    uni = synth.cipher_batch(2, T=T, bigram=False, seed=2)
    n, s, d, l, w = _per_lattice(uni, 0)
    assert n == T + 3 and len(s) == 26 * T + 2  # 26 parallel arcs per position: depth T, width 1
    bi = synth.cipher_batch(2, T=T, bigram=True, seed=2)
    n, s, d, l, w = _per_lattice(bi, 0)
    assert n == 26 * T + 4 and len(s) == 1 + 26 + 676 * (T - 1) + 26 + 1
    # arc scores are log-probabilities of a stochastic channel x language model: every column of the trellis sums to
    # a probability, so logZ <= 0
    logz = lo.forward_backward(n, s, d, w.astype(np.float64))[0]
    assert logz < 0.0
