"""Seeded generators of small lattices in the reference's dense-table format.

Table format (reference ``scorers.py:995-1035``): ``transition[S, V]`` int64 next state
(0 = no arc), ``emission[S, V]`` bool; row 0 is the start state and its only arc is
``bos``; the last row is the absorbing sink whose only arc is ``pad -> itself``; the arc
into the sink carries ``eos``.  Vocabulary ids as shipped: bos=1, eos=2, pad=3.
"""
from __future__ import annotations

import numpy as np

BOS, EOS, PAD = 1, 2, 3
FIRST_LABEL = 7  # 0 eps, 1-3 bos/eos/pad, 4-6 the three mark names (scorers.py:974-976)


def random_mark_lattice(
    rng: np.random.Generator,
    n_inner: int,
    vocab: int,
    *,
    max_out: int = 3,
    parallel_arcs: bool = False,
    shuffle_states: bool = True,
):
    """A random acyclic mark lattice with ``n_inner`` states between start and pre-final.

    States: 0 (start) --bos--> 1 ... inner ... --eos--> S-1 (sink, pad self-loop).
    Every state is reachable from 0 and reaches the sink.  With ``parallel_arcs=False`` no
    two arcs share a (src, dst) pair (the regime where the reference's batched DP is
    valid, quirk Q4).  ``shuffle_states`` permutes the inner state ids so that state
    numbering is not topological (the reference never assumes it is).
    """
    n = n_inner + 3  # start, inner..., prefinal, sink
    sink = n - 1
    prefinal = n - 2
    arcs = []  # (src, label, dst) in topological numbering
    used = [set() for _ in range(n)]  # labels used per source
    pairs = set()

    def add(s, d):
        if not parallel_arcs and (s, d) in pairs:
            return False
        free = [l for l in range(FIRST_LABEL, vocab) if l not in used[s]]
        if not free:
            return False
        l = int(rng.choice(free))
        used[s].add(l)
        pairs.add((s, d))
        arcs.append((s, l, d))
        return True

    arcs.append((0, BOS, 1))
    used[0].add(BOS)
    pairs.add((0, 1))
    for s in range(1, prefinal):
        k = int(rng.integers(1, max_out + 1))
        hi = min(prefinal, s + 4)
        for _ in range(k):
            d = int(rng.integers(s + 1, hi + 1))
            add(s, d)
        if not any(a[0] == s for a in arcs):
            add(s, s + 1)
    # make every inner state reachable
    has_in = {a[2] for a in arcs}
    for d in range(2, prefinal + 1):
        if d not in has_in:
            lo = max(1, d - 4)
            cands = [s for s in range(lo, d) if (parallel_arcs or (s, d) not in pairs)]
            s = int(rng.choice(cands))
            assert add(s, d)
    arcs.append((prefinal, EOS, sink))
    if parallel_arcs:
        # force at least one pair of parallel arcs
        s, _, d = arcs[len(arcs) // 2]
        if s != 0 and d != sink:
            add(s, d)

    perm = np.arange(n)
    if shuffle_states and n_inner > 1:
        inner = np.arange(1, prefinal + 1)
        perm[1 : prefinal + 1] = rng.permutation(inner)
    transition = np.zeros((n, vocab), dtype=np.int64)
    for s, l, d in arcs:
        assert transition[perm[s], l] == 0
        transition[perm[s], l] = perm[d]
    transition[sink, PAD] = sink
    emission = transition != 0
    return emission, transition


def chain_with_skips(n: int, vocab: int):
    """Deterministic deep lattice: chain 0..n with skip arcs s -> s+2 (probe lattice of
    SURVEY.md section 6)."""
    S = n + 2
    sink = S - 1
    t = np.zeros((S, vocab), dtype=np.int64)
    t[0, BOS] = 1
    for s in range(1, n):
        t[s, FIRST_LABEL + (s % (vocab - FIRST_LABEL))] = s + 1
        if s + 2 <= n:
            l2 = FIRST_LABEL + ((s + 3) % (vocab - FIRST_LABEL))
            if t[s, l2] == 0:
                t[s, l2] = s + 2
    t[n, EOS] = sink
    t[sink, PAD] = sink
    return t != 0, t
