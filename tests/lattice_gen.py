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


# ---- OpenFst-shaped acceptors without OpenFst (what get_state_mask_pynini reads, scorers.py:995-1035) ----
class FakeWeight:
    """The part of ``pynini.Weight`` that ``get_state_mask_pynini`` touches: ``zero(type)``, ``!=``, ``float()``."""

    def __init__(self, value: float):
        self.value = float(value)

    @classmethod
    def zero(cls, weight_type):
        assert weight_type == "tropical"
        return cls(float("inf"))

    def __eq__(self, other):
        return isinstance(other, FakeWeight) and self.value == other.value

    def __ne__(self, other):
        return not self.__eq__(other)

    def __float__(self):
        return self.value


class FakeArc:
    def __init__(self, ilabel, nextstate, weight):
        self.ilabel, self.olabel, self.nextstate, self.weight = int(ilabel), int(ilabel), int(nextstate), FakeWeight(weight)


class FakeFst:
    """``start() / num_states() / arcs(s) / final(s) / weight_type()`` over plain arrays."""

    def __init__(self, n_states, src, ilabel, nextstate, weight, is_final):
        self.n = int(n_states)
        self.rows = [[] for _ in range(self.n)]
        for s, l, t, w in zip(src, ilabel, nextstate, weight):
            self.rows[int(s)].append(FakeArc(l, t, w))
        self.is_final = np.asarray(is_final, dtype=bool)

    def start(self):
        return 0

    def num_states(self):
        return self.n

    def weight_type(self):
        return "tropical"

    def arcs(self, state):
        return iter(self.rows[state])

    def final(self, state):
        return FakeWeight(0.0 if self.is_final[state] else float("inf"))


def random_fst_arrays(rng: np.random.Generator, n_states: int, vocab: int):
    """Arrays of a random acyclic acceptor the way the offline pipeline leaves them: state 0 the start with one
    ``bos`` arc, one arc per (state, label), states without outgoing arcs final (reached by ``eos``), a few extra
    final states that also have outgoing arcs, one arc back into state 0 and one self-loop (both vanish under the
    DP's edge rule)."""
    n = max(int(n_states), 3)
    src, lab, nxt = [0], [BOS], [1]
    for s in range(1, n - 1):
        k = int(rng.integers(1, 4))
        labels = rng.choice(np.arange(FIRST_LABEL, vocab), size=k, replace=False)
        for l in labels:
            d = int(rng.integers(s + 1, n))
            src.append(s); lab.append(EOS if d == n - 1 else int(l)); nxt.append(d)
    # one arc per (state, label): eos may have been drawn twice for a state
    seen, keep = set(), []
    for i, (s, l) in enumerate(zip(src, lab)):
        if (s, l) not in seen:
            seen.add((s, l)); keep.append(i)
    src, lab, nxt = ([a[i] for i in keep] for a in (src, lab, nxt))
    mid = int(rng.integers(1, n - 1))
    for extra in ((mid, 4, 0), (mid, 5, mid)):  # into the start / a self-loop
        if (extra[0], extra[1]) not in seen:
            src.append(extra[0]); lab.append(extra[1]); nxt.append(extra[2])
    is_final = np.zeros(n, dtype=bool)
    is_final[n - 1] = True
    if n > 4:
        is_final[int(rng.integers(2, n - 1))] = True  # arcs into it are redirected to the sink row too
    w = np.round(rng.uniform(0.0, 4.0, size=len(src)), 3)
    return n, np.asarray(src), np.asarray(lab), np.asarray(nxt), w, is_final
