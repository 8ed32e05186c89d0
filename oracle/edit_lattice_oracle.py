"""Host restatement of the reference's lattice construction for the transliteration edit machine -- TEST
INFRASTRUCTURE, NOT PRODUCT (only tests/ may import this).

What the reference computes (with OpenFst, absent here): ``x o T o y`` for the one-state machine ``T`` of
``src/fsm/tr.py:321-390`` (insertion arcs eps:y with marks [output-mark, y]; deletion arcs x:eps with marks
[input-mark, x]; with add_sub substitution arcs x:y with marks [insertion-mark, input-mark, x, output-mark, y]),
``src/preprocess/tr.py:158-165``; every arc replaced by the chain of its marks
(``OwnAST.mfst_weight_projection``, ``src/modules/path_semiring.py:120-180``); bos in front, eos behind
(``src/preprocess/preprocess.py:51-174``).  Composition with a one-state machine whose arcs consume one input
symbol, one output symbol, or both is the edit grid over (i, j) = (symbols of x consumed, symbols of y produced).
PARITY UNPINNED against the reference's own output (mfst / pynini are not installable here, and pynini's
optimize() renumbers states): pinned by brute-force enumeration of the alignments instead (tests).
"""
from __future__ import annotations

import itertools
from typing import List, Sequence, Tuple


def edit_lattice(x: Sequence[int], y: Sequence[int], *, bos: int, eos: int, input_mark: int, output_mark: int,
                 sub_mark=None) -> Tuple[List[Tuple[int, int, int]], int]:
    """(arcs [(src, label, dst)], n_states) with plain Python loops; state 0 is the start, the last state the sink."""
    n, m = len(x), len(y)
    next_id = [1]
    grid = {}
    for i in range(n + 1):
        for j in range(m + 1):
            grid[(i, j)] = next_id[0]
            next_id[0] += 1

    def fresh():
        next_id[0] += 1
        return next_id[0] - 1

    arcs = [(0, bos, grid[(0, 0)])]

    def chain(s, marks, d):
        cur = s
        for k, l in enumerate(marks):
            nxt = d if k == len(marks) - 1 else fresh()
            arcs.append((cur, l, nxt))
            cur = nxt

    for i in range(n + 1):
        for j in range(m + 1):
            if i < n:  # deletion of x_i
                chain(grid[(i, j)], [input_mark, x[i]], grid[(i + 1, j)])
            if j < m:  # insertion of y_j
                chain(grid[(i, j)], [output_mark, y[j]], grid[(i, j + 1)])
            if sub_mark is not None and i < n and j < m:
                chain(grid[(i, j)], [sub_mark, input_mark, x[i], output_mark, y[j]], grid[(i + 1, j + 1)])
    sink = fresh()
    arcs.append((grid[(n, m)], eos, sink))
    return arcs, sink + 1


def mark_strings(x: Sequence[int], y: Sequence[int], *, bos: int, eos: int, input_mark: int, output_mark: int, sub_mark=None):
    """Every mark string of the lattice by brute force over the alignments (small inputs only): all interleavings
    of deletions, insertions and (with sub_mark) substitutions that consume x and produce y."""
    out = []

    def rec(i, j, acc):
        if i == len(x) and j == len(y):
            out.append(tuple([bos] + acc + [eos]))
            return
        if i < len(x):
            rec(i + 1, j, acc + [input_mark, x[i]])
        if j < len(y):
            rec(i, j + 1, acc + [output_mark, y[j]])
        if sub_mark is not None and i < len(x) and j < len(y):
            rec(i + 1, j + 1, acc + [sub_mark, input_mark, x[i], output_mark, y[j]])

    rec(0, 0, [])
    return out
