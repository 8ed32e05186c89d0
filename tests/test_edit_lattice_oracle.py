"""The host restatement of the edit-lattice construction (oracle/edit_lattice_oracle.py) against brute force: the
set of start -> sink label strings of the lattice is exactly the set of mark strings of all alignments."""
import numpy as np

from oracle import edit_lattice_oracle as eo

MARKS = dict(bos=1, eos=2, input_mark=5, output_mark=4, sub_mark=6)


def paths(arcs, n_states):
    out = {}
    for s, l, d in arcs:
        out.setdefault(s, []).append((l, d))
    res = []

    def rec(s, acc):
        if s not in out:
            res.append(tuple(acc))
            return
        for l, d in out[s]:
            rec(d, acc + [l])

    rec(0, [])
    return res


def test_lattice_paths_are_the_alignments():
    rng = np.random.default_rng(0)
    for n, m, sub in [(0, 0, True), (1, 0, True), (0, 2, False), (2, 2, True), (3, 2, True), (2, 3, False)]:
        x, y = rng.integers(10, 20, size=n).tolist(), rng.integers(20, 30, size=m).tolist()
        marks = dict(MARKS)
        if not sub:
            marks["sub_mark"] = None
        arcs, S = eo.edit_lattice(x, y, **marks)
        got = paths(arcs, S)
        want = eo.mark_strings(x, y, **marks)
        assert len(got) == len(set(got)) and sorted(got) == sorted(want)
        assert max(max(s, d) for s, _, d in arcs) == S - 1
