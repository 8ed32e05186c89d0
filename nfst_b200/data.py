"""Lattice files: the reference's dense ``.npz`` examples and a packed cache beside them.

The reference stores one example per ``<name>.npz`` with six arrays -- ``num_emission``, ``num_transition``,
``denom_emission``, ``denom_transition``, ``gs``, ``ps`` (``src/preprocess/tr.py:182-190``) -- and re-reads,
inflates and pads the dense ``[S, V]`` tables on every epoch (``Utils.load_fsa_from_npz``,
``src/util/preprocess_util.py:293-323``; ``FSADataset.__getitem__``, ``src/util/dataset_reader.py:30-40``;
``collate``, ``:175-186``) although fewer than 1 % of the cells are arcs.  Here an example is packed ONCE
(``pack_dense``: the reference's edge rule, levels, CSR / column layouts) and the packed arrays are written next
to the dense file (``<name>.packed.npz``); later epochs load that file -- a few kilobytes per example -- and
``collate`` concatenates the cached packs (``concat_packed``: a fixed number of tensor ops, no re-packing).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .pack import LaunchGroup, PackedLattices, build_groups, concat_packed, pack_arcs, pack_dense

DENSE_KEYS = ("num_emission", "num_transition", "denom_emission", "denom_transition", "gs", "ps")
PACKED_FORMAT = 2  # bump when the packed layout changes: stale caches are rebuilt


def load_fsa_from_npz(npz_fname: str, wfst_name=None, vocab_size=None, pad=None) -> Tuple[np.ndarray, ...]:
    """The six arrays of a reference example, in the reference's order and with its error for a missing file
    (``preprocess_util.py:293-310``).  With ``wfst_name`` the weighted proposal FST is read through pynini and its
    weighted tables are appended, as the reference does (``:314-322``) -- pynini is that branch's own dependency; the
    tables come from ``get_state_mask_pynini`` below."""
    assert os.path.exists(npz_fname), f"{npz_fname} does not exist! Please run preprocess_npz.py first."
    with np.load(npz_fname) as l:
        to_return = tuple(l[k] for k in DENSE_KEYS)
    if wfst_name is not None:
        from pynini import Fst

        loaded_fst = Fst.read(wfst_name)
        matrices = get_state_mask_pynini(loaded_fst, vocab_size, pad, to_numpy=True, weighted=True)
        to_return = tuple(list(to_return) + list(matrices))
    return to_return


def fst_arc_list(machine, final_zero=None):
    """Arc list of an OpenFst-style acceptor as ``get_state_mask_pynini`` reads it (``scorers.py:1005-1032``):
    ``(N, src[A], ilabel[A], nextstate[A], weight[A] float64)`` with every arc into a FINAL state redirected to the
    sink row ``N = machine.num_states()`` (``:1023-1025``).  ``machine`` needs ``start() / num_states() /
    arcs(state) / final(state) / weight_type()`` -- a ``pynini.Fst`` or anything shaped like one; ``final_zero`` is the
    semiring zero that marks non-final states (default: ``pynini.Weight.zero(machine.weight_type())``, as the
    reference computes it, ``:1002``).  Same assertions as the reference: the start is state 0 (``:1005``), one arc
    per (state, label) (``:1030``)."""
    if final_zero is None:
        from pynini import Weight  # the reference's own dependency for this function (scorers.py:999)

        final_zero = Weight.zero(machine.weight_type())
    assert machine.start() == 0
    n = int(machine.num_states())
    is_final = np.array([machine.final(s) != final_zero for s in range(n)], dtype=bool)
    src, lab, nxt, wgt = [], [], [], []
    for state in range(n):
        for arc in machine.arcs(state):
            src.append(state)
            lab.append(int(arc.ilabel))
            nxt.append(int(arc.nextstate))
            wgt.append(float(arc.weight))
    src, lab, nxt = (np.asarray(a, dtype=np.int64) for a in (src, lab, nxt))
    key = src * (int(lab.max()) + 1 if lab.size else 1) + lab
    assert np.unique(key).size == key.size  # "assert arc.ilabel not in ilabel_set"
    nxt = np.where(is_final[nxt], n, nxt) if nxt.size else nxt
    return n, src, lab, nxt, np.asarray(wgt, dtype=np.float64)


def get_state_mask_pynini(machine, vocab_size: int, pad: int, to_numpy: bool = False, weighted: bool = False, final_zero=None):
    """``FSAGRUScorer.get_state_mask_pynini`` (``scorers.py:995-1035``) -- the function that DEFINES the lattice
    tables -- without the per-arc Python writes: ``emission[N+1, V]`` (bool, or float64 log-weights ``-arc.weight``
    with ``-inf`` = no arc when ``weighted``), ``transition[N+1, V]`` int64 next state; row ``N`` is the sink with its
    ``pad`` self-loop (``:1011-1016``).  (The reference's ``weighted=True`` branch needs ``np.float``, gone from
    current numpy; the dtype it named is float64.)"""
    n, src, lab, nxt, wgt = fst_arc_list(machine, final_zero)
    if weighted:
        emission = np.full((n + 1, vocab_size), -np.inf, dtype=np.float64)
        emission[n, pad] = 0.0
        emission[src, lab] = -wgt
    else:
        emission = np.zeros((n + 1, vocab_size), dtype=bool)
        emission[n, pad] = True
        emission[src, lab] = True
    transition = np.zeros((n + 1, vocab_size), dtype=np.int64)
    transition[n, pad] = n
    transition[src, lab] = nxt
    if to_numpy:
        return emission, transition
    return torch.from_numpy(emission), torch.from_numpy(transition)


def pack_fsts(machines: Sequence, vocab_size: int, weighted: bool = False, device=None, final_zero=None, **pack_kw) -> PackedLattices:
    """A batch of acceptors straight to the packed layout, skipping the dense tables (fewer than 1 % of their cells
    are arcs): the arcs ``get_state_mask_pynini`` would write, filtered by the DP's edge rule -- a cell is an arc iff
    its next state is neither 0 nor the row itself (``scorers.py:704-716``) -- and packed with ``pack_arcs``.  Same
    lattices, state for state, as ``pack_dense`` of the collated tables; ``weighted`` keeps ``-arc.weight`` as static
    arc scores (``scorers.py:1026-1027``)."""
    lat, src, dst, lab, sc, ns = [], [], [], [], [], []
    for b, m in enumerate(machines):
        n, s, l, t, w = fst_arc_list(m, final_zero)
        keep = (t != 0) & (t != s)
        lat.append(np.full(int(keep.sum()), b, dtype=np.int64))
        src.append(s[keep]); dst.append(t[keep]); lab.append(l[keep]); sc.append(-w[keep]); ns.append(n + 1)
    if lab and max((int(l.max()) for l in lab if l.size), default=0) >= vocab_size:
        raise ValueError("label out of range")
    dev = torch.device(device) if device is not None else torch.device("cpu")
    cat = lambda xs: torch.from_numpy(np.concatenate(xs)).to(dev)  # noqa: E731
    static = cat(sc).to(torch.float32) if weighted else None
    return pack_arcs(cat(lat), cat(src), cat(dst), cat(lab), torch.tensor(ns, dtype=torch.int64, device=dev), vocab_size,
                     static_scores=static, **pack_kw)


def packed_to_dense(packed: PackedLattices, pad: int, weighted: bool = False):
    """The way back: ``(emission[B, S, V], transition[B, S, V])`` in the format ``collate`` hands the reference's modules
    (``scorers.py:995-1035`` tables, padded with the pad id, ``dataset_reader.py:175-186``) -- for feeding lattices that
    never had dense tables (``construct.edit_lattices``, ``pack_fsts``, ``pack_arcs``) to the reference's own consumers
    (``FSAGRUScorer.set_masks``, ``scorers.py:877-885``).  Rows are the PACKED local state ids (level order: the start
    is row 0, no arc enters it; every state without outgoing arcs gets the sink's ``pad`` self-loop, ``:1013-1016``);
    ``weighted`` writes the static arc scores as float log-weights with ``-inf`` for "no arc" (``:1011-1013``).  Needs
    one arc per (state, label), like the tables themselves (``:1030``)."""
    B, V = packed.n_lattices, packed.vocab
    state_off = packed.state_off.to(torch.int64)
    n_b = state_off[1:] - state_off[:-1]
    S = int(n_b.max())
    dev = packed.device
    out_ptr = packed.out_ptr[: packed.n_states + 1].to(torch.int64)
    deg = out_ptr[1:] - out_ptr[:-1]
    src = torch.repeat_interleave(torch.arange(packed.n_states, device=dev), deg)
    lat = torch.bucketize(src, state_off[1:], right=True)
    dst = packed.dst_out[: packed.n_arcs].to(torch.int64)
    lab = packed.label_out[: packed.n_arcs].to(torch.int64)
    cell = (lat * S + (src - state_off[lat])) * V + lab
    if cell.numel() and int(torch.unique(cell).numel()) != int(cell.numel()):
        raise ValueError("two arcs of one state share a label: the dense tables hold one arc per (state, label)")
    transition = torch.zeros(B * S * V, dtype=torch.int64, device=dev)
    transition[cell] = dst - state_off[lat]
    if weighted:
        if packed.static_scores is None:
            raise ValueError("weighted tables need a batch packed with static arc scores")
        emission = torch.full((B * S * V,), float("-inf"), dtype=torch.float64, device=dev)
        emission[cell] = packed.static_scores[: packed.n_arcs].to(torch.float64)
        on = 0.0
    else:
        emission = torch.zeros(B * S * V, dtype=torch.bool, device=dev)
        emission[cell] = True
        on = True
    sinks = torch.nonzero(deg == 0).squeeze(1)
    sl = torch.bucketize(sinks, state_off[1:], right=True)
    scell = (sl * S + (sinks - state_off[sl])) * V + int(pad)
    transition[scell] = sinks - state_off[sl]
    emission[scell] = on
    transition, emission = transition.view(B, S, V), emission.view(B, S, V)
    padded = torch.arange(S, device=dev)[None, :] >= n_b[:, None]  # collate's rows: the pad id everywhere (quirk Q5)
    transition[padded] = int(pad)
    emission[padded] = float(pad) if weighted else True
    return emission, transition


def _group_fields(g: LaunchGroup):
    return {k: v for k, v in g.__dict__.items() if not isinstance(v, torch.Tensor) and v is not None}


def save_packed(path: str, packed: PackedLattices, compress: bool = False) -> None:
    """Write a packed batch (usually one example) as an ``.npz``: every array of ``PackedLattices`` plus the
    per-lattice statistics the launch groups are rebuilt from."""
    arrays = {"__format__": np.array([PACKED_FORMAT]), "__scalars__": np.array(
        [packed.n_lattices, packed.n_states, packed.n_arcs, packed.vocab, packed.max_levels], dtype=np.int64)}
    for name in packed.tensors():
        arrays[name] = getattr(packed, name).cpu().numpy()
    for k, v in packed.stats.items():
        arrays["stat__" + k] = v.cpu().numpy()
    if packed.dense_shape is not None:
        arrays["__dense_shape__"] = np.array(packed.dense_shape, dtype=np.int64)
    tmp = path + ".tmp.npz"
    (np.savez_compressed if compress else np.savez)(tmp, **arrays)
    os.replace(tmp, path)


def load_packed(path: str, device="cpu") -> PackedLattices:
    """Inverse of ``save_packed``.  Raises ``ValueError`` for a cache written by another layout version."""
    with np.load(path) as l:
        if "__format__" not in l.files or int(l["__format__"][0]) != PACKED_FORMAT:
            raise ValueError(f"{path}: packed-lattice cache of another format; re-pack it")
        B, S, A, V, max_levels = (int(x) for x in l["__scalars__"])
        stats = {k[6:]: torch.from_numpy(l[k]) for k in l.files if k.startswith("stat__")}
        kw = {k: torch.from_numpy(l[k]) for k in l.files if not k.startswith("__") and not k.startswith("stat__")}
        dense_shape = tuple(int(x) for x in l["__dense_shape__"]) if "__dense_shape__" in l.files else None
    kw = {k: v.to(device) for k, v in kw.items()}
    kw.setdefault("static_scores", None)
    lazy = A > 0 and int(kw["in2out"].numel()) != A  # packed without the in-order arrays (column-major lattices)
    ci = None if lazy else {d: (kw[f"{d}_chunk_off"].to(torch.int64), kw[f"{d}_chunks"], kw[f"{d}_chunk_level"]) for d in ("fwd", "bwd")}
    groups = build_groups(stats, torch.device(device), ci)
    return PackedLattices(n_lattices=B, n_states=S, n_arcs=A, vocab=V, dense_shape=dense_shape, groups=groups,
                          max_levels=max_levels, stats=stats, **kw)


class PackedExample:
    """What ``LatticeDataset.__getitem__`` returns: the numerator lattice packed, and ``gs`` / ``ps`` as read."""

    def __init__(self, packed: PackedLattices, gs: np.ndarray, ps: np.ndarray, name: str, proposal_tables=None):
        self.packed, self.gs, self.ps, self.name = packed, gs, ps, name
        self.proposal_tables = proposal_tables  # (emission float64 [N+1, V], transition int64) of the weighted proposal FST, if any


class LatticeDataset(torch.utils.data.Dataset):
    """Counterpart of ``FSADataset`` (``dataset_reader.py:16-43``) over the same list of example names.

    ``__getitem__`` returns the numerator lattice of ``<name>.npz`` packed (the denominator tables are cyclic
    and never fed to the DP: ``denom_prob`` is hard-wired to zero, ``lightning.py:473``).  The first access packs
    the dense tables and writes ``<name>.packed.npz`` (in ``cache_dir`` if given); later accesses read that."""

    def __init__(self, list_of_machines: Sequence[str], vocab_size: Optional[int] = None, pad: Optional[int] = None,
                 cache_dir: Optional[str] = None, weighted: Optional[bool] = None, *,
                 list_of_wfst_proposals: Optional[Sequence[str]] = None):
        self.l = list(list_of_machines)
        self.proposals = None if list_of_wfst_proposals is None else list(list_of_wfst_proposals)  # dataset_reader.py:18,23
        self.vocab_size, self.pad = vocab_size, pad
        self.cache_dir, self.weighted = cache_dir, weighted

    def _proposal(self, index: int):
        """the weighted proposal tables of example ``index`` (``dataset_reader.py:33-36`` -> ``preprocess_util.py:314-322``)"""
        if self.proposals is None:
            return None
        from pynini import Fst

        return get_state_mask_pynini(Fst.read(self.proposals[index]), self.vocab_size, self.pad, to_numpy=True, weighted=True)

    def __len__(self) -> int:
        return len(self.l)

    def cache_path(self, name: str) -> str:
        if self.cache_dir is None:
            return f"{name}.packed.npz"
        return os.path.join(self.cache_dir, name.strip(os.sep).replace(os.sep, "__") + ".packed.npz")

    def __getitem__(self, index: int) -> PackedExample:
        name = self.l[index]
        dense, cache = f"{name}.npz", self.cache_path(name)
        if os.path.exists(cache) and (not os.path.exists(dense) or os.path.getmtime(cache) >= os.path.getmtime(dense)):
            try:
                with np.load(dense) as l:
                    gs, ps = l["gs"], l["ps"]
                return PackedExample(load_packed(cache), gs, ps, name, self._proposal(index))
            except (ValueError, OSError, KeyError):
                pass  # stale or unreadable cache: rebuild it (the reference re-serialises unreadable files too, tr.py:135-141)
        ne, nt, _, _, gs, ps = load_fsa_from_npz(dense, None, self.vocab_size, self.pad)
        em = torch.from_numpy(np.ascontiguousarray(ne))[None]
        tr = torch.from_numpy(np.ascontiguousarray(nt).astype(np.int64))[None]
        packed = pack_dense(em, tr, weighted=self.weighted)
        if self.cache_dir is not None:
            os.makedirs(self.cache_dir, exist_ok=True)
        save_packed(cache, packed)
        return PackedExample(packed, gs, ps, name, self._proposal(index))


def pad_sequence_1d(seqs: List[np.ndarray], padding_value: int) -> torch.Tensor:
    """``Utils.pad_sequence`` (``preprocess_util.py:368-392``) for the 1-D ``gs`` / ``ps`` arrays."""
    n = max(len(s) for s in seqs)
    out = torch.full((len(seqs), n), int(padding_value), dtype=torch.int64)
    for i, s in enumerate(seqs):
        out[i, : len(s)] = torch.from_numpy(np.asarray(s).astype(np.int64))
    return out


def collate(batch: List[PackedExample], pad: int, device=None):
    """Counterpart of ``T9FSADataModule.collate`` (``dataset_reader.py:175-186``): (packed batch, gs[B, Lx],
    ps[B, Ly]) -- plus the padded proposal tables when the examples carry them.  No numerator table is padded or
    copied; the per-example packs are concatenated."""
    packed = concat_packed([b.packed for b in batch]) if len(batch) > 1 else batch[0].packed
    if device is not None:
        packed = packed.to(device)
    out = (packed, pad_sequence_1d([b.gs for b in batch], pad), pad_sequence_1d([b.ps for b in batch], pad))
    if batch[0].proposal_tables is not None:
        # the two weighted proposal tables ride along dense, padded like every array of the reference's batch --
        # with the pad id, the float emission table included (quirk Q5)
        out = out + tuple(pad_tables([b.proposal_tables[i] for b in batch], pad) for i in (0, 1))
    return out


def pad_tables(tables: List[np.ndarray], padding_value) -> torch.Tensor:
    """``Utils.pad_sequence(batch_first=True)`` (``preprocess_util.py:368-392``) for ``[S_b, V]`` tables: ``[B, max S_b, V]``
    filled with ``padding_value`` in the tables' own dtype."""
    n = max(t.shape[0] for t in tables)
    out = np.full((len(tables), n) + tables[0].shape[1:], padding_value, dtype=tables[0].dtype)
    for i, t in enumerate(tables):
        out[i, : t.shape[0], ...] = t
    return torch.from_numpy(out)
