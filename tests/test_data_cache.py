"""``nfst_b200.data`` on the CPU: the reference's ``.npz`` examples, the packed cache beside them, ``collate``.

The reference path is ``FSADataset.__getitem__`` -> ``Utils.load_fsa_from_npz`` -> ``T9FSADataModule.collate``
(``util/dataset_reader.py:30-40,175-186``, ``util/preprocess_util.py:293-323,368-392``): six arrays per example, each
padded to the batch maximum with the PAD ID (quirk Q5), dense tables included.  The expected values here restate that
with numpy (``np.full(..., pad)`` + copy) and the DP over the result is checked against the oracle by replaying the
packed arrays, so that a cache or concatenation bug shows up without a GPU.
"""
import os

import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import data as nd
from oracle import lattice_oracle as lo
from tests.lattice_gen import PAD, random_mark_lattice
from tests.test_pack import _np, check_structure, replay_beta

V = 24


def reference_collate(examples, pad):
    """``collate`` of the reference: every one of the six arrays padded along axis 0 with the pad id."""
    out = []
    for i in range(6):
        seqs = [e[i] for e in examples]
        n = max(s.shape[0] for s in seqs)
        buf = np.full((len(seqs), n) + seqs[0].shape[1:], pad, dtype=seqs[0].dtype)
        for k, s in enumerate(seqs):
            buf[k, : s.shape[0], ...] = s
        out.append(buf)
    return tuple(out)


@pytest.fixture()
def examples(tmp_path):
    rng = np.random.default_rng(21)
    names, arrays = [], []
    for i in range(5):
        em, tr = random_mark_lattice(rng, int(rng.integers(3, 16)), V, parallel_arcs=bool(i % 2))
        gs = rng.integers(4, V, size=3 + i).astype(np.int64)
        ps = rng.integers(4, V, size=1 + 2 * i).astype(np.int64)
        name = str(tmp_path / f"ex{i}")
        # written the way src/preprocess/tr.py:182-190 writes them
        np.savez_compressed(name + ".npz", num_emission=em, num_transition=tr, denom_emission=em[:2], denom_transition=tr[:2],
                            gs=gs, ps=ps)
        names.append(name)
        arrays.append((em, tr, em[:2], tr[:2], gs, ps))
    return names, arrays


def test_load_fsa_from_npz_returns_the_six_arrays_in_the_reference_order(examples, tmp_path):
    names, arrays = examples
    got = nd.load_fsa_from_npz(names[1] + ".npz", None, V, PAD)
    assert len(got) == 6
    for g, w in zip(got, arrays[1]):
        assert g.dtype == w.dtype and np.array_equal(g, w)
    # the reference's own assertion text for a missing file (preprocess_util.py:294-296)
    missing = str(tmp_path / "nope.npz")
    with pytest.raises(AssertionError, match="does not exist! Please run preprocess_npz.py first."):
        nd.load_fsa_from_npz(missing, None, V, PAD)


def test_load_fsa_from_npz_appends_the_weighted_proposal_tables(examples, monkeypatch):
    """The ``wfst_name`` branch (preprocess_util.py:314-322): ``pynini.Fst.read`` + the weighted tables.  pynini is absent
    here, so a stand-in module supplies ``Fst.read`` / ``Weight`` (tests/lattice_gen.py); without any pynini the branch
    fails on its import, as the reference's does."""
    import sys
    import types

    from tests.lattice_gen import FakeFst, FakeWeight, random_fst_arrays

    names, arrays = examples
    monkeypatch.delitem(sys.modules, "pynini", raising=False)
    with pytest.raises(ImportError):
        nd.load_fsa_from_npz(names[0] + ".npz", "proposal.fst", V, PAD)
    machine = FakeFst(*random_fst_arrays(np.random.default_rng(2), 9, V))
    read = []

    class Fst:
        @staticmethod
        def read(path):
            read.append(path)
            return machine

    fake = types.ModuleType("pynini")
    fake.Fst, fake.Weight = Fst, FakeWeight
    monkeypatch.setitem(sys.modules, "pynini", fake)
    got = nd.load_fsa_from_npz(names[0] + ".npz", "proposal.fst", V, PAD)
    assert read == ["proposal.fst"] and len(got) == 8
    for g, w in zip(got[:6], arrays[0]):
        assert np.array_equal(g, w)
    em, tr = nd.get_state_mask_pynini(machine, V, PAD, to_numpy=True, weighted=True)
    assert got[6].dtype == np.float64 and np.array_equal(got[6], em) and np.array_equal(got[7], tr)


def test_dataset_packs_once_then_reads_the_cache(examples):
    names, arrays = examples
    ds = nd.LatticeDataset(names, V, PAD)
    assert len(ds) == len(names)
    first = [ds[i] for i in range(len(ds))]
    assert all(os.path.exists(n + ".packed.npz") for n in names)
    stamp = [os.path.getmtime(n + ".packed.npz") for n in names]
    again = [ds[i] for i in range(len(ds))]
    assert stamp == [os.path.getmtime(n + ".packed.npz") for n in names]  # read, not rewritten
    for a, b, arr in zip(first, again, arrays):
        assert np.array_equal(a.gs, arr[4]) and np.array_equal(b.ps, arr[5])
        assert (a.packed.n_lattices, a.packed.n_states, a.packed.n_arcs, a.packed.vocab) == \
               (b.packed.n_lattices, b.packed.n_states, b.packed.n_arcs, b.packed.vocab)
        assert a.packed.dense_shape == b.packed.dense_shape == (1,) + arr[1].shape
        assert sorted(a.packed.tensors()) == sorted(b.packed.tensors())
        for f in a.packed.tensors():
            ta, tb = getattr(a.packed, f), getattr(b.packed, f)
            assert ta.dtype == tb.dtype and torch.equal(ta, tb), f
        assert len(a.packed.groups) == len(b.packed.groups)
        for ga, gb in zip(a.packed.groups, b.packed.groups):
            assert nd._group_fields(ga) == nd._group_fields(gb)
            assert torch.equal(ga.ids, gb.ids)
        check_structure(b.packed)


def test_stale_or_foreign_cache_is_rebuilt(examples, tmp_path):
    names, _ = examples
    cache_dir = str(tmp_path / "cache")
    ds = nd.LatticeDataset(names[:2], V, PAD, cache_dir=cache_dir)
    good = ds[0].packed
    path = ds.cache_path(names[0])
    assert os.path.dirname(path) == cache_dir and os.path.exists(path) and not os.path.exists(names[0] + ".packed.npz")
    # a cache of another layout version
    with np.load(path) as l:
        arrays = {k: l[k] for k in l.files}
    arrays["__format__"] = np.array([nd.PACKED_FORMAT + 1])
    np.savez(path, **arrays)
    with pytest.raises(ValueError, match="another format"):
        nd.load_packed(path)
    rebuilt = ds[0].packed
    assert torch.equal(rebuilt.dst_out, good.dst_out)
    assert int(np.load(path)["__format__"][0]) == nd.PACKED_FORMAT
    # a truncated file
    with open(path, "wb") as f:
        f.write(b"not an npz")
    assert torch.equal(ds[0].packed.label_out, good.label_out)
    # a dense file newer than its cache wins
    nd.load_packed(path)
    past = os.path.getmtime(path) - 1000.0
    os.utime(path, (past, past))
    os.utime(names[0] + ".npz", (past + 500.0, past + 500.0))
    ds[0]
    assert os.path.getmtime(path) > past + 500.0  # rewritten


def test_collate_of_cached_examples_equals_the_reference_collate_then_pack(examples):
    names, arrays = examples
    ds = nd.LatticeDataset(names, V, PAD)
    [ds[i] for i in range(len(ds))]
    batch = [ds[i] for i in range(len(ds))]  # from the caches
    packed, gs, ps = nd.collate(batch, PAD)
    ne, nt, de, dt, gs_ref, ps_ref = reference_collate(arrays, PAD)
    assert gs.dtype == torch.int64 and np.array_equal(gs.numpy(), gs_ref) and np.array_equal(ps.numpy(), ps_ref)
    # the reference would now hand the padded dense tables to the DP: same lattices, state for state
    joint = nb.pack_dense(torch.from_numpy(ne), torch.from_numpy(nt.astype(np.int64)))
    check_structure(packed)
    assert (packed.n_lattices, packed.n_states, packed.n_arcs) == (joint.n_lattices, joint.n_states, joint.n_arcs)
    for f in ("state_off", "level_off", "level_ptr", "out_ptr", "dst_out", "label_out", "orig_state", "start_state"):
        assert torch.equal(getattr(packed, f), getattr(joint, f)), f
    theta = np.random.default_rng(4).normal(size=V)
    beta = replay_beta(packed, theta[_np(packed.label_out)])
    state_off, orig = _np(packed.state_off), _np(packed.orig_state)
    for b, arr in enumerate(arrays):
        tr = arr[1]
        src, lab, dst, _ = lo.arcs_from_dense(tr)
        logz, _, be, _ = lo.forward_backward(tr.shape[0], src, dst, theta[lab])
        sl = slice(state_off[b], state_off[b + 1])
        np.testing.assert_allclose(beta[sl], be[orig[sl]], atol=1e-12)
        assert abs(beta[_np(packed.start_state)[b]] - logz) < 1e-12
    # a batch of one is the example's own pack
    one, g1, p1 = nd.collate(batch[3:4], PAD)
    assert one is batch[3].packed and g1.shape == (1, len(arrays[3][4])) and p1.shape == (1, len(arrays[3][5]))


def test_dataset_under_a_multi_worker_dataloader(examples):
    """As the reference drives its dataset (``dataset_reader.py:148-155``: DataLoader workers + ``collate_fn``): the packed
    batches cross the process boundary intact, epoch after epoch (the second epoch reads the caches)."""
    import functools

    names, arrays = examples
    ds = nd.LatticeDataset(names[:4], V, PAD)
    dl = torch.utils.data.DataLoader(ds, batch_size=2, num_workers=2, collate_fn=functools.partial(nd.collate, pad=PAD),
                                     multiprocessing_context="spawn", persistent_workers=True)
    want = [nd.collate([ds[0], ds[1]], PAD), nd.collate([ds[2], ds[3]], PAD)]
    for epoch in range(2):
        got = list(dl)
        assert len(got) == 2
        for (p, gs, ps), (q, gs_w, ps_w) in zip(got, want):
            assert torch.equal(gs, gs_w) and torch.equal(ps, ps_w)
            assert (p.n_lattices, p.n_states, p.n_arcs) == (q.n_lattices, q.n_states, q.n_arcs)
            for f in q.tensors():
                assert torch.equal(getattr(p, f), getattr(q, f)), f
            assert [nd._group_fields(g) for g in p.groups] == [nd._group_fields(g) for g in q.groups]
            check_structure(p)


def test_dataset_is_constructed_the_way_the_reference_constructs_fsadataset(examples, monkeypatch):
    """``FSADataset(current_split, vocab_size=..., pad=..., list_of_wfst_proposals=...)`` (dataset_reader.py:98-103): the
    same keywords work; with proposals every example also carries the weighted proposal tables."""
    import sys
    import types

    from tests.lattice_gen import FakeFst, FakeWeight, random_fst_arrays

    names, _ = examples
    ds = nd.LatticeDataset(names[:2], vocab_size=V, pad=PAD, list_of_wfst_proposals=None)
    assert ds[0].proposal_tables is None
    machines = {f"p{i}.fst": FakeFst(*random_fst_arrays(np.random.default_rng(30 + i), 7 + i, V)) for i in range(2)}
    fake = types.ModuleType("pynini")
    fake.Fst = type("Fst", (), {"read": staticmethod(lambda path: machines[path])})
    fake.Weight = FakeWeight
    monkeypatch.setitem(sys.modules, "pynini", fake)
    ds = nd.LatticeDataset(names[:2], vocab_size=V, pad=PAD, list_of_wfst_proposals=list(machines))
    for i in range(2):
        for ex in (ds[i], ds[i]):  # packed now, from the cache next
            em, tr = ex.proposal_tables
            want = nd.get_state_mask_pynini(machines[f"p{i}.fst"], V, PAD, to_numpy=True, weighted=True)
            assert np.array_equal(em, want[0]) and np.array_equal(tr, want[1])
    # collate: the proposal tables ride along, padded with the pad id like every array of the reference's batch (Q5)
    out = nd.collate([ds[0], ds[1]], PAD)
    assert len(out) == 5
    tabs = [nd.get_state_mask_pynini(machines[f"p{i}.fst"], V, PAD, to_numpy=True, weighted=True) for i in range(2)]
    for k, got in ((0, out[3]), (1, out[4])):
        want = reference_collate([(t[k],) * 6 for t in tabs], PAD)[0]
        assert got.dtype == torch.from_numpy(want).dtype and np.array_equal(got.numpy(), want)
    assert out[3].dtype == torch.float64 and float(out[3][0, -1, 0]) == float(PAD)  # the float table too
