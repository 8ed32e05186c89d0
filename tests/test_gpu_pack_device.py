"""The device packer (nfst_pack.cu: nfst_pack_small) against the tensor-op packer and the oracle.

Structure: the packed arrays describe the same lattice (arc_origin / orig_state are permutations onto the kept
arcs / states, CSR by source is in (state, label) order, CSR by destination in (state, canonical id) order, states
are numbered by level and every arc goes to a deeper level, the rows collate() padding adds are trimmed).
Results: beta / logZ / posteriors / Viterbi of the device-packed batch meet the same bounds against the C oracle as
everything else, and equal the tensor-op packer's results in the original numbering.
"""
import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import pack as P
from nfst_b200 import synth
from oracle import c_oracle
from tests.lattice_gen import random_mark_lattice

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


def _np(t):
    return t.cpu().numpy().astype(np.int64)


def check_structure(p: nb.PackedLattices, n_raw_arcs: int):
    S, A, B = p.n_states, p.n_arcs, p.n_lattices
    state_off, arc_off, level_off, level_ptr = _np(p.state_off), _np(p.arc_off), _np(p.level_off), _np(p.level_ptr)
    out_ptr, in_ptr, dst, lab, src_out = _np(p.out_ptr), _np(p.in_ptr), _np(p.dst_out), _np(p.label_out), _np(p.src_out)
    src_in, lab_in, in2out = _np(p.src_in), _np(p.label_in), _np(p.in2out)
    assert state_off[0] == 0 and state_off[-1] == S and arc_off[-1] == A and out_ptr[S] == A and in_ptr[S] == A
    assert len(level_ptr) == level_off[-1]
    level_of = np.full(S, -1)
    for b in range(B):
        lp = level_ptr[level_off[b]:level_off[b + 1]]
        assert lp[0] == state_off[b] and lp[-1] == state_off[b + 1] and np.all(np.diff(lp) > 0)
        assert len(lp) - 1 == int(p.n_levels[b])
        for l in range(len(lp) - 1):
            level_of[lp[l]:lp[l + 1]] = l
        assert int(p.start_state[b]) == state_off[b], "the start state is the only state of level 0"
        assert np.all(out_ptr[state_off[b]:state_off[b + 1] + 1] >= arc_off[b]) and out_ptr[state_off[b + 1]] == arc_off[b + 1]
    # CSR by source, (state, label) order; every arc goes to a deeper level of the same lattice
    np.testing.assert_array_equal(src_out, np.repeat(np.arange(S), np.diff(out_ptr)))
    assert np.all(level_of[dst] > level_of[src_out])
    same = src_out[1:] == src_out[:-1]
    assert np.all(lab[1:][same] >= lab[:-1][same])
    # CSR by destination, (state, canonical id) order
    np.testing.assert_array_equal(dst[in2out], np.repeat(np.arange(S), np.diff(in_ptr)))
    np.testing.assert_array_equal(src_in, src_out[in2out])
    np.testing.assert_array_equal(lab_in, lab[in2out])
    d_in = dst[in2out]
    same = d_in[1:] == d_in[:-1]
    assert np.all(in2out[1:][same] > in2out[:-1][same])
    assert sorted(in2out.tolist()) == list(range(A))
    # sinks
    sinks, sink_off = _np(p.sinks), _np(p.sink_off)
    np.testing.assert_array_equal(sinks, np.nonzero(np.diff(out_ptr) == 0)[0])
    assert sink_off[-1] == len(sinks)
    origin = _np(p.arc_origin)
    assert len(set(origin.tolist())) == A and origin.min() >= 0 and origin.max() < n_raw_arcs


def results_in_original_numbering(ab, p, sc):
    logz, alpha, beta, post = nb.lattice_forward_backward(p, arc_scores=sc)
    score, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=sc)
    torch.cuda.synchronize()
    so = np.concatenate([[0], np.cumsum(ab.n_states.cpu().numpy())])
    lat = np.repeat(np.arange(p.n_lattices), np.diff(_np(p.state_off)))
    g2o = so[lat] + _np(p.orig_state)
    beta_o = np.full(int(so[-1]), np.nan)
    beta_o[g2o] = beta.cpu().numpy()
    post_o = np.zeros(ab.src.numel())
    post_o[_np(p.arc_origin)] = post.cpu().numpy()
    offc, lab = _np(off), _np(labels)
    return logz.cpu().numpy(), beta_o, post_o, score.cpu().numpy(), [lab[offc[b]:offc[b + 1]].tolist() for b in range(p.n_lattices)]


@pytest.mark.parametrize("gen", ["translit32", "snips8", "dag_small", "translit_ties"])
def test_device_packer_matches_tensor_op_packer_and_oracle(gen, monkeypatch):
    ab = {"translit32": lambda: synth.transliteration_batch(32, seed=0), "snips8": lambda: synth.snips_batch(8, seed=1),
          "dag_small": lambda: synth.random_dag_batch(5, 3000, levels=40, seed=4),  # 19 states per level: not a tile-stream lattice
          "translit_ties": lambda: synth.transliteration_batch(16, seed=4, integer_scores=True)}[gen]()
    abd = ab.to(DEV)
    before = P.launch_count
    p_dev, sc_dev = abd.pack()
    assert P.launch_count == before + 4, "the device packer ran"
    assert all(g.small_max_arcs > 0 for g in p_dev.groups)
    check_structure(p_dev, ab.src.numel())
    monkeypatch.setattr(P, "DEVICE_PACK", 0)
    p_ref, sc_ref = abd.pack()
    assert P.launch_count == before + 4
    assert (p_dev.n_states, p_dev.n_arcs, p_dev.max_levels) == (p_ref.n_states, p_ref.n_arcs, p_ref.max_levels)
    np.testing.assert_array_equal(_np(p_dev.n_levels), _np(p_ref.n_levels))
    r_dev = results_in_original_numbering(ab, p_dev, sc_dev)
    r_ref = results_in_original_numbering(ab, p_ref, sc_ref)
    np.testing.assert_allclose(r_dev[0], r_ref[0], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(r_dev[1], r_ref[1], rtol=2e-6, atol=2e-6)  # NaN = trimmed, in the same places
    np.testing.assert_allclose(r_dev[2], r_ref[2], rtol=2e-5, atol=1e-7)
    assert np.array_equal(r_dev[3].view(np.uint32), r_ref[3].view(np.uint32)) and r_dev[4] == r_ref[4], "Viterbi is bit-exact"
    ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(), ab.n_states.numpy())
    o_logz, _, o_beta, o_post = c_oracle.forward_backward(ob)
    o_score, _, o_labels = c_oracle.viterbi(ob)
    np.testing.assert_allclose(r_dev[0], o_logz, rtol=1e-5, atol=1e-5)
    assert np.all(np.abs(r_dev[2] - o_post) <= 1e-5 * o_post + 1e-7)
    assert np.array_equal(r_dev[3].view(np.uint32), o_score.view(np.uint32))
    assert r_dev[4] == [list(x) for x in o_labels]


def test_device_packer_dense_tables_trim_padding_and_unreachable_rows(monkeypatch):
    rng = np.random.default_rng(3)
    tabs, S = [], 0
    for i in range(6):
        em, tr = random_mark_lattice(rng, 6 + 5 * i, 24)  # state ids shuffled: not topological
        tabs.append(tr)
        S = max(S, tr.shape[0])
    V = tabs[0].shape[1]
    from tests.lattice_gen import PAD
    batch = np.full((len(tabs), S + 2, V), PAD, dtype=np.int64)  # collate(): pad rows hold the pad id everywhere (quirk Q5)
    for i, t in enumerate(tabs):
        batch[i, :t.shape[0]] = t
    tr = torch.from_numpy(batch).to(DEV)
    before = P.launch_count
    p = nb.pack_dense(tr != 0, tr)
    assert P.launch_count == before + 4
    row, lab, dst = P.dense_arcs(tr)
    check_structure(p, 10**12)
    # arc_origin holds dense cell indices of real arcs; the pad rows' arcs are gone
    cells = _np(p.arc_origin)
    b, rem = cells // ((S + 2) * V), cells % ((S + 2) * V)
    s, l = rem // V, rem % V
    for i, t in enumerate(tabs):
        assert np.all(s[b == i] < t.shape[0])
    np.testing.assert_array_equal(batch[b, s, l], _np(p.orig_state)[_np(p.dst_out)])
    np.testing.assert_array_equal(l, _np(p.label_out))
    theta = torch.randn(V, device=DEV)
    beta = nb.compute_beta(tr != 0, tr, theta, k=1)
    monkeypatch.setattr(P, "DEVICE_PACK", 0)
    beta_ref = nb.compute_beta(tr != 0, tr, theta, k=1)
    assert P.launch_count == before + 8
    torch.testing.assert_close(beta, beta_ref, rtol=2e-6, atol=0)


def test_device_packer_rejects_cycles_and_bad_endpoints():
    lat = torch.zeros(3, dtype=torch.int64, device=DEV)
    src, dst = torch.tensor([0, 1, 2], device=DEV), torch.tensor([1, 2, 1], device=DEV)
    lab = torch.tensor([4, 5, 6], device=DEV)
    with pytest.raises(ValueError, match="cyclic"):
        nb.pack_arcs(lat, src, dst, lab, torch.tensor([3]), 16)
    tr = torch.zeros((1, 3, 8), dtype=torch.int64, device=DEV)
    tr[0, 0, 4] = 7  # points outside the table
    with pytest.raises(ValueError, match="outside the table|out of range"):
        nb.pack_dense(tr != 0, tr)


def test_device_packer_falls_back_for_wide_lattices():
    before = P.launch_count
    p, _ = synth.random_dag_batch(2, 60_000, levels=16, seed=1, device=DEV).pack()
    assert p.has_tiles and P.launch_count == before, "wide lattices keep the column-major layouts of the tensor-op packer"


def test_tensor_op_packer_levels_on_the_device_and_rejects_cycles():
    # tiles=True keeps the batch on the tensor-op packer, whose level sweeps run in the library (nfst_level_sweeps)
    ab = synth.random_dag_batch(3, 5000, levels=10, seed=2)
    p_gpu, _ = ab.to(DEV).pack(tiles=True)
    p_cpu, _ = ab.pack(tiles=True)  # the same packer on the CPU: torch scatter sweeps
    assert p_gpu.has_tiles
    np.testing.assert_array_equal(_np(p_gpu.level_ptr), _np(p_cpu.level_ptr))
    np.testing.assert_array_equal(_np(p_gpu.orig_state), _np(p_cpu.orig_state))
    np.testing.assert_array_equal(_np(p_gpu.dst_out), _np(p_cpu.dst_out))
    assert torch.equal(p_gpu.tile_stream.cpu(), p_cpu.tile_stream)
    n = 200  # a long cycle 1 -> 2 -> ... -> n -> 1 behind the start state
    src = torch.arange(0, n + 1, device=DEV)
    dst = torch.cat([torch.arange(1, n + 1, device=DEV), torch.tensor([1], device=DEV)])
    with pytest.raises(ValueError, match="cyclic"):
        nb.pack_arcs(torch.zeros(n + 1, dtype=torch.int64, device=DEV), src, dst, torch.full((n + 1,), 5, device=DEV),
                     torch.tensor([n + 1]), 16, tiles=True)


def test_c_abi_pack_dense_equals_the_python_path():
    """nfst_pack_dense: dense tables -> packed arrays through the C ABI alone, both phases, against
    nb.pack_dense; an arc list that overflows its capacity is reported through totals[4] = 5."""
    import ctypes as C

    from nfst_b200 import _lib

    rng = np.random.default_rng(5)
    tabs = [random_mark_lattice(rng, 4 + 3 * i, 20)[1] for i in range(5)]
    S, V = max(t.shape[0] for t in tabs) + 1, 20
    from tests.lattice_gen import PAD
    batch = np.full((len(tabs), S, V), PAD, dtype=np.int64)
    for i, t in enumerate(tabs):
        batch[i, :t.shape[0]] = t
    tr = torch.from_numpy(batch).to(DEV)
    ref = nb.pack_dense(tr != 0, tr)
    lib = _lib.load()
    B = len(tabs)
    cap = int(((tr != 0) & (tr != torch.arange(S, device=DEV).view(1, S, 1))).sum())

    def run(capacity):
        i32 = dict(dtype=torch.int32, device=DEV)
        o = {k: torch.zeros(n, **i32) for k, n in (("state_off", B + 1), ("arc_off", B + 1), ("level_off", B + 1), ("sink_off", B + 1),
                                                   ("n_levels", B), ("start_state", B), ("lattice_stats", 8 * B), ("totals", 8))}
        out = _lib.PackOutC()
        for k, t in o.items():
            setattr(out, k, t.data_ptr())
        nbytes = int(lib.nfst_pack_workspace_bytes(B, S, capacity))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.nfst_pack_dense(tr.data_ptr(), B, S, V, capacity, 4096, C.byref(out), ws.data_ptr(), nbytes, 1, st))
        totals = o["totals"].cpu().tolist()
        if totals[4]:
            return totals, None
        Sk, A, n_lp, n_sink = totals[:4]
        big = {k: torch.zeros(n + 4, **i32) for k, n in (("level_ptr", n_lp), ("sinks", n_sink), ("orig_state", Sk), ("in_ptr", Sk + 1),
                                                         ("out_ptr", Sk + 1), ("src_in", A), ("label_in", A), ("in2out", A),
                                                         ("dst_out", A), ("label_out", A), ("src_out", A))}
        origin = torch.zeros(A, dtype=torch.int64, device=DEV)
        for k, t in big.items():
            setattr(out, k, t.data_ptr())
        out.arc_origin = origin.data_ptr()
        _lib.check(lib.nfst_pack_dense(tr.data_ptr(), B, S, V, capacity, 4096, C.byref(out), ws.data_ptr(), nbytes, 2, st))
        torch.cuda.synchronize()
        return totals, {**{k: v for k, v in o.items()}, **{k: v[:-4] for k, v in big.items()}, "arc_origin": origin}

    totals, got = run(cap)
    assert totals[4] == 0 and (totals[0], totals[1]) == (ref.n_states, ref.n_arcs)
    for k in ("state_off", "arc_off", "level_off", "sink_off", "n_levels", "start_state", "level_ptr", "sinks", "orig_state", "in_ptr",
              "out_ptr", "src_in", "label_in", "in2out", "dst_out", "label_out", "src_out", "arc_origin"):
        assert torch.equal(got[k].to(torch.int64), getattr(ref, k).to(torch.int64)), k
    totals, got = run(cap - 1)
    assert totals[4] == 5 and got is None


def test_pack_fsts_on_the_device_matches_the_oracle():
    """``data.pack_fsts``: OpenFst-shaped acceptors -> packed batch on the device (row a1 without the dense tables,
    ``scorers.py:995-1035``): logZ with theta scores and with the machines' own weights against the numpy oracle."""
    from nfst_b200 import data as nd
    from oracle import lattice_oracle as lo
    from tests.lattice_gen import PAD, FakeFst, FakeWeight, random_fst_arrays

    rng = np.random.default_rng(12)
    V = 24
    zero = FakeWeight.zero("tropical")
    machines = [FakeFst(*random_fst_arrays(rng, int(n), V)) for n in rng.integers(3, 40, size=9)]
    theta = rng.normal(size=V).astype(np.float32)
    p = nd.pack_fsts(machines, V, device=DEV, final_zero=zero)
    pw = nd.pack_fsts(machines, V, weighted=True, device=DEV, final_zero=zero)
    assert p.device.type == "cuda" and pw.static_scores is not None
    logz = nb.lattice_log_partition(p, theta=torch.from_numpy(theta).to(DEV)).cpu().numpy()
    logz_w = nb.lattice_log_partition(pw).cpu().numpy()  # the static scores (-arc.weight) are the arc scores
    for b, m in enumerate(machines):
        em, tr = nd.get_state_mask_pynini(m, V, PAD, to_numpy=True, weighted=True, final_zero=zero)
        s, l, d, sc = lo.arcs_from_dense(tr, em)
        want = lo.forward_backward(tr.shape[0], s, d, theta[l].astype(np.float64))[0]
        want_w = lo.forward_backward(tr.shape[0], s, d, sc.astype(np.float32).astype(np.float64))[0]
        assert abs(logz[b] - want) <= 1e-5 * max(1.0, abs(want)), (b, logz[b], want)
        assert abs(logz_w[b] - want_w) <= 1e-5 * max(1.0, abs(want_w)), (b, logz_w[b], want_w)
