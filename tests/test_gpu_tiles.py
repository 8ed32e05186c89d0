"""GPU parity of the tile-stream kernels (nfst_tiles.cu) against the CPU oracle: log-partition, beta, posteriors
(fixed-point conditional-probability flow), autograd in both score modes, Viterbi bit-exact with forced ties,
every warp count, far arcs (small ring), heavy states cut into pieces, float64 state, mixed batches, stage depths,
and bit-reproducibility of the posteriors."""
import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import synth
from nfst_b200 import tiles as T
from nfst_b200.pack import concat_packed
from oracle import c_oracle
from tests.test_gpu_parity import DEV, check_fwd_bwd, oracle_batch, post_rtol

pytestmark = pytest.mark.gpu


def viterbi_matches(ab, p, sc):
    score, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=sc)
    o_score, o_paths, o_labels = c_oracle.viterbi(oracle_batch(ab))
    assert np.array_equal(score.cpu().numpy().view(np.uint32), o_score.view(np.uint32))
    origin, offc, arcs_c, lab = p.arc_origin.cpu().numpy(), off.cpu().numpy(), arcs.cpu().numpy(), labels.cpu().numpy()
    for b in range(p.n_lattices):
        np.testing.assert_array_equal(origin[arcs_c[offc[b]:offc[b + 1]]], o_paths[b])
        np.testing.assert_array_equal(lab[offc[b]:offc[b + 1]], o_labels[b])


@pytest.mark.parametrize("arcs,levels,B,warps", [(3_000, 8, 5, 0), (10_000, 64, 6, 0), (100_000, 64, 4, 0), (100_000, 64, 3, 1),
                                                 (100_000, 64, 3, 4), (400_000, 64, 2, 0), (400_000, 64, 2, 16),
                                                 (20_000, 16, 4, 16)])
def test_tile_forward_backward_and_viterbi(arcs, levels, B, warps, monkeypatch):
    monkeypatch.setattr(T, "TILE_WARPS", warps)
    ab = synth.random_dag_batch(B, arcs, levels=levels, seed=7)
    p, sc, _ = check_fwd_bwd(ab, strict=True)  # 1e-5 flat, no depth scaling
    assert all(g.tiles for g in p.groups)
    viterbi_matches(ab, p, sc)


@pytest.mark.parametrize("stages", [2, 3, 5])
def test_tile_stage_depths(stages, monkeypatch):
    monkeypatch.setattr(nb.ops, "TILE_STAGES", stages)
    ab = synth.random_dag_batch(3, 30_000, levels=24, seed=8)
    p, sc, _ = check_fwd_bwd(ab, strict=True)
    assert p.has_tiles
    viterbi_matches(ab, p, sc)


def test_tile_small_ring_far_arcs(monkeypatch):
    monkeypatch.setattr(T, "FORCE_RING_SLICES", 20)
    monkeypatch.setattr(T, "TILE_WARPS", 2)
    ab = synth.random_dag_batch(3, 12_000, levels=24, seed=5)
    p, sc, _ = check_fwd_bwd(ab, strict=True)
    assert p.has_tiles and any(g.tile_far for g in p.groups)
    viterbi_matches(ab, p, sc)
    # autograd through the far table
    w = sc.clone().requires_grad_(True)
    nb.lattice_log_partition(p, arc_scores=w).sum().backward()
    _, _, _, o_post = c_oracle.forward_backward(oracle_batch(ab))
    ref = o_post[p.arc_origin.cpu().numpy()]
    assert np.all(np.abs(w.grad.cpu().numpy() - ref) <= 1e-5 * ref + 1e-7)


def test_tile_heavy_states_in_pieces(monkeypatch):
    monkeypatch.setattr(T, "TILE_ARCS", 64)
    ab = synth.random_dag_batch(3, 8_000, levels=6, seed=9)  # sources with hundreds of arcs: several pieces
    p, sc, _ = check_fwd_bwd(ab, strict=True)
    assert p.has_tiles
    viterbi_matches(ab, p, sc)


def test_tile_extension_columns_and_theta():
    # states with 1..33 arcs (extension blocks, one heavy state) in theta mode and in per-arc mode
    n = 300
    src = [torch.zeros(n, dtype=torch.int64)]
    dst = [torch.arange(1, n + 1)]
    lab = [torch.arange(n) % 250 + 4]
    for s in range(1, n + 1):
        k = 1 + s % 33
        src.append(torch.full((k,), s)); dst.append(n + 1 + (torch.arange(k) * 7 + s) % n); lab.append(torch.arange(k) + 4)
    src.append(torch.arange(n + 1, 2 * n + 1)); dst.append(torch.full((n,), 2 * n + 1)); lab.append(torch.full((n,), 3))
    src, dst, lab = torch.cat(src), torch.cat(dst), torch.cat(lab)
    g = torch.Generator().manual_seed(0)
    sc = torch.randn(src.numel(), generator=g)
    ab = synth.ArcBatch(torch.zeros(src.numel(), dtype=torch.int64), src, dst, lab, sc, torch.tensor([2 * n + 2]), 256)
    p, scp, _ = check_fwd_bwd(ab, strict=True)
    assert p.has_tiles
    viterbi_matches(ab, p, scp)
    theta = torch.randn(256, generator=g)
    ab2 = synth.ArcBatch(ab.arc_lattice, src, dst, lab, theta[lab], ab.n_states, 256)
    o_logz, _, _, o_post = c_oracle.forward_backward(oracle_batch(ab2))
    th = theta.to(DEV).requires_grad_(True)
    logz = nb.lattice_log_partition(p, theta=th)
    logz.sum().backward()
    np.testing.assert_allclose(logz.detach().cpu().numpy(), o_logz, rtol=1e-6)
    dth = np.zeros(256)
    np.add.at(dth, lab.numpy(), o_post)
    np.testing.assert_allclose(th.grad.cpu().numpy(), dth, rtol=1e-5, atol=1e-6)
    score, off, arcs, labels = nb.lattice_viterbi(p, theta=th.detach())
    o_score, o_paths, o_labels = c_oracle.viterbi(oracle_batch(ab2))
    assert np.array_equal(score.cpu().numpy().view(np.uint32), o_score.view(np.uint32))
    np.testing.assert_array_equal(labels.cpu().numpy(), o_labels[0])


def test_tile_float64_state_and_deep_lattice():
    ab = synth.random_dag_batch(3, 60_000, levels=400, seed=9)  # 400 levels: "auto" returns float64 state
    p, sc, _ = check_fwd_bwd(ab)
    assert p.has_tiles and nb.ops.resolve_state_dtype(p) == torch.float64
    # float32 state: the ring holds float32 values of magnitude ~600 (ulp 6e-5): depth-scaled tolerance
    check_fwd_bwd(ab, state_dtype=torch.float32)


def test_tile_viterbi_forced_ties():
    ab = synth.random_dag_batch(4, 20_000, levels=32, seed=11)
    g = torch.Generator().manual_seed(0)
    ab.scores = -torch.randint(0, 3, ab.scores.shape, generator=g).float()  # {-2,-1,0}: many exact ties
    p, sc = ab.to(DEV).pack()
    assert p.has_tiles
    viterbi_matches(ab, p, sc)


def test_tile_agrees_with_csr_kernels_and_is_bit_reproducible(monkeypatch):
    ab = synth.random_dag_batch(4, 50_000, levels=32, seed=13).to(DEV)
    pt, sct = ab.pack()
    monkeypatch.setattr(T, "TILES", 0)
    monkeypatch.setattr(nb.pack, "SELL", 0)
    pc, scc = ab.pack()
    assert pt.has_tiles and not pc.has_columns
    lt, at, bt, postt = nb.lattice_forward_backward(pt, arc_scores=sct)
    lc, ac, bc, postc = nb.lattice_forward_backward(pc, arc_scores=scc)
    assert torch.allclose(lt, lc, rtol=1e-6, atol=1e-4)
    a = torch.empty_like(postt); a[pt.arc_origin] = postt
    c = torch.empty_like(postc); c[pc.arc_origin] = postc
    assert torch.allclose(a, c, rtol=3e-5, atol=1e-7)
    vt = nb.lattice_viterbi(pt, arc_scores=sct)
    vc = nb.lattice_viterbi(pc, arc_scores=scc)
    assert torch.equal(vt[0], vc[0]) and torch.equal(vt[1], vc[1])
    assert torch.equal(pt.arc_origin[vt[2].long()], pc.arc_origin[vc[2].long()]) and torch.equal(vt[3], vc[3])
    # integer atomics: the flow is the same bit for bit however the warps interleave
    for _ in range(3):
        again = nb.lattice_forward_backward(pt, arc_scores=sct)[3]
        assert torch.equal(again, postt)


def test_tile_autograd_arc_scores_and_theta():
    ab = synth.random_dag_batch(5, 8_000, levels=16, seed=17)
    p, sc = ab.to(DEV).pack()
    assert p.has_tiles
    o_logz, _, _, o_post = c_oracle.forward_backward(oracle_batch(ab))
    origin = p.arc_origin.cpu().numpy()
    w = sc.clone().requires_grad_(True)
    coef = torch.linspace(0.5, 2.0, p.n_lattices, device=DEV)
    logz = nb.lattice_log_partition(p, arc_scores=w)
    (logz * coef).sum().backward()
    np.testing.assert_allclose(logz.detach().cpu().numpy(), o_logz, rtol=1e-6)
    arc_lat = np.repeat(np.arange(p.n_lattices), np.diff(p.arc_off.cpu().numpy()))
    ref = o_post[origin] * coef.cpu().numpy()[arc_lat]
    got = w.grad.cpu().numpy()
    assert np.all(np.abs(got - ref) <= 1e-5 * ref + 2e-7)
    theta = torch.randn(p.vocab, device=DEV, requires_grad=True)
    nb.lattice_log_partition(p, theta=theta).sum().backward()
    wt = theta.detach().cpu().numpy()[ab.label.numpy()]
    ab2 = synth.ArcBatch(ab.arc_lattice, ab.src, ab.dst, ab.label, torch.from_numpy(wt), ab.n_states, ab.vocab)
    _, _, _, po = c_oracle.forward_backward(oracle_batch(ab2))
    dth = np.zeros(p.vocab)
    np.add.at(dth, ab.label.numpy(), po)
    np.testing.assert_allclose(theta.grad.cpu().numpy(), dth, rtol=1e-5, atol=1e-5)
    logz2, alpha, beta, post, dtheta = nb.lattice_forward_backward(p, arc_scores=sc, theta=theta.detach(), want_dtheta=True)
    wb = (ab.scores.numpy() + wt).astype(np.float32)
    ab3 = synth.ArcBatch(ab.arc_lattice, ab.src, ab.dst, ab.label, torch.from_numpy(wb), ab.n_states, ab.vocab)
    o3_logz, _, _, po3 = c_oracle.forward_backward(oracle_batch(ab3))
    np.testing.assert_allclose(logz2.cpu().numpy(), o3_logz, rtol=1e-5)
    d3 = np.zeros(p.vocab)
    np.add.at(d3, ab.label.numpy(), po3)
    np.testing.assert_allclose(dtheta.cpu().numpy(), d3, rtol=1e-5, atol=1e-5)


def test_tile_mixed_batch_and_lattice_backward_outputs():
    parts = [synth.transliteration_batch(5, seed=2), synth.random_dag_batch(3, 60_000, levels=16, seed=4),
             synth.snips_batch(4, seed=1), synth.random_dag_batch(2, 6_000, levels=48, seed=6)]
    packs, scores = zip(*[ab.to(DEV).pack() for ab in parts])
    p = concat_packed(list(packs))
    sc = torch.cat(scores)
    assert {g.tiles for g in p.groups} == {True, False}
    logz, alpha, beta, post = nb.lattice_forward_backward(p, arc_scores=sc)
    off_a = off_b = 0
    for ab, pk in zip(parts, packs):
        o_logz, o_alpha, o_beta, o_post = c_oracle.forward_backward(oracle_batch(ab))
        np.testing.assert_allclose(logz[off_b:off_b + pk.n_lattices].cpu().numpy(), o_logz, rtol=1e-5)
        ref = o_post[pk.arc_origin.cpu().numpy()]
        got = post[off_a:off_a + pk.n_arcs].cpu().numpy().astype(np.float64)
        rt = 1e-5 if pk.has_tiles else post_rtol(float(np.abs(o_alpha).max() + np.abs(o_beta).max()))
        assert np.all(np.abs(got - ref) <= rt * ref + 1e-7), (rt, float(np.max(np.abs(got - ref) / (ref + 1e-7))), pk.n_arcs)
        off_a += pk.n_arcs
        off_b += pk.n_lattices
    alpha2, logz2 = nb.lattice_forward(p, arc_scores=sc)
    assert torch.allclose(alpha2, alpha, rtol=1e-5, atol=1e-4) and torch.allclose(logz2, logz, rtol=1e-6, atol=1e-5)
    r = nb.ops.lattice_backward(p, sc, alpha=alpha2, logz=logz2, want_beta=True, want_post=True, want_viterbi=True)
    assert torch.allclose(r["beta"], beta, rtol=1e-6, atol=1e-5)
    assert torch.allclose(r["post"], post, rtol=1e-5, atol=1e-7)
    vs, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=sc)
    assert torch.equal(r["vit_score"], vs)


def test_viterbi_padded_rows_equal_the_ragged_result():
    """``lattice_viterbi_padded`` (no host read; the layout of the reference's best sample, lightning.py:474-479) against
    ``lattice_viterbi`` on a batch that mixes small-lattice, tile-stream and CSR groups; with and without a skipped first label."""
    parts = [synth.transliteration_batch(6, seed=2), synth.random_dag_batch(2, 6_000, levels=12, seed=4), synth.snips_batch(3, seed=1)]
    packs, scores = zip(*[ab.to(DEV).pack() for ab in parts])
    p = concat_packed(list(packs))
    sc = torch.cat(scores)
    vs, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=sc)
    PADL = 3
    first = int(labels[int(off[0])])
    for skip in (-1, first):
        vs2, rows, lens = nb.lattice_viterbi_padded(p, arc_scores=sc, pad_label=PADL, skip_label=skip)
        assert torch.equal(vs, vs2) and rows.shape == (p.n_lattices, p.max_levels - 1) and rows.dtype == torch.int64
        offc, lab, rows_c, lens_c = off.cpu().tolist(), labels.cpu(), rows.cpu(), lens.cpu().tolist()
        for b in range(p.n_lattices):
            want = lab[offc[b]:offc[b + 1]].to(torch.int64)
            if skip >= 0 and want.numel() and int(want[0]) == skip:
                want = want[1:]
            assert lens_c[b] == want.numel() and torch.equal(rows_c[b, :lens_c[b]], want)
            assert bool((rows_c[b, lens_c[b]:] == PADL).all())


def test_tile_large_ring_leaves_no_room_for_a_second_staged_array():
    """1M-arc lattices: the DP ring fills the shared memory, so (a) the flow pass reads the labels of the dtheta
    histogram from global memory instead of staging them and (b) per-arc scores + theta are summed once before the
    pull pass.  Both against the C oracle."""
    ab = synth.random_dag_batch(2, 1_000_000, seed=5)
    p, sc = ab.to(DEV).pack()
    assert p.has_tiles and nb.ops._two_arrays_do_not_fit(p)
    g = torch.Generator().manual_seed(1)
    theta = torch.randn(p.vocab, generator=g) * 0.3
    lab = ab.label.numpy()
    for with_scores in (False, True):
        w = theta.numpy()[lab] + (ab.scores.numpy() if with_scores else 0.0)
        ab2 = synth.ArcBatch(ab.arc_lattice, ab.src, ab.dst, ab.label, torch.from_numpy(w.astype(np.float32)), ab.n_states, ab.vocab)
        o_logz, _, _, o_post = c_oracle.forward_backward(oracle_batch(ab2))
        logz, _, _, post, dth = nb.lattice_forward_backward(p, arc_scores=sc if with_scores else None, theta=theta.to(DEV),
                                                            want_dtheta=True)
        np.testing.assert_allclose(logz.cpu().numpy(), o_logz, rtol=1e-6)
        ref = o_post[p.arc_origin.cpu().numpy()]
        assert np.all(np.abs(post.cpu().numpy() - ref) <= 1e-5 * ref + 2e-9)
        want = np.zeros(p.vocab)
        np.add.at(want, lab, o_post)
        np.testing.assert_allclose(dth.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
        score, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=sc if with_scores else None, theta=theta.to(DEV))
        o_score, o_paths, _ = c_oracle.viterbi(oracle_batch(ab2))
        assert np.array_equal(score.cpu().numpy().view(np.uint32), o_score.view(np.uint32))


def test_deep_and_wide_lattice_runs_with_float64_state():
    """More than 96 levels: float64 state by default, i.e. 8-byte ring slots -- the packer sizes the ring for them (or
    leaves the lattice to the level-major kernels); the call must not be refused for shared memory."""
    ab = synth.random_dag_batch(1, 2_000_000, levels=128, seed=2)
    p, sc, _ = check_fwd_bwd(ab)
    assert nb.ops.resolve_state_dtype(p) == torch.float64
    for g in p.groups:
        assert not g.tiles or (g.tile_ring + 32) * 8 < 200 * 1024
    viterbi_matches(ab, p, sc)


def test_tile_rejects_misuse_and_misalignment():
    ab = synth.random_dag_batch(2, 5_000, levels=8, seed=3).to(DEV)
    p, sc = ab.pack()
    assert p.has_tiles
    lib = nb._lib.load()
    lc = nb.ops._launch(p.groups[0], torch.float32)
    scs, keep = nb.ops._scores(p, sc, None)
    a = torch.empty(p.n_states, device=DEV)
    z = torch.empty(p.n_lattices, device=DEV)
    rc = lib.nfst_fwd_f32(p.c_struct(), lc, scs, a.data_ptr(), z.data_ptr(), None)
    assert rc < 0 and b"tile-stream" in lib.nfst_last_error_string()
    rc = lib.nfst_tile_flow_f32(p.c_struct(), lc, None, None, None, None, None)
    assert rc < 0
    logz, _, cond = nb.ops.lattice_pull(p, arc_scores=sc)
    bad = torch.empty(p.n_arcs + 1, device=DEV)[1:]
    bad.copy_(cond)
    with pytest.raises(RuntimeError, match="16-byte aligned"):
        nb.ops.lattice_backward(p, sc, logz=logz, cond=bad, want_beta=False, want_post=True)
    # a misaligned view of the scores: the operator copies it once
    view = torch.empty(p.n_arcs + 1, device=DEV)[1:]
    view.copy_(sc)
    l2 = nb.lattice_log_partition(p, arc_scores=view)
    assert torch.equal(l2, logz)


def test_tile_no_finite_path_is_minus_infinity():
    """A -inf score that blocks every path: logZ = -inf, posteriors 0, no NaN; Viterbi score -inf with a valid
    (possibly empty) path and no out-of-range backpointer."""
    ab = synth.random_dag_batch(3, 4_000, levels=8, seed=21)
    p, sc = ab.to(DEV).pack()
    assert p.has_tiles
    w = sc.clone()
    a0, a1 = int(p.arc_off[0]), int(p.arc_off[1])
    src = p.src_out[a0:a1]
    w[a0:a1][src == p.start_state[0]] = float("-inf")  # lattice 0: every arc out of the start state
    logz, alpha, beta, post = nb.lattice_forward_backward(p, arc_scores=w)
    assert torch.isinf(logz[0]) and logz[0] < 0 and torch.isfinite(logz[1:]).all()
    assert not torch.isnan(post).any() and float(post[a0:a1].abs().max()) == 0.0
    score, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=w)
    assert torch.isinf(score[0]) and score[0] < 0 and torch.isfinite(score[1:]).all()
    assert int(arcs.max()) < p.n_arcs and int(arcs.min()) >= 0


@pytest.mark.parametrize("layout", ["tiles", "sell"])
def test_single_state_consumers_read_column_layouts(layout, monkeypatch):
    """The sampling-loop kernels and the beta-hat recurrence visit the arcs of ONE state; on column-major lattices
    they find them through ``out_arc``.  Same lattices packed CSR and column-major must give the same walk step,
    the same exact samples (same uniforms) and the same beta-hat."""
    monkeypatch.setattr(nb.pack, "SELL_MIN_WIDTH", 32)
    ab = synth.random_dag_batch(3, 5_000, levels=12, seed=23).to(DEV)
    monkeypatch.setattr(T, "TILES", int(layout == "tiles"))
    monkeypatch.setattr(nb.pack, "SELL", int(layout == "sell"))
    pcol, sccol = ab.pack()
    monkeypatch.setattr(T, "TILES", 0)
    monkeypatch.setattr(nb.pack, "SELL", 0)
    pcsr, sccsr = ab.pack()
    assert pcol.has_columns and not pcsr.has_columns and pcol.out_arc.numel() == pcol.n_arcs
    # out_arc lists every state's arcs in label order
    oa, optr = pcol.out_arc.long(), pcol.out_ptr.long()
    src = torch.repeat_interleave(torch.arange(pcol.n_states, device=DEV), optr[1:] - optr[:-1])
    assert torch.equal(pcol.src_out[oa].long(), src)
    lab = pcol.label_out[oa]
    same_state = src[1:] == src[:-1]
    assert bool((lab[1:][same_state] >= lab[:-1][same_state]).all())  # random DAGs repeat labels
    # states are numbered differently in the two packs: compare through original ids
    g = torch.Generator(device=DEV).manual_seed(5)
    k = 4
    N = pcol.n_lattices * k
    T_ = max(pcol.max_levels - 1, 1)
    u = torch.rand(N, T_, device=DEV, generator=g)
    la, ln, lq, arcs_a, lz_a = nb.sample_paths(pcol, k, arc_scores=sccol, uniform=u)
    lb, lnb, lqb, arcs_b, lz_b = nb.sample_paths(pcsr, k, arc_scores=sccsr, uniform=u)
    assert torch.equal(la, lb) and torch.equal(ln, lnb)
    assert torch.allclose(lq, lqb, rtol=1e-5, atol=1e-5) and torch.allclose(lz_a, lz_b, rtol=1e-6, atol=1e-5)
    valid = arcs_a >= 0
    assert torch.equal(pcol.arc_origin[arcs_a[valid].long()], pcsr.arc_origin[arcs_b[valid].long()])
    # one walk step from the start states with a random prefix, sampling and scoring
    V = pcol.vocab
    prefix = torch.randn(N, V, device=DEV, generator=g)
    rb = nb.ops.lattice_backward(pcsr, sccsr, want_beta=True)
    # the same look-ahead values for both walks (real-space beta amplifies the kernels' rounding differences): the CSR
    # pack's beta carried over to the column-major pack's state numbering through the original state ids
    ia = torch.argsort(pcol.state_off[:-1].long().repeat_interleave((pcol.state_off[1:] - pcol.state_off[:-1]).long()) * 10**7 + pcol.orig_state.long())
    ib = torch.argsort(pcsr.state_off[:-1].long().repeat_interleave((pcsr.state_off[1:] - pcsr.state_off[:-1]).long()) * 10**7 + pcsr.orig_state.long())
    beta_b = rb["beta"].float().exp()
    beta_a = torch.empty_like(beta_b)
    beta_a[ia] = beta_b[ib]
    wa = nb.LatticeWalker(pcol, k, beta_a, pad_id=3, faithful=False)
    wb = nb.LatticeWalker(pcsr, k, beta_b, pad_id=3, faithful=False)
    uu = torch.rand(N, device=DEV, generator=g)
    for _ in range(3):
        sa, pa, za = wa.step(prefix, uniform=uu)
        sb, pb, zb = wb.step(prefix, uniform=uu)
        assert torch.equal(sa, sb) and torch.allclose(pa, pb, rtol=1e-5, atol=1e-5) and torch.allclose(za, zb, rtol=1e-5, atol=1e-5)
        assert torch.equal(wa.dense_state(), wb.dense_state())
    # the beta-hat recurrence (Wh != 0)
    H = 16
    proj = torch.randn(V, H, device=DEV, generator=g) * 0.3
    Wh = torch.randn(H, H, device=DEV, generator=g) * 0.2
    W = torch.randn(H, device=DEV, generator=g) * 0.3
    lba, bha = nb.lattice_beta_hat(pcol, proj, Wh, W)
    lbb, bhb = nb.lattice_beta_hat(pcsr, proj, Wh, W)
    assert torch.allclose(lba[ia], lbb[ib], rtol=1e-5, atol=1e-5) and torch.allclose(bha[ia], bhb[ib], rtol=0, atol=1e-5)


def test_stripping_pad_matches_the_reference_loop():
    """``Sampler.stripping_pad`` (samplers.py:162-180) restated with its own loop (torch, CPU) against the kernels:
    zeros inside and at the end of rows, rows of different lengths, a batch without an all-pad column."""
    def reference(sequences, pad):
        batch_size, seq_len = sequences.shape
        to_return = torch.full_like(sequences, pad)
        indices = torch.zeros(batch_size, dtype=torch.long)
        for i in range(seq_len):
            to_return[torch.arange(batch_size), indices] = sequences[:, i]
            indices = indices + (sequences[:, i] != 0).long()
            if torch.all(sequences[:, i] == pad):
                break
        return to_return[:, : i + 1].contiguous()

    g = torch.Generator().manual_seed(3)
    pad = 3
    for N, T_, with_pad_tail in ((7, 19, True), (64, 301, True), (5, 40, False), (1, 1, False), (33, 70, True)):
        seq = torch.randint(0, 9, (N, T_), generator=g)
        seq[torch.rand(N, T_, generator=g) < 0.3] = 0
        if with_pad_tail:
            lens = torch.randint(1, T_, (N,), generator=g)
            seq[torch.arange(T_)[None, :] >= lens[:, None]] = pad
        want = reference(seq.clone(), pad)
        got = nb.stripping_pad(seq.to(DEV), pad).cpu()
        assert got.shape == want.shape and torch.equal(got, want), (N, T_)


def test_exact_joint_prob_adapter(tmp_path):
    """``ExactJointProb.forward`` / ``decode_from_npz`` return the reference's shapes (lightning.py:442-480, 647-658):
    num_prob[B] = logZ - theta[bos] (checked against the oracle), denom_prob = 0, the best path without bos."""
    from nfst_b200.data import LatticeDataset, collate
    from oracle import lattice_oracle as lo
    from tests.lattice_gen import BOS, EOS, PAD, random_mark_lattice

    rng = np.random.default_rng(11)
    V = 24
    theta = torch.from_numpy(rng.normal(size=V).astype(np.float32))
    tables, names = [], []
    for i in range(4):
        em, tr = random_mark_lattice(rng, int(rng.integers(4, 14)), V)
        tables.append(tr)
        name = str(tmp_path / f"ex{i}")
        np.savez_compressed(name + ".npz", num_emission=em, num_transition=tr, denom_emission=em, denom_transition=tr,
                            gs=np.arange(3 + i), ps=np.arange(2 + i))
        names.append(name)
    S = max(t.shape[0] for t in tables)
    tr_b = np.stack([lo.collate_pad([t], PAD)[0] if t.shape[0] == S else lo.collate_pad([t, np.zeros((S, V), dtype=np.int64)], PAD)[0] for t in tables])
    model = nb.ExactJointProb(theta.to(DEV), bos=BOS, eos=EOS, pad=PAD)
    trt = torch.from_numpy(tr_b).to(DEV)
    num, den, best = model(trt != 0, trt, None, None, None, None, return_samples=True)
    assert num.shape == (4,) and den.shape == (4,) and float(den.abs().max()) == 0.0 and best.shape[0] == 4
    for b, tr in enumerate(tables):
        s, l, d, _ = lo.arcs_from_dense(tr)
        logz, _, _, _ = lo.forward_backward(tr.shape[0], s, d, theta.numpy()[l].astype(np.float64))
        assert abs(float(num[b]) - (logz - float(theta[BOS]))) < 1e-4
        _, _, vl, _, _ = lo.viterbi_f32(tr.shape[0], s, l, d, theta.numpy()[l])
        row = best[b].cpu().tolist()
        assert row[: len(vl) - 1] == list(vl)[1:] and all(x == PAD for x in row[len(vl) - 1:])
    # one file, as decode/decoder.py:77-79 calls it; then the cached dataset + collate
    prob, mark = model.decode_from_npz(names[2] + ".npz", V, PAD)
    assert abs(prob - float(num[2])) < 1e-5 and mark.dim() == 1 and mark.tolist()[-1] == EOS
    ds = LatticeDataset(names, V, PAD)
    first = [ds[i] for i in range(4)]  # packs and writes <name>.packed.npz
    again = [ds[i] for i in range(4)]  # reads the caches
    import os
    assert all(os.path.exists(n + ".packed.npz") for n in names)
    for a, b2 in zip(first, again):
        for f in ("dst_out", "label_out", "out_ptr", "level_ptr", "arc_origin"):
            assert torch.equal(getattr(a.packed, f), getattr(b2.packed, f))
    packed, gs, ps = collate(again, PAD, device=DEV)
    assert gs.shape == (4, 6) and ps.shape == (4, 5) and int(gs[0, 3]) == PAD
    num2, den2 = model(packed, None)
    assert torch.allclose(num2, num, rtol=1e-6, atol=1e-5)
