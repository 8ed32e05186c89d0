"""GPU parity of the sliced-column kernels (nfst_sell.cu) against the CPU oracle: log-partition, beta,
alpha (from the state posterior), posteriors (conditional-probability flow), autograd in both score
modes, Viterbi bit-exact with forced ties, heavy states, mixed batches, and agreement with the CSR
kernels on the same lattices."""
import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import synth
from nfst_b200.pack import concat_packed
from oracle import c_oracle
from tests.test_gpu_parity import DEV, check_fwd_bwd, gpu_state_to_orig, oracle_batch, post_rtol

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def narrow_lattices_too(monkeypatch):
    """The packer only sends lattices of >= 96 states per level down the sliced-column path (below that the
    CSR kernels are faster); the tests want small lattices there too."""
    monkeypatch.setattr(nb.pack, "SELL_MIN_WIDTH", 32)
    monkeypatch.setattr(nb.tiles, "TILES", 0)  # these tests are about the sliced-column layout (tiles win by default)


def viterbi_matches(ab, p, sc):
    score, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=sc)
    o_score, o_paths, o_labels = c_oracle.viterbi(oracle_batch(ab))
    assert np.array_equal(score.cpu().numpy().view(np.uint32), o_score.view(np.uint32))
    origin, offc, arcs_c, lab = p.arc_origin.cpu().numpy(), off.cpu().numpy(), arcs.cpu().numpy(), labels.cpu().numpy()
    for b in range(p.n_lattices):
        np.testing.assert_array_equal(origin[arcs_c[offc[b]:offc[b + 1]]], o_paths[b])
        np.testing.assert_array_equal(lab[offc[b]:offc[b + 1]], o_labels[b])


@pytest.mark.parametrize("arcs,levels,B", [(3_000, 8, 5), (10_000, 64, 6), (100_000, 64, 4), (400_000, 64, 2)])
def test_sell_forward_backward_and_viterbi(arcs, levels, B):
    ab = synth.random_dag_batch(B, arcs, levels=levels, seed=7)
    # strict: sums and differences run in float64 inside the kernel, 1e-5 holds without depth scaling
    p, sc, _ = check_fwd_bwd(ab, strict=True)
    assert all(g.sell for g in p.groups)
    viterbi_matches(ab, p, sc)


def test_sell_float64_state_and_deep_lattice():
    ab = synth.random_dag_batch(3, 60_000, levels=400, seed=9)  # 400 levels: "auto" returns float64 state
    p, sc, (logz, alpha, beta, post) = check_fwd_bwd(ab)
    assert p.has_sell and nb.ops.resolve_state_dtype(p) == torch.float64
    # float32 state: the ring holds float32 values of magnitude ~600 (ulp 6e-5): depth-scaled tolerance
    check_fwd_bwd(ab, state_dtype=torch.float32)


def test_sell_viterbi_forced_ties():
    ab = synth.random_dag_batch(4, 20_000, levels=32, seed=11)
    g = torch.Generator().manual_seed(0)
    ab.scores = -torch.randint(0, 3, ab.scores.shape, generator=g).float()  # {-2,-1,0}: many exact ties
    p, sc = ab.to(DEV).pack()
    assert p.has_sell
    viterbi_matches(ab, p, sc)


def test_sell_agrees_with_csr_kernels(monkeypatch):
    ab = synth.random_dag_batch(4, 50_000, levels=32, seed=13).to(DEV)
    ps, scs = ab.pack()
    monkeypatch.setattr(nb.pack, "SELL", 0)
    pc, scc = ab.pack()
    assert ps.has_sell and not pc.has_sell
    ls, as_, bs, posts = nb.lattice_forward_backward(ps, arc_scores=scs)
    lc, ac, bc, postc = nb.lattice_forward_backward(pc, arc_scores=scc)
    assert torch.allclose(ls, lc, rtol=1e-6, atol=1e-4)
    # same arcs, different canonical order: compare through the caller's arc ids
    a = torch.empty_like(posts); a[ps.arc_origin] = posts
    c = torch.empty_like(postc); c[pc.arc_origin] = postc
    assert torch.allclose(a, c, rtol=3e-5, atol=1e-7)
    vs = nb.lattice_viterbi(ps, arc_scores=scs)
    vc = nb.lattice_viterbi(pc, arc_scores=scc)
    assert torch.equal(vs[0], vc[0]) and torch.equal(vs[1], vc[1])
    assert torch.equal(ps.arc_origin[vs[2].long()], pc.arc_origin[vc[2].long()]) and torch.equal(vs[3], vc[3])


def test_sell_autograd_arc_scores_and_theta():
    ab = synth.random_dag_batch(5, 8_000, levels=16, seed=17)
    p, sc = ab.to(DEV).pack()
    assert p.has_sell
    o_logz, _, _, o_post = c_oracle.forward_backward(oracle_batch(ab))
    origin = p.arc_origin.cpu().numpy()
    # per-arc scores, non-trivial upstream gradient
    w = sc.clone().requires_grad_(True)
    coef = torch.linspace(0.5, 2.0, p.n_lattices, device=DEV)
    logz = nb.lattice_log_partition(p, arc_scores=w)
    (logz * coef).sum().backward()
    np.testing.assert_allclose(logz.detach().cpu().numpy(), o_logz, rtol=1e-6)
    arc_lat = np.repeat(np.arange(p.n_lattices), np.diff(p.arc_off.cpu().numpy()))
    ref = o_post[origin] * coef.cpu().numpy()[arc_lat]
    got = w.grad.cpu().numpy()
    assert np.all(np.abs(got - ref) <= 1e-5 * ref + 1e-7)
    # theta mode: d logZ / d theta[l] = sum of the posteriors of the arcs labelled l
    theta = torch.randn(p.vocab, device=DEV, requires_grad=True)
    nb.lattice_log_partition(p, theta=theta).sum().backward()
    wt = theta.detach().cpu().numpy()[ab.label.numpy()]
    ab2 = synth.ArcBatch(ab.arc_lattice, ab.src, ab.dst, ab.label, torch.from_numpy(wt), ab.n_states, ab.vocab)
    o2_logz, _, _, po = c_oracle.forward_backward(oracle_batch(ab2))
    dth = np.zeros(p.vocab)
    np.add.at(dth, ab.label.numpy(), po)
    np.testing.assert_allclose(theta.grad.cpu().numpy(), dth, rtol=1e-4, atol=1e-5)
    # theta + per-arc scores together, fused call with the histogram
    logz2, alpha, beta, post, dtheta = nb.lattice_forward_backward(p, arc_scores=sc, theta=theta.detach(), want_dtheta=True)
    wb = (ab.scores.numpy() + wt).astype(np.float32)
    ab3 = synth.ArcBatch(ab.arc_lattice, ab.src, ab.dst, ab.label, torch.from_numpy(wb), ab.n_states, ab.vocab)
    o3_logz, _, _, po3 = c_oracle.forward_backward(oracle_batch(ab3))
    np.testing.assert_allclose(logz2.cpu().numpy(), o3_logz, rtol=1e-5)
    d3 = np.zeros(p.vocab)
    np.add.at(d3, ab.label.numpy(), po3)
    np.testing.assert_allclose(dtheta.cpu().numpy(), d3, rtol=1e-4, atol=1e-5)


def test_sell_mixed_batch_and_lattice_backward_outputs():
    parts = [synth.transliteration_batch(5, seed=2), synth.random_dag_batch(3, 30_000, levels=32, seed=4),
             synth.snips_batch(4, seed=1)]
    packs, scores = zip(*[ab.to(DEV).pack() for ab in parts])
    p = concat_packed(list(packs))
    sc = torch.cat(scores)
    kinds = {g.sell for g in p.groups}
    assert kinds == {True, False}
    logz, alpha, beta, post = nb.lattice_forward_backward(p, arc_scores=sc)
    off_a = 0
    off_b = 0
    for ab, pk in zip(parts, packs):
        o_logz, o_alpha, o_beta, o_post = c_oracle.forward_backward(oracle_batch(ab))
        np.testing.assert_allclose(logz[off_b:off_b + pk.n_lattices].cpu().numpy(), o_logz, rtol=1e-5)
        ref = o_post[pk.arc_origin.cpu().numpy()]
        got = post[off_a:off_a + pk.n_arcs].cpu().numpy().astype(np.float64)
        # sliced-column part: 1e-5 flat; the fp32 CSR kernels: depth-scaled (tests/test_gpu_parity.py)
        rt = 1e-5 if pk.has_sell else post_rtol(float(np.abs(o_alpha).max() + np.abs(o_beta).max()))
        assert np.all(np.abs(got - ref) <= rt * ref + 1e-7), (rt, float(np.max(np.abs(got - ref) / (ref + 1e-7))), pk.n_arcs)
        off_a += pk.n_arcs
        off_b += pk.n_lattices
    # separate calls: forward (alpha, logZ) and the fused backward with every output
    alpha2, logz2 = nb.lattice_forward(p, arc_scores=sc)
    assert torch.allclose(alpha2, alpha, rtol=1e-5, atol=1e-4) and torch.allclose(logz2, logz, rtol=1e-6, atol=1e-5)
    r = nb.ops.lattice_backward(p, sc, alpha=alpha2, logz=logz2, want_beta=True, want_post=True, want_viterbi=True)
    assert torch.allclose(r["beta"], beta, rtol=1e-6, atol=1e-5)
    assert torch.allclose(r["post"], post, rtol=1e-5, atol=1e-7)
    vs, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=sc)
    assert torch.equal(r["vit_score"], vs)


def test_sell_rejects_misuse():
    ab = synth.random_dag_batch(2, 5_000, levels=8, seed=3).to(DEV)
    p, sc = ab.pack()
    assert p.has_sell
    lib = nb._lib.load()
    lc = nb.ops._launch(p.groups[0], torch.float32)
    scs, keep = nb.ops._scores(p, sc, None)
    a = torch.empty(p.n_states, device=DEV)
    z = torch.empty(p.n_lattices, device=DEV)
    # the CSR entry points refuse a sliced-column group
    rc = lib.nfst_fwd_f32(p.c_struct(), lc, scs, a.data_ptr(), z.data_ptr(), None)
    assert rc < 0 and b"sliced-column" in lib.nfst_last_error_string()
    # the flow pass needs cond and post
    rc = lib.nfst_sell_flow_f32(p.c_struct(), lc, None, None, None, None, None, None, None, None, None)
    assert rc < 0


def test_sell_staging_edges():
    """The 16-byte staging copies: an arc count that is not a multiple of 4 (the last slice of the batch is copied
    by element), a misaligned view of the scores (the operator copies it once), and the C ABI's refusal of a
    misaligned conditional buffer."""
    for seed in range(20, 40):
        ab = synth.random_dag_batch(3, 9_000, levels=12, seed=seed)
        p, sc = ab.to(DEV).pack()
        if p.n_arcs % 4:
            break
    assert p.has_sell and p.n_arcs % 4
    o_logz, _, _, o_post = c_oracle.forward_backward(oracle_batch(ab))
    ref = o_post[p.arc_origin.cpu().numpy()]
    buf = torch.empty(p.n_arcs + 1, device=DEV)
    view = buf[1:]  # 4 bytes off a 16-byte boundary
    view.copy_(sc)
    assert view.data_ptr() % 16
    for w in (sc, view):
        logz, _, _, post = nb.lattice_forward_backward(p, arc_scores=w)
        np.testing.assert_allclose(logz.cpu().numpy(), o_logz, rtol=1e-6)
        assert np.all(np.abs(post.cpu().numpy() - ref) <= 1e-5 * ref + 1e-7)
    logz, _, cond = nb.ops.lattice_pull(p, arc_scores=sc)
    bad = torch.empty(p.n_arcs + 1, device=DEV)[1:]
    bad.copy_(cond)
    with pytest.raises(RuntimeError, match="16-byte aligned"):
        nb.lattice_backward(p, arc_scores=sc, logz=logz, cond=bad, want_beta=False, want_post=True)
