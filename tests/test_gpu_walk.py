"""GPU parity of the fused sampling-loop step (nfst_walk_step_f32) against the reference's own loop (golden
vectors) and against the numpy restatement on larger random cases, in both look-ahead modes, for boolean and
weighted tables, scoring and sampling."""
import os

import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200.sampler import LatticeWalker, walk_step
from oracle import lattice_oracle as lo
from oracle import walk_oracle as wo
from tests.lattice_gen import PAD, random_mark_lattice

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = os.path.join(os.path.dirname(__file__), "golden", "walk_step.npz")


def to_packed_states(p, k, dense_state):
    """dense (reference) state ids of the rows -> packed ids"""
    B, S = p.dense_shape[0], p.dense_shape[1]
    inv = torch.full((B, S), -1, dtype=torch.int64)
    so = p.state_off.cpu().long()
    orig = p.orig_state.cpu().long()
    for b in range(B):
        inv[b, orig[so[b]:so[b + 1]]] = torch.arange(so[b], so[b + 1])
    rows = torch.arange(B * k) // k
    out = inv[rows, torch.as_tensor(dense_state).long()]
    assert (out >= 0).all()
    return out.to(torch.int32).to(DEV)


def packed_beta(p, k, beta_dense):
    """beta[B*k, S] (reference layout, real space) -> one value per packed state"""
    so = p.state_off.cpu().numpy()
    lat = np.repeat(np.arange(p.n_lattices), np.diff(so))
    return torch.from_numpy(beta_dense[lat * k, p.orig_state.cpu().numpy()].astype(np.float32)).to(DEV)


def test_walk_step_matches_the_reference_loop():
    g = np.load(G)
    k, T, pad = int(g["k"]), float(g["temperature"]), int(g["pad"])
    tr = torch.from_numpy(g["tr"]).to(DEV)
    p = nb.pack_dense(tr != 0, tr, weighted=False)
    beta = packed_beta(p, k, g["beta"])
    # (1) every recorded step on its own
    for t in range(int(g["steps"])):
        sym, logp, nxt, logz = walk_step(
            p, k, to_packed_states(p, k, g[f"state_new_{t}"]), torch.from_numpy(g[f"prefix_{t}"]).to(DEV), beta, pad,
            base_mask=torch.from_numpy(g[f"base_{t}"]).to(DEV), symbols=torch.from_numpy(g[f"sym_{t}"]).to(DEV),
            temperature=T, look_state=to_packed_states(p, k, g[f"state_old_{t}"]))
        np.testing.assert_array_equal(sym.cpu().numpy(), g[f"sym_{t}"])
        np.testing.assert_allclose(logp.cpu().numpy(), g[f"logp_{t}"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(logz.cpu().numpy(), g[f"zs_{t}"], rtol=1e-5, atol=2e-6)
        np.testing.assert_array_equal(p.orig_state[nxt.long()].cpu().numpy(), g[f"next_{t}"])
    # (2) the walker object replays the whole loop: bos is consumed first (scorers.py:230-231), then one
    # scored symbol per step; its states are the reference's metadata["state"]
    w = LatticeWalker(p, k, beta, pad, temperature=T, faithful=True)
    w.consume(torch.full((w.n_rows,), int(g["bos"]), dtype=torch.int32, device=DEV))
    total = torch.zeros(w.n_rows, device=DEV)
    for t in range(int(g["steps"])):
        np.testing.assert_array_equal(w.dense_state().cpu().numpy(), g[f"state_new_{t}"])
        sym, logp, logz = w.step(torch.from_numpy(g[f"prefix_{t}"]).to(DEV), base_mask=torch.from_numpy(g[f"base_{t}"]).to(DEV),
                                 symbols=torch.from_numpy(g[f"sym_{t}"]).to(DEV))
        np.testing.assert_allclose(logp.cpu().numpy(), g[f"logp_{t}"], rtol=1e-5, atol=2e-6)
        total += logp
    ref_total = sum(g[f"logp_{t}"] for t in range(int(g["steps"])))
    np.testing.assert_allclose(total.cpu().numpy(), ref_total, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("faithful", [True, False])
def test_walk_step_matches_oracle_on_random_walks(weighted, faithful):
    rng = np.random.default_rng(7 + weighted)
    V, k, T = 40, 4, 1.3
    tabs = [random_mark_lattice(rng, n, V, parallel_arcs=False) for n in (12, 30, 21, 5)]
    tr = lo.collate_pad([t[1] for t in tabs], PAD)
    em_bool = lo.collate_pad([t[1] != 0 for t in tabs], PAD)
    if weighted:  # float log-weights, -inf = no arc (scorers.py:1011-1013,1026-1027)
        em = np.where(tr != 0, rng.normal(size=tr.shape), -np.inf)
        for b, t in enumerate(tabs):  # the sink's pad loop carries weight 0 (scorers.py:1013-1016)
            em[b, t[1].shape[0] - 1, PAD] = 0.0
    else:
        em = em_bool
    B, S, _ = tr.shape
    N = B * k
    trt = torch.from_numpy(tr).to(DEV)
    p = nb.pack_dense(torch.from_numpy(em).to(DEV), trt, weighted=weighted)
    beta_dense = np.abs(rng.normal(size=(N, S))).astype(np.float32) + 0.1  # any positive numbers will do
    beta_dense = np.repeat(beta_dense[::k], k, axis=0)  # rows of one lattice share beta
    beta = packed_beta(p, k, beta_dense)
    state_old = np.zeros(N, dtype=np.int64)
    state_new = tr[np.arange(N) // k, 0, 1].astype(np.int64)  # after bos
    for step in range(10):
        prefix = rng.normal(size=(N, V)).astype(np.float32)
        base = np.where(rng.random((N, V)) < 0.1, -1.5, 0.0).astype(np.float32)
        # pick a valid symbol per row (pad at the sink, whose pad loop the edge rule drops)
        syms = np.empty(N, dtype=np.int64)
        for n in range(N):
            valid = np.nonzero(tr[n // k, state_new[n]] != 0)[0]
            syms[n] = rng.choice(valid)
        look_old = state_old if faithful else None
        st_old_for_oracle = state_old if faithful else state_new  # aligned: the row of the state itself
        masked, o_logp, o_logz, o_next = wo.walk_step_dense(em, tr, k, beta_dense, st_old_for_oracle, state_new, prefix, base,
                                                            PAD, T, syms)
        kw = dict(base_mask=torch.from_numpy(base).to(DEV), temperature=T,
                  look_state=None if look_old is None else to_packed_states(p, k, look_old))
        sym, logp, nxt, logz = walk_step(p, k, to_packed_states(p, k, state_new), torch.from_numpy(prefix).to(DEV), beta, PAD,
                                         symbols=torch.from_numpy(syms).to(DEV), **kw)
        np.testing.assert_allclose(logp.cpu().numpy(), o_logp, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(logz.cpu().numpy(), o_logz, rtol=1e-5, atol=1e-5)
        np.testing.assert_array_equal(p.orig_state[nxt.long()].cpu().numpy(), o_next)
        # sampling: inverse CDF over the arcs in label order == over the dense row in label order
        u = rng.random(N).astype(np.float32)
        s_sym, s_logp, s_nxt, _ = walk_step(p, k, to_packed_states(p, k, state_new), torch.from_numpy(prefix).to(DEV), beta, PAD,
                                            uniform=torch.from_numpy(u).to(DEV), **kw)
        probs = np.exp(masked - o_logz[:, None])
        cdf = np.cumsum(probs, axis=1)
        s_sym = s_sym.cpu().numpy()
        for n in range(N):
            want = int(np.searchsorted(cdf[n], u[n], side="right"))
            if want != s_sym[n]:  # only at a boundary of the CDF (float32 vs float64 round-off)
                assert abs(cdf[n, min(want, V - 1)] - u[n]) < 1e-5 or abs(cdf[n, s_sym[n]] - u[n]) < 1e-5
            assert probs[n, s_sym[n]] > 0
            assert abs(s_logp[n].item() - np.log(probs[n, s_sym[n]])) < 1e-4
        state_old, state_new = state_new, o_next


def test_walk_sampling_frequencies_and_misuse():
    rng = np.random.default_rng(3)
    V, k = 24, 4096
    _, tr = random_mark_lattice(rng, 6, V, parallel_arcs=False)
    trt = torch.from_numpy(tr)[None].to(DEV)
    p = nb.pack_dense(trt != 0, trt, weighted=False)
    beta = torch.ones(p.n_states, device=DEV)
    w = LatticeWalker(p, k, beta, PAD, faithful=False)
    w.consume(torch.full((k,), 1, dtype=torch.int32, device=DEV))
    prefix = torch.randn(V, device=DEV).expand(k, V).contiguous()
    s0 = int(w.dense_state()[0])
    sym, logp, logz = w.step(prefix)
    valid = np.nonzero(tr[s0] != 0)[0]
    pr = torch.softmax(prefix[0, valid] + 1.0, 0).cpu().numpy()
    freq = np.array([(sym == int(l)).float().mean().item() for l in valid])
    assert np.all(np.abs(freq - pr) < 4 * np.sqrt(pr * (1 - pr) / k) + 1e-3)
    with pytest.raises(ValueError):
        walk_step(p, k, w.state, prefix, beta, PAD)  # neither symbols nor uniform
    lib = nb._lib.load()
    rc = lib.nfst_walk_step_f32(p.c_struct(), k, k, w.state.data_ptr(), None, prefix.data_ptr(), None, beta.data_ptr(), None,
                                0.0, PAD, None, prefix.data_ptr(), sym.data_ptr(), logp.data_ptr(), sym.data_ptr(), None, None)
    assert rc < 0 and b"temperature" in lib.nfst_last_error_string()


def test_sample_paths_are_exact_posterior_samples():
    """The whole loop in one launch for an arc-factored model: every sampled path is a start->sink path, its
    log q is score(path) - logZ (so every IWAE weight, estimatros.py:10-44, equals logZ), arc frequencies match
    the posteriors, and the walk is the inverse-CDF walk of the float64 restatement on the same uniforms."""
    from nfst_b200 import synth
    from nfst_b200.sampler import sample_paths
    from oracle import c_oracle

    ab = synth.transliteration_batch(6, seed=9)
    p, sc = ab.to(DEV).pack()
    k = 2048
    T = p.max_levels - 1
    g = torch.Generator(device=DEV).manual_seed(5)
    u = torch.rand(p.n_lattices * k, T, device=DEV, generator=g)
    labels, length, log_q, arcs, logz = sample_paths(p, k, arc_scores=sc, uniform=u, pad_id=PAD)
    ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(),
                        ab.n_states.numpy())
    o_logz, _, o_beta, o_post = c_oracle.forward_backward(ob)
    np.testing.assert_allclose(logz.cpu().numpy(), o_logz, rtol=1e-5)
    arcs_c, len_c, lq = arcs.cpu().numpy(), length.cpu().numpy(), log_q.cpu().numpy()
    w, dst, src = sc.cpu().numpy().astype(np.float64), p.dst_out.cpu().numpy(), p.src_out.cpu().numpy()
    out_ptr, start = p.out_ptr.cpu().numpy(), p.start_state.cpu().numpy()
    lab = labels.cpu().numpy()
    for r in range(0, p.n_lattices * k, 97):  # a sample of the rows, path by path
        b, path = r // k, arcs_c[r, :len_c[r]]
        assert src[path[0]] == start[b] and np.array_equal(dst[path[:-1]], src[path[1:]])
        assert out_ptr[dst[path[-1]] + 1] == out_ptr[dst[path[-1]]]  # ends at a state without arcs
        assert np.all(arcs_c[r, len_c[r]:] == -1) and np.all(lab[r, len_c[r]:] == PAD)
        np.testing.assert_array_equal(lab[r, :len_c[r]], p.label_out.cpu().numpy()[path])
        assert abs(lq[r] - (w[path].sum() - o_logz[b])) < 2e-4  # log q = score - logZ: the IWAE weight is logZ
    # zero-variance estimate: logsumexp_k(score - log q) - log k == logZ for every lattice
    score = torch.zeros_like(log_q)
    valid = arcs >= 0
    score = (sc[arcs.clamp(min=0).long()] * valid).sum(1)
    est = torch.logsumexp((score - log_q).view(p.n_lattices, k), 1) - np.log(k)
    np.testing.assert_allclose(est.cpu().numpy(), o_logz, rtol=1e-5, atol=1e-4)
    # arc frequencies ~ posteriors
    counts = torch.bincount(arcs[valid].long(), minlength=p.n_arcs).cpu().numpy() / k
    ref = o_post[p.arc_origin.cpu().numpy()]
    assert np.all(np.abs(counts - ref) < 5 * np.sqrt(np.maximum(ref * (1 - ref), 1e-9) / k) + 2e-3)
    # inverse-CDF walk of the float64 restatement on the same uniforms (first rows of lattice 0)
    so = np.concatenate([[0], np.cumsum(ab.n_states.numpy())])
    state_off = p.state_off.cpu().numpy()
    lat = np.repeat(np.arange(p.n_lattices), np.diff(state_off))
    beta = o_beta[so[lat] + p.orig_state.cpu().numpy()]
    uu = u.cpu().numpy()
    mism = 0
    for r in range(64):
        s, t = start[0], 0
        while out_ptr[s + 1] > out_ptr[s]:
            a = np.arange(out_ptr[s], out_ptr[s + 1])
            c = np.exp(w[a] + beta[dst[a]] - beta[s])
            j = min(int(np.searchsorted(np.cumsum(c), uu[r, t], side="right")), len(a) - 1)
            if a[j] != arcs_c[r, t]:
                mism += 1
                break
            s, t = dst[a[j]], t + 1
    assert mism <= 1  # a uniform number can fall on a CDF boundary (fp32 beta on the GPU side)


def test_full_stateful_sample_call_replays_through_the_patched_module_and_the_walker():
    """tests/golden/stateful_sample.npz is ONE FULL CALL of the reference's Sampler.stateful_sample on a real
    FSAGRUScorer(use_beta=True) (samplers.py:182-335).  Here: an object with that module's attribute names and its
    recorded parameters gets patch_compute_beta -> its compute_beta() equals the reference's beta; LatticeWalker then
    replays the whole loop on the recorded per-step network outputs and vocabulary masks, scoring the recorded samples:
    the summed log-probabilities equal what the reference's call returned."""
    from types import SimpleNamespace

    from nfst_b200.scorer import patch_compute_beta

    g = np.load(os.path.join(os.path.dirname(G), "stateful_sample.npz"))
    k, T, pad, bos, steps = int(g["k"]), float(g["temperature"]), int(g["pad"]), int(g["bos"]), int(g["steps"])
    tr, em = torch.from_numpy(g["tr"]).to(DEV), torch.from_numpy(g["em"]).to(DEV)
    dev = lambda name: torch.from_numpy(g["p_" + name]).to(DEV)  # noqa: E731
    mod = SimpleNamespace(embeddings=SimpleNamespace(weight=dev("emb")), Wx=dev("Wx"), Wh=dev("Wh"), W=dev("W"),
                          beta_bias=dev("bias"), emission=em, transition=tr, k=k)
    patch_compute_beta(mod)
    beta_dense = mod.compute_beta()
    assert beta_dense.shape == g["beta"].shape and beta_dense.dtype == torch.float32
    B, S = tr.shape[0], tr.shape[1]
    p = nb.pack_dense(em, tr, weighted=False)
    # compare where the reference's value is meaningful: the states the start reaches (the rest are collate padding)
    so = p.state_off.cpu().numpy()
    lat = np.repeat(np.arange(B), np.diff(so))
    orig = p.orig_state.cpu().numpy()
    np.testing.assert_allclose(beta_dense.cpu().numpy()[lat * k, orig], g["beta"][lat * k, orig], rtol=2e-5)
    w = LatticeWalker(p, k, packed_beta(p, k, beta_dense.cpu().numpy()), pad, temperature=T, faithful=True)
    w.consume(torch.full((w.n_rows,), bos, dtype=torch.int32, device=DEV))
    seqs = g["sequences"]  # [N, steps - 1]; the call pops the final (all-pad) step
    total = torch.zeros(w.n_rows, device=DEV)
    for t in range(steps):
        sym = torch.from_numpy(seqs[:, t] if t < seqs.shape[1] else np.full(w.n_rows, pad)).to(DEV)
        _, logp, _ = w.step(torch.from_numpy(g[f"prefix_{t}"]).to(DEV), base_mask=torch.from_numpy(g[f"base_{t}"]).to(DEV), symbols=sym)
        total += logp
    np.testing.assert_allclose(total.cpu().numpy(), g["summed_log_probs"], rtol=1e-5, atol=1e-5)
