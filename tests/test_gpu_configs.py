"""GPU parity at the FULL batch sizes BASELINE.json states for its five configs.

Every test generates the config's synthetic batch at its stated size (in chunks, on the device), packs it,
runs log-semiring forward-backward and Viterbi + backtrace over the WHOLE batch through the C ABI, and compares
against the C oracle (oracle/lattice_oracle.c, float64):

  * logZ of EVERY lattice against a float64 restatement of the recurrence in plain torch ops on the ORIGINAL arc
    list (Bellman sweeps until nothing changes; independent of the packer and of the kernels), which is itself
    held to the C oracle on the sample;
  * arc posteriors (|p - ref| <= 1e-5 ref + 2e-9), Viterbi scores (bit-exact) and best-path label sequences
    (equal) on a seeded random sample of 16 lattices taken out of the full-batch results;
  * size-independent identities on all lattices: the posterior mass leaving the start state is 1, logZ is finite.

Tolerances are BASELINE.json north_star's: 1e-5 relative in fp32 (for logZ, a log value that may be near zero:
1e-5 relative on Z, i.e. rtol = atol = 1e-5 on logZ), Viterbi bit-exact.
"""
import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import synth
from nfst_b200.pack import concat_packed
from oracle import c_oracle

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = torch.device("cuda", 0)
N_SAMPLE = 16
# posteriors: 1e-5 relative plus an ABSOLUTE 2e-9 -- the tile-stream flow pass accumulates state posteriors in 2^-31
# fixed point (native integer shared-memory atomics, bit-reproducible), so a posterior carries an absolute error of a
# few 5e-10 units: 1e-5 relative down to posteriors of ~1e-4, absolute below (NFST_FLOW_BITS=64 is relative at any size)
ATOL_POST = 2e-9


def logz_by_sweeps(ab: synth.ArcBatch) -> torch.Tensor:
    """float64 logZ[B] of a batch from its original arc list: beta = 0 at arc-less states (scorers.py:720),
    beta[s] = logsumexp over the arcs out of s of (w + beta[dst]) (scorers.py:741-749 in log space), swept until nothing
    changes -- on an acyclic lattice that happens after `depth` sweeps and is the exact recurrence."""
    n = int(ab.n_states.numel())
    off = torch.zeros(n + 1, dtype=torch.int64, device=ab.src.device)
    off[1:] = torch.cumsum(ab.n_states, 0)
    gs, gd = off[ab.arc_lattice] + ab.src, off[ab.arc_lattice] + ab.dst
    w = ab.scores.double()
    S = int(off[-1])
    has_out = torch.bincount(gs, minlength=S) > 0
    beta = torch.zeros(S, dtype=torch.float64, device=gs.device)
    ninf = torch.full((S,), float("-inf"), dtype=torch.float64, device=gs.device)
    # a state at distance k from the arc-less states is final after k sweeps; the float64 atomics of index_add_ sum
    # in no fixed order, so "final" means equal to rounding, not bit for bit
    for it in range(int(ab.n_states.max()) + 8):
        t = w + beta[gd]
        m = ninf.scatter_reduce(0, gs, t, reduce="amax")
        acc = torch.zeros_like(beta).index_add_(0, gs, torch.exp(t - m[gs]))
        new = torch.where(has_out, m + torch.log(acc), torch.zeros_like(beta))
        if it % 16 == 15 and float((new - beta).abs().max()) <= 1e-13 * max(1.0, float(new.abs().max())):
            beta = new
            break
        beta = new
    else:
        raise AssertionError("the sweeps did not settle: cyclic lattice?")
    return beta[off[:-1]]  # the start state is state 0 of its lattice (scorers.py:1005)


def build_with_sample(gen, B, per_chunk, seed):
    """(packed batch, scores, sample, reference logZ of every lattice): the batch generated and packed chunk by
    chunk; for N_SAMPLE seeded lattices the ORIGINAL arcs (host copies) and, per packed arc of the lattice, its
    position in them."""
    rng = np.random.default_rng(seed)
    ids = np.sort(rng.choice(B, size=min(N_SAMPLE, B), replace=False))
    parts, scores, sample, ref_logz = [], [], [], []
    done, arc_base = 0, 0
    while done < B:
        n = min(per_chunk, B - done)
        ab = gen(n, done)
        ab = ab if ab.src.device == DEV else ab.to(DEV)
        p, sc = ab.pack()
        ref_logz.append(logz_by_sweeps(ab))
        arc_off = p.arc_off.cpu().numpy()
        for b in ids[(ids >= done) & (ids < done + n)] - done:
            idx = torch.nonzero(ab.arc_lattice == int(b)).squeeze(1)  # ascending positions in the chunk's arc list
            a0, a1 = int(arc_off[b]), int(arc_off[b + 1])
            pos = torch.searchsorted(idx, p.arc_origin[a0:a1])
            assert bool((idx[pos] == p.arc_origin[a0:a1]).all()), "arc_origin points into the lattice's own arcs"
            sample.append(dict(lattice=done + int(b), a0=arc_base + a0, a1=arc_base + a1, pos=pos.cpu().numpy(),
                               src=ab.src[idx].cpu().numpy(), dst=ab.dst[idx].cpu().numpy(), label=ab.label[idx].cpu().numpy(),
                               scores=ab.scores[idx].cpu().numpy(), n_states=int(ab.n_states[b])))
        p.arc_origin = torch.empty(0, dtype=torch.int64, device=DEV)  # per-chunk; not needed any more
        parts.append(p)
        scores.append(sc)
        arc_base += p.n_arcs
        done += n
        del ab
    packed = concat_packed(parts) if len(parts) > 1 else parts[0]
    return packed, torch.cat(scores), sample, torch.cat(ref_logz).cpu().numpy()


def oracle_on(sample):
    lat = np.concatenate([np.full(len(s["src"]), j, dtype=np.int64) for j, s in enumerate(sample)])
    cat = lambda k: np.concatenate([s[k] for s in sample])  # noqa: E731
    ob = c_oracle.Batch(lat, cat("src"), cat("dst"), cat("label"), cat("scores"), [s["n_states"] for s in sample])
    logz, _, _, post = c_oracle.forward_backward(ob)
    vscore, vpaths, vlabels = c_oracle.viterbi(ob)
    off = np.concatenate([[0], np.cumsum([len(s["src"]) for s in sample])])
    return logz, post, vscore, vlabels, off


def check_config(gen, B, per_chunk, *, seed=0, expect=None):
    packed, sc, sample, ref_logz = build_with_sample(gen, B, per_chunk, seed)
    assert packed.n_lattices == B
    if expect is not None:
        assert expect(packed), [(g.tiles, g.sell, g.small, g.n) for g in packed.groups]
    logz, _, _, post = nb.lattice_forward_backward(packed, arc_scores=sc)
    vscore, voff, varcs, vlabels = nb.lattice_viterbi(packed, arc_scores=sc)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(logz).all())
    # identity on every lattice: the posterior mass that leaves the start state is 1
    src_out = packed.src_out.long()
    start = packed.start_state.long()
    lat_of_arc = torch.repeat_interleave(torch.arange(B, device=DEV), (packed.arc_off[1:] - packed.arc_off[:-1]).long())
    from_start = src_out == start[lat_of_arc]
    mass = torch.zeros(B, dtype=torch.float64, device=DEV).index_add_(0, lat_of_arc[from_start], post[from_start].double())
    np.testing.assert_allclose(mass.cpu().numpy(), 1.0, rtol=2e-5)
    # the sample against the oracle
    o_logz, o_post, o_vs, o_vlab, off = oracle_on(sample)
    logz_c, vs_c, voff_c, vlab_c = logz.cpu().numpy().astype(np.float64), vscore.cpu().numpy(), voff.cpu().numpy(), vlabels.cpu().numpy()
    worst = 0.0
    for j, s in enumerate(sample):
        b = s["lattice"]
        np.testing.assert_allclose(logz_c[b], o_logz[j], rtol=1e-5, atol=1e-5)
        got = post[s["a0"]:s["a1"]].cpu().numpy().astype(np.float64)
        ref = o_post[off[j] + s["pos"]]
        err = np.abs(got - ref) - (1e-5 * ref + ATOL_POST)
        worst = max(worst, float(np.max(np.abs(got - ref) / (ref + 1e-7))))
        assert np.all(err <= 0), (b, float(np.max(np.abs(got - ref) / (ref + 1e-7))))
        assert vs_c[b:b + 1].view(np.uint32)[0] == o_vs[j:j + 1].view(np.uint32)[0], "Viterbi score bit-exact"
        assert list(vlab_c[voff_c[b]:voff_c[b + 1]]) == list(o_vlab[j]), "Viterbi path"
    # every lattice's logZ against the float64 sweeps, and the sweeps against the oracle on the sample
    np.testing.assert_allclose(logz_c, ref_logz, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ref_logz[[s["lattice"] for s in sample]], o_logz, rtol=1e-12)
    return packed, worst


def test_config1_transliteration_b32():
    check_config(lambda n, o: synth.transliteration_batch(n, seed=o), 32, 32)


def test_config2_snips_b256():
    check_config(lambda n, o: synth.snips_batch(n, seed=1 + o), 256, 256)


@pytest.mark.parametrize("bigram", [False, True])
def test_config3_cipher_t1000_b64(bigram):
    # depth 1000: the state vectors are float64 ("auto"), which is what holds 1e-5 at this depth
    p, _ = check_config(lambda n, o: synth.cipher_batch(n, T=1000, bigram=bigram, seed=2 + o, device=DEV), 64, 16 if bigram else 64)
    assert nb.ops.resolve_state_dtype(p) == torch.float64


@pytest.mark.parametrize("arcs", [10_000, 100_000, 300_000, 1_000_000])
def test_config4_random_dag_b1024(arcs):
    check_config(lambda n, o: synth.random_dag_batch(n, arcs, levels=64, seed=3 + o, device=DEV), 1024, max(1, 60_000_000 // arcs),
                 expect=lambda p: all(g.tiles for g in p.groups))


@pytest.mark.parametrize("integer_scores", [False, True])
def test_config5_viterbi_b4096(integer_scores):
    check_config(lambda n, o: synth.transliteration_batch(n, seed=4 + o, integer_scores=integer_scores), 4096, 4096)
