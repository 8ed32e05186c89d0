"""Throughput + parity over the five BASELINE.json configs on one GPU (not the bench line).

    python tests/report/config_report.py [--quick] > profiles/rN_configs.txt

For every config: pack time, forward+fused-backward (beta + posteriors) and Viterbi+backtrace
time (CUDA events, 3 warm-ups, L2 flushed between iterations when the inputs fit in L2),
arcs/s, algorithmic GB/s, the C oracle (CPU port, all host threads) on a bounded sample, and
parity of the GPU results against that oracle on the same sample.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import synth  # noqa: E402
from nfst_b200.pack import concat_packed  # noqa: E402
from oracle import c_oracle  # noqa: E402

DEV = torch.device("cuda", 0)
ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-dag", action="store_true", help="skip the config-4 sweep")
args = ap.parse_args()

flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)  # > 126 MB L2


def timed(fn, steps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(steps):
        if flush:
            flush_buf.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def build(gen, B, per_chunk):
    parts, scores, sample = [], [], None
    done = 0
    t0 = time.perf_counter()
    while done < B:
        n = min(per_chunk, B - done)
        ab = gen(n, done)
        if sample is None:
            sample = ab
        p, sc = ab.to(DEV).pack()
        parts.append(p)
        scores.append(sc)
        done += n
    packed = concat_packed(parts) if len(parts) > 1 else parts[0]
    torch.cuda.synchronize()
    return packed, torch.cat(scores), sample, time.perf_counter() - t0


def parity(sample: synth.ArcBatch, n_max=16):
    """GPU vs C oracle on (a slice of) the first generated chunk."""
    B = min(int(sample.n_states.numel()), n_max)
    sel = (sample.arc_lattice < B).cpu()
    ab = synth.ArcBatch(sample.arc_lattice.cpu()[sel], sample.src.cpu()[sel], sample.dst.cpu()[sel], sample.label.cpu()[sel],
                        sample.scores.cpu()[sel], sample.n_states.cpu()[:B], sample.vocab)
    p, sc = ab.to(DEV).pack()
    logz, alpha, beta, post = nb.lattice_forward_backward(p, arc_scores=sc)
    vs, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=sc)
    ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(),
                        ab.n_states.numpy())
    t0 = time.perf_counter()
    o_logz, _, _, o_post = c_oracle.forward_backward(ob)
    t_cpu = time.perf_counter() - t0
    o_vs, o_paths, _ = c_oracle.viterbi(ob)
    ref = o_post[p.arc_origin.cpu().numpy()]
    got = post.cpu().numpy().astype(np.float64)
    big = ref >= 1e-4  # relative error where 1e-5 relative is above the fixed-point unit; absolute error below
    perr = float(np.max(np.abs(got[big] - ref[big]) / ref[big])) if big.any() else 0.0
    aerr = float(np.max(np.abs(got[~big] - ref[~big]))) if (~big).any() else 0.0
    zerr = float(np.max(np.abs(logz.cpu().numpy() - o_logz) / np.maximum(1.0, np.abs(o_logz))))
    vexact = int(np.sum(vs.cpu().numpy().view(np.uint32) == o_vs.view(np.uint32)))
    origin, offc, arcs_c = p.arc_origin.cpu().numpy(), off.cpu().numpy(), arcs.cpu().numpy()
    pexact = sum(int(np.array_equal(origin[arcs_c[offc[b]:offc[b + 1]]], o_paths[b])) for b in range(B))
    return dict(B=B, state=str(alpha.dtype).replace("torch.", ""), logz_rel=zerr, post_rel=perr, post_abs=aerr, vit_scores_exact=f"{vexact}/{B}",
                vit_paths_exact=f"{pexact}/{B}", cpu_arcs_per_s=ob.n_arcs / t_cpu, cpu_threads=c_oracle.max_threads())


def recurrent_beta_report(B=32, H=256):
    """compute_beta with Wh != 0 (the beta-hat recurrence) on config-1 lattices,
    GPU (level-stepped kernel) next to the float64 numpy port of the reference recurrence."""
    from oracle import lattice_oracle as lo

    ab = synth.transliteration_batch(B, seed=0)
    p, _ = ab.to(DEV).pack()
    gen = torch.Generator().manual_seed(0)
    V = ab.vocab
    emb = torch.randn(V, H, generator=gen, dtype=torch.float64)
    Wx, Wh = (torch.randn(H, H, generator=gen, dtype=torch.float64) / H ** 0.5 for _ in range(2))
    W = torch.randn(1, H, generator=gen, dtype=torch.float64) / H ** 0.5
    bias = 0.1 * torch.randn(H, generator=gen, dtype=torch.float64)
    proj, Whd, Wd = (emb @ Wx.T + bias).to(DEV), Wh.to(DEV), W.to(DEV)
    ms = timed(lambda: nb.ops.lattice_beta_hat(p, proj, Whd, Wd), args.steps, True)
    log_beta, _ = nb.ops.lattice_beta_hat(p, proj, Whd, Wd)
    lat, src, dst, lab = (t.numpy() for t in (ab.arc_lattice, ab.src, ab.dst, ab.label))
    n_cpu = 2
    t0 = time.perf_counter()
    refs = [lo.beta_recurrent(int(ab.n_states[b]), src[lat == b], lab[lat == b], dst[lat == b], emb.numpy(), Wx.numpy(),
                              Wh.numpy(), W.numpy(), bias.numpy())[0] for b in range(n_cpu)]
    t_cpu = time.perf_counter() - t0
    arcs_cpu = int(sum((lat == b).sum() for b in range(n_cpu)))
    so = p.state_off.cpu().numpy()
    orig = p.orig_state.cpu().numpy()
    err = 0.0
    for b in range(n_cpu):
        got = log_beta[so[b]:so[b + 1]].cpu().numpy().astype(np.float64)
        err = max(err, float(np.max(np.abs(got - np.log(refs[b][orig[so[b]:so[b + 1]]])))))
    print(f"{'config1 beta-hat recurrence H=' + str(H) + ' B=' + str(B):44s} {p.n_arcs:11d} {p.n_states:10d} {p.max_levels:5d} "
          f"{'':7s} {ms:8.3f} {p.n_arcs / ms / 1e6:10.3f}  (compute_beta, Wh != 0; one launch per level) | "
          f"numpy float64 port {arcs_cpu / t_cpu / 1e6:.4f} Marc/s 1 thr | max |log beta - oracle| {err:.1e} (first {n_cpu} lattices)",
          flush=True)


def training_step_report():
    """Config 2 as the reference trains it: arc score = theta[label] (WFSTScorer, scorers.py:1663-1687),
    loss = -mean logZ, gradient d loss / d theta through the autograd Function (posterior label counts)."""
    ab = synth.snips_batch(256, seed=1).to(DEV)
    p, _ = ab.pack()
    theta = (-torch.rand(p.vocab, device=DEV)).requires_grad_(True)

    def step():
        theta.grad = None
        logz = nb.lattice_log_partition(p, theta=theta)
        (-logz.mean()).backward()

    ms = timed(step, args.steps, True)
    step()
    g = theta.grad.detach().cpu().numpy().astype(np.float64)
    # oracle: posteriors summed per label, on the first 16 lattices
    B = 16
    sel = (ab.arc_lattice < B).cpu()
    sub = synth.ArcBatch(ab.arc_lattice.cpu()[sel], ab.src.cpu()[sel], ab.dst.cpu()[sel], ab.label.cpu()[sel],
                         theta.detach().cpu()[ab.label.cpu()[sel]], ab.n_states.cpu()[:B], ab.vocab)
    ob = c_oracle.Batch(sub.arc_lattice.numpy(), sub.src.numpy(), sub.dst.numpy(), sub.label.numpy(), sub.scores.numpy(),
                        sub.n_states.numpy())
    _, _, _, o_post = c_oracle.forward_backward(ob)
    ref = np.bincount(sub.label.numpy(), weights=o_post, minlength=ab.vocab)
    p16, _ = sub.to(DEV).pack()
    th2 = theta.detach().clone().requires_grad_(True)
    nb.lattice_log_partition(p16, theta=th2).sum().backward()
    err = float(np.max(np.abs(th2.grad.cpu().numpy() - ref) / np.maximum(ref, 1e-3)))
    print(f"{'config2 SNIPS training step (theta mode) B=256':44s} {p.n_arcs:11d} {p.n_states:10d} {p.max_levels:5d} {'':7s} {ms:8.3f} "
          f"{p.n_arcs / ms / 1e6:10.2f}  (logZ forward + autograd backward -> d theta[V]; grad sum {g.sum():.3f}) | "
          f"d theta vs C oracle rel {err:.1e} (first {B} lattices)", flush=True)


def walk_step_report(B=32, k=16):
    """one time step of the sampling loop on config-1 lattices (B*k rows, V = 256): the fused
    kernel next to the reference's own sequence of eager torch ops on the same GPU (dense [B*k, S, V] tables,
    scorers.py:577-593, :683-690, :1037-1054, samplers.py:251-283)."""
    from nfst_b200.sampler import walk_step
    from torch.distributions import Categorical

    ab = synth.transliteration_batch(B, seed=0)
    p, _ = ab.to(DEV).pack()
    V, N = ab.vocab, B * k
    S = int(ab.n_states.max()) + 1
    # dense tables as the reference holds them (k-fold expanded, scorers.py:887-918)
    tr = torch.zeros(B, S, V, dtype=torch.int64)
    tr[ab.arc_lattice, ab.src, ab.label] = ab.dst
    tr_k = tr.to(DEV).repeat_interleave(k, 0)
    em_k = tr_k != 0
    beta_dense = torch.rand(N, S, device=DEV) + 0.1
    zero_v, ninf_v = torch.zeros(V, device=DEV), torch.full((V,), float("-inf"), device=DEV)
    prefix = torch.randn(N, V, device=DEV)
    base = torch.zeros(N, V, device=DEV)
    rows = torch.arange(N, device=DEV)
    state_dense = tr_k[rows, 0, synth.BOS]  # after bos
    look_dense = torch.zeros(N, dtype=torch.int64, device=DEV)

    def reference_ops():
        trans = tr_k[rows, look_dense]                       # scorers.py:586-589
        final = torch.gather(beta_dense, 1, trans) + prefix  # :590-592
        mask = torch.where(em_k[rows, state_dense], zero_v, ninf_v)  # :1042-1049
        masked = (final + base + mask) / 1.0
        dist = Categorical(logits=masked)
        sym = dist.sample()
        lp = dist.log_prob(sym)
        nxt = tr_k[rows, state_dense][rows, sym]             # :683-690
        return sym, lp, nxt

    # packed equivalents of the same states
    so = p.state_off.long()
    inv = torch.full((B, S), 0, dtype=torch.int64, device=DEV)
    lat = torch.repeat_interleave(torch.arange(B, device=DEV), (so[1:] - so[:-1]))
    inv[lat, p.orig_state.long()] = torch.arange(p.n_states, device=DEV)
    st = inv[rows // k, state_dense].to(torch.int32)
    lk = inv[rows // k, look_dense].to(torch.int32)
    beta_packed = beta_dense[lat * k, p.orig_state.long()].contiguous()
    u = torch.rand(N, device=DEV)

    def fused():
        return walk_step(p, k, st, prefix, beta_packed, synth.PAD, base_mask=base, uniform=u, look_state=lk)

    ms_ref = timed(reference_ops, args.steps * 5, False)
    ms_new = timed(fused, args.steps * 5, False)
    print(f"{'config1 sampling-loop step B=32 k=16':44s} rows {N}, V {V}: fused kernel {ms_new * 1e3:.1f} us/step, the reference's eager "
          f"torch ops on the same GPU {ms_ref * 1e3:.1f} us/step ({ms_ref / ms_new:.1f}x); dense tables {tr_k.numel() * 9 / 1e6:.0f} MB "
          f"vs packed arcs {p.n_arcs * 8 / 1e6:.2f} MB", flush=True)


def compute_beta_report(B=32, k=16):
    """The drop-in call itself on config 1 as the reference holds it: FSAGRUScorer.compute_beta() on collate()-padded
    dense [B, S, V] tables (scorers.py:753-875) -> nb.compute_beta(emission, transition, theta, k): dense scan + device
    packer + fused backward kernel + expansion to beta[B*k, S]; and the same with the pack cached."""
    ab = synth.transliteration_batch(B, seed=0)
    S, V = int(ab.n_states.max()), ab.vocab
    tr = torch.full((B, S, V), synth.PAD, dtype=torch.int64)
    for b in range(B):
        tr[b, : int(ab.n_states[b])] = 0
    tr[ab.arc_lattice, ab.src, ab.label] = ab.dst
    tr = tr.to(DEV)
    em = tr != 0
    theta = torch.randn(V, device=DEV)
    ms_full = timed(lambda: nb.compute_beta(em, tr, theta, k=k), args.steps * 3, False)
    p = nb.pack_dense(em, tr)
    ms_dp = timed(lambda: nb.compute_beta(em, tr, theta, k=k, packed=p), args.steps * 3, False)
    print(f"{'config1 compute_beta(em, tr, theta, k=16) drop-in call':44s} dense tables [{B}, {S}, {V}] ({tr.numel() * 8 / 1e6:.0f} MB), {p.n_arcs} arcs: "
          f"{ms_full:.3f} ms per call from the dense tables (scan + device pack + DP + beta[{B * k}, {S}]), {ms_dp:.3f} ms with the pack cached; "
          f"the unmodified reference's compute_beta() on lattices of this batch, timed in the build container (CPU, 1 core): 0.94-1.16 s "
          f"per lattice = ~600 arcs/s, i.e. ~{p.n_arcs / 600:.0f} s for the batch", flush=True)


q = args.quick
CONFIGS = [
    ("config1 transliteration B=32", lambda n, o: synth.transliteration_batch(n, seed=o), 32, 32),
    ("config2 SNIPS B=256", lambda n, o: synth.snips_batch(n, seed=1 + o), 256, 256),
    ("config3 cipher unigram T=1000 B=64", lambda n, o: synth.cipher_batch(n, T=1000, bigram=False, seed=2 + o, device=DEV), 64, 64),
    ("config3 cipher bigram T=1000 B=64", lambda n, o: synth.cipher_batch(n, T=1000, bigram=True, seed=2 + o, device=DEV), 16 if q else 64, 16),
    ("config5 Viterbi transliteration B=4096", lambda n, o: synth.transliteration_batch(n, seed=4 + o), 512 if q else 4096, 512),
    ("config5 Viterbi integer scores B=4096", lambda n, o: synth.transliteration_batch(n, seed=4 + o, integer_scores=True), 512 if q else 4096, 512),
]
for arcs in (() if args.no_dag else (10_000, 30_000, 100_000, 300_000, 1_000_000)):
    B = 1024 if not q else 128
    CONFIGS.append((f"config4 random DAG A={arcs} B={B}", (lambda a: lambda n, o: synth.random_dag_batch(n, a, seed=3 + o, device=DEV))(arcs),
                    B, max(1, 60_000_000 // arcs)))

print(f"{'config':44s} {'arcs':>11s} {'states':>10s} {'lvls':>5s} {'pack s':>7s} {'f+b ms':>8s} {'f+b Garc/s':>10s} {'alg GB/s':>8s} "
      f"{'vit ms':>8s} {'vit Garc/s':>10s} | {'CPU Marc/s':>10s} {'thr':>3s} | parity (GPU vs C oracle sample)")
for name, gen, B, per_chunk in CONFIGS:
    packed, sc, sample, t_pack = build(gen, B, per_chunk)
    A, S = packed.n_arcs, packed.n_states
    flush = 20 * A < 2 * 126e6

    all_sell = all(g.sell or g.tiles for g in packed.groups)  # column-major groups: two passes, no alpha
    beta_buf = torch.empty(S, dtype=nb.ops.resolve_state_dtype(packed), device=DEV) if packed.has_columns else None

    def fb():  # first pass: logZ (+ beta / alpha), second pass: posteriors (+ beta for CSR groups) -- bench.py's step
        lz, al, cond = nb.ops.lattice_pull(packed, arc_scores=sc, beta_out=beta_buf)
        nb.lattice_backward(packed, arc_scores=sc, alpha=al, logz=lz, cond=cond, want_beta=not all_sell, want_post=True)

    def vit():
        nb.lattice_viterbi(packed, arc_scores=sc)

    ms_fb = timed(fb, args.steps, flush)
    ms_v = timed(vit, args.steps, flush)
    par = parity(sample)
    name = name + (" [tiles]" if all(g.tiles for g in packed.groups) else " [sell]" if all_sell else
                   " [small]" if all(g.small_max_arcs > 0 for g in packed.groups) else "")
    print(f"{name:44s} {A:11d} {S:10d} {packed.max_levels:5d} {t_pack:7.2f} {ms_fb:8.3f} {A / ms_fb / 1e6:10.2f} "
          f"{(20 * A + 20 * S) / ms_fb / 1e6:8.0f} {ms_v:8.3f} {A / ms_v / 1e6:10.2f} | {par['cpu_arcs_per_s'] / 1e6:10.1f} {par['cpu_threads']:3d} | "
          f"state {par['state']} logZ rel {par['logz_rel']:.1e} post rel {par['post_rel']:.1e} (p >= 1e-4) abs {par['post_abs']:.1e} (p < 1e-4) "
          f"Viterbi scores {par['vit_scores_exact']} paths {par['vit_paths_exact']} (first {par['B']} lattices)", flush=True)
    del packed, sc, sample
    torch.cuda.empty_cache()
compute_beta_report()
recurrent_beta_report()
training_step_report()
walk_step_report()
