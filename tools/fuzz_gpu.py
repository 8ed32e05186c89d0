"""Random mixed batches through every execution model against the C oracle (test infrastructure: imports oracle/ and
the parity helpers of tests/): transliteration / SNIPS / cipher / random-DAG lattices of random sizes concatenated into
one batch, forward-backward (logZ, alpha, beta, posteriors: the tests' tolerances), Viterbi (bit-exact scores, equal
paths) and the theta-mode gradient.  python tools/fuzz_gpu.py [seconds] [seed]"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import synth  # noqa: E402
from oracle import c_oracle  # noqa: E402
from tests.test_gpu_parity import DEV, check_fwd_bwd, oracle_batch  # noqa: E402
from tests.test_gpu_tiles import viterbi_matches  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 90.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def cat(parts):
    lat, off = [], 0
    for ab in parts:
        lat.append(ab.arc_lattice.cpu() + off)
        off += int(ab.n_states.numel())
    return synth.ArcBatch(torch.cat(lat), torch.cat([a.src.cpu() for a in parts]), torch.cat([a.dst.cpu() for a in parts]),
                          torch.cat([a.label.cpu() for a in parts]), torch.cat([a.scores.cpu() for a in parts]),
                          torch.cat([a.n_states.cpu() for a in parts]), max(a.vocab for a in parts))


def part():
    k = int(rng.integers(0, 4))
    s = int(rng.integers(0, 10**6))
    if k == 0:
        B = int(rng.integers(1, 40))
        return f"translit x{B} s{s}", synth.transliteration_batch(B, seed=s)
    if k == 1:
        B = int(rng.integers(1, 8))
        return f"snips x{B} s{s}", synth.snips_batch(B, seed=s)
    if k == 2:
        B, T, bi = int(rng.integers(1, 3)), int(rng.choice([40, 200])), bool(rng.integers(0, 2))
        return f"cipher{'-bi' if bi else ''} T={T} x{B} s{s}", synth.cipher_batch(B, T=T, bigram=bi, seed=s)
    B = int(rng.integers(1, 6))
    arcs = int(np.exp(rng.uniform(np.log(500), np.log(400_000))))
    levels = int(rng.choice([3, 8, 24, 64, 130, 300]))
    return f"dag {arcs} arcs, {levels} levels x{B} s{s}", synth.random_dag_batch(B, arcs, levels=levels, seed=s)


t0 = time.time()
n = fails = refused = 0
while time.time() - t0 < budget:
    names, parts = zip(*[part() for _ in range(int(rng.integers(1, 4)))])
    ab = cat(list(parts))
    n += 1
    try:
        p, sc, _ = check_fwd_bwd(ab)
        viterbi_matches(ab, p, sc)
        theta = torch.from_numpy(rng.normal(size=p.vocab).astype(np.float32) * 0.3)
        w = theta.numpy()[ab.label.numpy()]
        ab2 = synth.ArcBatch(ab.arc_lattice, ab.src, ab.dst, ab.label, torch.from_numpy(w), ab.n_states, ab.vocab)
        o_logz, _, _, o_post = c_oracle.forward_backward(oracle_batch(ab2))
        th = theta.to(DEV).requires_grad_(True)
        logz = nb.lattice_log_partition(p, theta=th)
        logz.sum().backward()
        np.testing.assert_allclose(logz.detach().cpu().numpy(), o_logz, rtol=1e-5, atol=1e-5)
        want = np.zeros(p.vocab)
        np.add.at(want, ab.label.numpy(), o_post)
        # (a label's gradient sums thousands of posteriors; the CSR kernels add them with float32 atomics: 2.4e-5 measured
        # on 200-level bigram cipher lattices -- the tile-stream kernels accumulate in fixed point)
        np.testing.assert_allclose(th.grad.cpu().numpy(), want, rtol=5e-5, atol=5e-5)
        kinds = sorted({"tiles" if g.tiles else "sell" if g.sell else "small" if g.small_max_arcs > 0 else
                        "level" if g.fwd_level_chunks is not None else "csr" for g in p.groups})
        print(f"ok   {' + '.join(names)}  [{', '.join(kinds)}; {p.n_arcs} arcs, {p.max_levels} levels]", flush=True)
    except RuntimeError as e:
        if "shared memory" in str(e):
            refused += 1
            print(f"REFUSED {' + '.join(names)}: {e}", flush=True)
        else:
            fails += 1
            print(f"FAIL {' + '.join(names)}\n{traceback.format_exc()}", flush=True)
    except Exception:  # noqa: BLE001
        fails += 1
        print(f"FAIL {' + '.join(names)}\n{traceback.format_exc()}", flush=True)
        try:  # which lattices / launch groups carry the posterior error
            p, sc = ab.to(DEV).pack()
            logz, alpha, beta, post = nb.lattice_forward_backward(p, arc_scores=sc)
            _, o_alpha, o_beta, o_post = c_oracle.forward_backward(oracle_batch(ab))
            ref = o_post[p.arc_origin.cpu().numpy()]
            got = post.cpu().numpy().astype(np.float64)
            bad = np.abs(got - ref) > 1e-5 * ref + 1e-7
            aoff = p.arc_off.cpu().numpy()
            kind = {}
            for g in p.groups:
                for b in g.ids.cpu().tolist():
                    kind[b] = "tiles" if g.tiles else "sell" if g.sell else "small" if g.small_max_arcs > 0 else "csr/level"
            for b in range(p.n_lattices):
                nb_ = int(bad[aoff[b]:aoff[b + 1]].sum())
                if nb_:
                    r = ref[aoff[b]:aoff[b + 1]]; d = np.abs(got[aoff[b]:aoff[b + 1]] - r)
                    i = int(np.argmax(d / (r + 1e-7)))
                    print(f"   lattice {b} [{kind[b]}] {aoff[b + 1] - aoff[b]} arcs, levels {int(p.n_levels[b])}, state {alpha.dtype}: {nb_} bad arcs, worst ref {r[i]:.3e} got-ref {d[i]:.3e}; "
                          f"max|alpha|+|beta| {float(np.abs(o_alpha).max() + np.abs(o_beta).max()):.1f}", flush=True)
        except Exception as e2:  # noqa: BLE001
            print("   (diagnosis failed:", e2, ")")
print(f"{n} batches, {fails} failures, {refused} refused for shared memory, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
