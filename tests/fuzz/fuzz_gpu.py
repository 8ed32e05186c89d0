"""Random mixed batches through every execution model against the C oracle (test infrastructure: imports oracle/ and
the parity helpers of tests/): transliteration / SNIPS / cipher / random-DAG lattices of random sizes concatenated into
one batch, forward-backward (logZ, alpha, beta, posteriors: the tests' tolerances), Viterbi (bit-exact scores, equal
paths) and the theta-mode gradient.  python tests/fuzz/fuzz_gpu.py [seconds] [seed]"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import synth  # noqa: E402
from oracle import c_oracle  # noqa: E402
from tests.test_gpu_parity import DEV, oracle_batch  # noqa: E402

rng = np.random.default_rng(0)  # main() / the caller reseeds: fuzz_gpu.rng = np.random.default_rng(seed)


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


def run_case(names, parts):
    """one random batch: packed jointly or part by part (concat_packed, the per-step collate of cached examples), scored
    per arc / by theta / both, sometimes with integer scores (exact Viterbi ties)"""
    from nfst_b200.pack import concat_packed
    from tests.test_gpu_parity import gpu_state_to_orig

    ab = cat(list(parts))
    A = ab.src.numel()
    if rng.integers(0, 4) == 0:
        ab.scores = -torch.from_numpy(rng.integers(0, 3, size=A)).float()
    if rng.integers(0, 2) and len(parts) > 1:
        how = "concat"
        packs, off, origin = [], 0, []
        for a in parts:
            a.vocab = ab.vocab  # concat_packed: all parts share one vocabulary
            pk, _ = a.to(DEV).pack()
            packs.append(pk)
            origin.append(pk.arc_origin.cpu() + off)
            off += a.src.numel()
        p, origin = concat_packed(packs), torch.cat(origin).numpy()
    else:
        how = "joint"
        p, _ = ab.to(DEV).pack()
        origin = p.arc_origin.cpu().numpy()
    mode = ["arcs", "theta", "both"][int(rng.integers(0, 3))]
    theta = torch.from_numpy(rng.normal(size=p.vocab).astype(np.float32) * 0.3)
    lab = ab.label.numpy()
    w = np.zeros(A, dtype=np.float32)
    if mode != "theta":
        w = w + ab.scores.numpy()
    if mode != "arcs":
        w = (w + theta.numpy()[lab]).astype(np.float32)
    sc = torch.from_numpy(ab.scores.numpy()[origin]).to(DEV) if mode != "theta" else None
    th = theta.to(DEV) if mode != "arcs" else None
    ab2 = synth.ArcBatch(ab.arc_lattice, ab.src, ab.dst, ab.label, torch.from_numpy(w), ab.n_states, ab.vocab)
    ob = oracle_batch(ab2)
    o_logz, o_alpha, o_beta, o_post = c_oracle.forward_backward(ob)
    st = "auto" if rng.integers(0, 5) else torch.float64
    out = nb.lattice_forward_backward(p, arc_scores=sc, theta=th, want_dtheta=th is not None, state_dtype=st)
    logz, alpha, beta, post = (t.cpu().numpy().astype(np.float64) for t in out[:4])
    g2o = gpu_state_to_orig(p, ab.n_states.numpy())
    np.testing.assert_allclose(logz, o_logz, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(alpha, o_alpha[g2o], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(beta, o_beta[g2o], rtol=1e-5, atol=1e-5)
    ref = o_post[origin]
    bad = np.abs(post - ref) > 1e-5 * ref + 1e-7
    assert not bad.any(), f"posteriors: {int(bad.sum())} arcs beyond 1e-5 (worst {float(np.max(np.abs(post - ref) / (ref + 1e-7))):.2e})"
    if th is not None:
        want = np.zeros(p.vocab)
        np.add.at(want, lab, o_post)
        # (level-major groups flush one partial histogram per chunk into dtheta with float atomics: thousands of float adds
        # per label, 2.7e-5 measured on a 200-level bigram cipher lattice; inside a block the sums are fixed point)
        np.testing.assert_allclose(out[4].cpu().numpy(), want, rtol=5e-5, atol=5e-5)
    score, off, arcs, labels = nb.lattice_viterbi(p, arc_scores=sc, theta=th)
    o_score, o_paths, o_labels = c_oracle.viterbi(ob)
    assert np.array_equal(score.cpu().numpy().view(np.uint32), o_score.view(np.uint32)), "Viterbi scores differ"
    offc, arcs_c, lab_c = off.cpu().numpy(), arcs.cpu().numpy(), labels.cpu().numpy()
    for b in range(p.n_lattices):
        assert np.array_equal(origin[arcs_c[offc[b]:offc[b + 1]]], o_paths[b]), f"Viterbi path of lattice {b} differs"
        assert np.array_equal(lab_c[offc[b]:offc[b + 1]], o_labels[b])
    kinds = sorted({"tiles" if g.tiles else "sell" if g.sell else "small" if g.small_max_arcs > 0 else
                    "level" if g.fwd_level_chunks is not None else "csr" for g in p.groups})
    return f"[{', '.join(kinds)}; {how}, {mode}, state {str(alpha.dtype) if False else out[1].dtype}; {p.n_arcs} arcs, {p.max_levels} levels]"


def main(budget: float, seed: int) -> int:
    global rng
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = fails = refused = 0
    while time.time() - t0 < budget:
        names, parts = zip(*[part() for _ in range(int(rng.integers(1, 4)))])
        n += 1
        try:
            print(f"ok   {' + '.join(names)}  {run_case(names, parts)}", flush=True)
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
    print(f"{n} batches, {fails} failures, {refused} refused for shared memory, {time.time() - t0:.0f} s")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main(float(sys.argv[1]) if len(sys.argv) > 1 else 90.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0))
