"""Random (x, y) pairs through the on-device lattice construction (construct.edit_lattices) against the host restatement
(oracle/edit_lattice_oracle.py) and the C oracle: sizes, chains between grid states, logZ, best path.  Test
infrastructure (imports oracle/ and tests/).  python tests/fuzz/fuzz_construct.py [seconds] [seed]"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import nfst_b200 as nb  # noqa: E402
from nfst_b200 import construct  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle import edit_lattice_oracle as eo  # noqa: E402
from tests.test_gpu_construct import DEV, MARKS, V, grid_chains  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 45.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)


def run_case():
    B = int(rng.integers(1, 48))
    hi = int(rng.choice([3, 8, 16, 40]))
    add_sub = bool(rng.integers(0, 2))
    xs = [rng.integers(8, 30, size=int(rng.integers(0, hi + 1))).tolist() for _ in range(B)]
    ys = [rng.integers(30, 60, size=int(rng.integers(0, hi + 1))).tolist() for _ in range(B)]
    marks = dict(MARKS)
    if not add_sub:
        marks["sub_mark"] = None
    p = construct.edit_lattices(xs, ys, vocab=V, **marks)
    theta = torch.from_numpy(rng.normal(size=V).astype(np.float32))
    logz = nb.lattice_log_partition(p, theta=theta.to(DEV)).cpu().numpy()
    vs, voff, varcs, vlab = nb.lattice_viterbi(p, theta=theta.to(DEV))
    voff, vlab = voff.cpu().numpy(), vlab.cpu().numpy()
    so, ao = p.state_off.cpu().numpy(), p.arc_off.cpu().numpy()
    src_out, dst_out, lab_out, orig = (t.cpu().numpy() for t in (p.src_out, p.dst_out, p.label_out, p.orig_state))
    check = rng.choice(B, size=min(B, 6), replace=False)  # the host construction is slow Python: a sample per batch
    for b in check:
        x, y = xs[b], ys[b]
        arcs, S = eo.edit_lattice(x, y, **marks)
        assert (S, len(arcs)) == construct.edit_lattice_size(len(x), len(y), add_sub)
        assert (so[b + 1] - so[b], ao[b + 1] - ao[b]) == (S, len(arcs))
        a = slice(ao[b], ao[b + 1])
        got = list(zip(orig[src_out[a]].tolist(), lab_out[a].tolist(), orig[dst_out[a]].tolist()))
        assert grid_chains(got, len(x), len(y)) == grid_chains(arcs, len(x), len(y))
        s_, l_, d_ = (np.array(c) for c in zip(*arcs))
        ob = c_oracle.Batch(np.zeros(len(arcs), dtype=np.int64), s_, d_, l_, theta.numpy()[l_], [S])
        o_logz = c_oracle.forward_backward(ob, want_post=False, want_states=False)[0]
        np.testing.assert_allclose(logz[b], o_logz[0], rtol=1e-5, atol=1e-5)
        o_vs, _, o_vlab = c_oracle.viterbi(ob)
        assert vs[b:b + 1].cpu().numpy().view(np.uint32)[0] == o_vs.view(np.uint32)[0]
        assert vlab[voff[b]:voff[b + 1]].tolist() == list(o_vlab[0])
    return f"B={B} lengths<={hi} sub={add_sub} states={p.n_states} arcs={p.n_arcs} levels={p.max_levels}"


t0 = time.time()
n = fails = 0
while time.time() - t0 < budget:
    n += 1
    try:
        msg = run_case()
        if n % 20 == 1:
            print("ok  ", msg, flush=True)
    except Exception:  # noqa: BLE001
        fails += 1
        print(f"FAIL case {n}\n{traceback.format_exc()}", flush=True)
print(f"{n} batches of (x, y) pairs, {fails} failures, {time.time() - t0:.0f} s")
# one long pair: (|x|+1)(|y|+1) grid cells beyond the small-lattice kernels' 16-bit indices
try:
    p = construct.edit_lattices([list(range(8, 8 + 22)) * 6], [list(range(30, 30 + 20)) * 6], vocab=V, **MARKS)
    lz = nb.lattice_log_partition(p, theta=torch.zeros(V, device=DEV))
    print(f"long pair |x|=132 |y|=120: states={p.n_states} arcs={p.n_arcs} logZ={float(lz[0]):.4f} groups={[(g.tiles, g.small_max_arcs > 0, g.block_threads) for g in p.groups]}")
except Exception as e:  # noqa: BLE001
    print("long pair |x|=132 |y|=120:", type(e).__name__, e)
sys.exit(1 if fails else 0)
