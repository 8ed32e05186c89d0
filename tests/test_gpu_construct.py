"""On-device construction of the transliteration lattices (nfst_b200.construct.edit_lattices, SURVEY section 8 row
f-4) against the host restatement of the reference's construction (oracle/edit_lattice_oracle.py): the same arcs
(as (source, label, destination) triples -- both number the grid row-major, so the grid states coincide and the
chains are compared as label strings between grid states), the same logZ / posteriors / best path through the C
oracle, and the closed-form sizes."""
import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import construct
from nfst_b200 import pack as P
from oracle import c_oracle
from oracle import edit_lattice_oracle as eo

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"
V = 64
MARKS = dict(bos=1, eos=2, input_mark=5, output_mark=4, sub_mark=6)


def pairs(rng, B, lo=0, hi=7):
    return ([rng.integers(8, 30, size=int(rng.integers(lo, hi + 1))).tolist() for _ in range(B)],
            [rng.integers(30, 60, size=int(rng.integers(lo, hi + 1))).tolist() for _ in range(B)])


def grid_chains(arcs, n, m):
    """{(grid source, grid destination): sorted list of label strings of the chains between them}; the start state,
    the grid ((n+1)(m+1) states from id 1, row-major) and the sink count as grid states."""
    S = 1 + max(max(s, d) for s, _, d in arcs)
    G = (n + 1) * (m + 1)
    is_grid = lambda s: s <= G or s == S - 1  # noqa: E731
    out_of = {}
    for s, l, d in arcs:
        out_of.setdefault(s, []).append((l, d))
    res = {}
    for s in [q for q in out_of if is_grid(q)]:
        for l, d in out_of[s]:
            labs, cur = [l], d
            while not is_grid(cur):
                assert len(out_of[cur]) == 1
                l2, cur = out_of[cur][0]
                labs.append(l2)
            res.setdefault((s, cur), []).append(tuple(labs))
    return {k: sorted(v) for k, v in res.items()}


@pytest.mark.parametrize("add_sub", [True, False])
def test_device_built_lattices_equal_the_host_construction(add_sub):
    rng = np.random.default_rng(1 + add_sub)
    xs, ys = pairs(rng, 24)
    xs[0], ys[0] = [], []  # the empty pair: bos, eos
    xs[1], ys[1] = [9, 9, 9], []
    marks = dict(MARKS)
    if not add_sub:
        marks["sub_mark"] = None
    before = P.launch_count
    p = construct.edit_lattices(xs, ys, vocab=V, **marks)
    assert P.launch_count == before + 4 and all(g.small_max_arcs > 0 for g in p.groups), "packed on the device"
    theta = torch.randn(V, generator=torch.Generator().manual_seed(3))
    logz = nb.lattice_log_partition(p, theta=theta.to(DEV)).cpu().numpy()
    vs, voff, varcs, vlab = nb.lattice_viterbi(p, theta=theta.to(DEV))
    voff, vlab = voff.cpu().numpy(), vlab.cpu().numpy()
    so, ao = p.state_off.cpu().numpy(), p.arc_off.cpu().numpy()
    src_out, dst_out, lab_out, orig = (t.cpu().numpy() for t in (p.src_out, p.dst_out, p.label_out, p.orig_state))
    for b, (x, y) in enumerate(zip(xs, ys)):
        arcs, S = eo.edit_lattice(x, y, **marks)
        assert (S, len(arcs)) == construct.edit_lattice_size(len(x), len(y), add_sub)
        assert (so[b + 1] - so[b], ao[b + 1] - ao[b]) == (S, len(arcs)), "every state is reachable: nothing is trimmed"
        a = slice(ao[b], ao[b + 1])
        got = list(zip(orig[src_out[a]].tolist(), lab_out[a].tolist(), orig[dst_out[a]].tolist()))
        assert grid_chains(got, len(x), len(y)) == grid_chains(arcs, len(x), len(y))
        # the DP on the device-built lattice == the C oracle on the host-built one
        s_, l_, d_ = (np.array(c) for c in zip(*arcs))
        ob = c_oracle.Batch(np.zeros(len(arcs), dtype=np.int64), s_, d_, l_, theta.numpy()[l_], [S])
        o_logz = c_oracle.forward_backward(ob, want_post=False, want_states=False)[0]
        np.testing.assert_allclose(logz[b], o_logz[0], rtol=1e-5, atol=1e-5)
        o_vs, _, o_vlab = c_oracle.viterbi(ob)
        assert vs[b:b + 1].cpu().numpy().view(np.uint32)[0] == o_vs.view(np.uint32)[0]
        assert vlab[voff[b]:voff[b + 1]].tolist() == list(o_vlab[0])


def test_logz_counts_the_alignments():
    # theta = 0: Z is the number of mark strings = the number of alignments (brute force on small pairs)
    xs, ys = [[8, 9], [8], [8, 9, 10]], [[30], [30, 31, 32], [30, 31]]
    p = construct.edit_lattices(xs, ys, vocab=V, **MARKS)
    logz = nb.lattice_log_partition(p, theta=torch.zeros(V, device=DEV)).cpu().numpy()
    want = [len(eo.mark_strings(x, y, **MARKS)) for x, y in zip(xs, ys)]
    np.testing.assert_allclose(np.exp(logz), want, rtol=1e-5)


def test_tensor_inputs_and_argument_checks():
    x = torch.tensor([[8, 9, 0], [10, 11, 12]], device=DEV)
    y = torch.tensor([[30, 0], [31, 32]], device=DEV)
    p = construct.edit_lattices(x, y, x_len=torch.tensor([2, 3]), y_len=torch.tensor([1, 2]), vocab=V, **MARKS)
    q = construct.edit_lattices([[8, 9], [10, 11, 12]], [[30], [31, 32]], vocab=V, **MARKS)
    assert torch.equal(p.label_out, q.label_out) and torch.equal(p.dst_out, q.dst_out)
    with pytest.raises(RuntimeError, match="distinct"):
        construct.edit_lattices([[8]], [[30]], vocab=V, bos=1, eos=2, input_mark=5, output_mark=5)
    with pytest.raises(ValueError, match="exceeds"):
        construct.edit_lattices(x, y, x_len=torch.tensor([4, 3]), y_len=torch.tensor([1, 2]), vocab=V, **MARKS)
