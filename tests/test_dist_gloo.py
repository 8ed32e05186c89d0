"""Multi-rank host logic on CPU: gloo backend, world_size 2 (one process per rank)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nfst_b200 import dist as nd
from nfst_b200 import synth
from oracle import c_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ab = synth.transliteration_batch(9, seed=7)
        arc_counts = np.bincount(ab.arc_lattice.numpy(), minlength=9)
        mine = nd.my_shard(arc_counts)
        # this rank's lattices -> oracle logZ / per-label posterior mass (the GPU kernels' role)
        sel = np.isin(ab.arc_lattice.numpy(), mine)
        remap = {b: i for i, b in enumerate(mine)}
        lat = np.array([remap[b] for b in ab.arc_lattice.numpy()[sel]])
        ob = c_oracle.Batch(lat, ab.src.numpy()[sel], ab.dst.numpy()[sel], ab.label.numpy()[sel], ab.scores.numpy()[sel],
                            ab.n_states.numpy()[mine])
        logz, _, _, post = c_oracle.forward_backward(ob)
        dtheta = np.zeros(ab.vocab)
        np.add.at(dtheta, ab.label.numpy()[sel], post)
        loss, grad = nd.all_reduce_loss_and_grad(torch.tensor(logz.sum(), dtype=torch.float32),
                                                 torch.from_numpy(dtheta).float())
        # the same sums through the asynchronous form (what bench.py's step uses)
        pend = nd.all_reduce_loss_and_grad(torch.tensor(logz.sum(), dtype=torch.float32), torch.from_numpy(dtheta).float(),
                                           async_op=True)
        loss2, grad2 = pend.wait()
        assert float(loss2) == float(loss) and torch.equal(grad2, grad)
        scores = nd.gather_ragged(torch.from_numpy(logz).float(), mine, 9)
        q.put((rank, mine, float(loss), grad.numpy(), scores.numpy()))
    finally:
        dist.destroy_process_group()


def test_sharding_and_single_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, l0, g0, s0), (r1, m1, l1, g1, s1) = res
    assert sorted(m0 + m1) == list(range(9)) and not set(m0) & set(m1)  # a partition
    # single-process reference
    ab = synth.transliteration_batch(9, seed=7)
    ob = c_oracle.Batch(ab.arc_lattice.numpy(), ab.src.numpy(), ab.dst.numpy(), ab.label.numpy(), ab.scores.numpy(),
                        ab.n_states.numpy())
    logz, _, _, post = c_oracle.forward_backward(ob)
    dtheta = np.zeros(ab.vocab)
    np.add.at(dtheta, ab.label.numpy(), post)
    assert abs(l0 - logz.sum()) < 1e-3 and l0 == l1  # every rank holds the reduced loss
    np.testing.assert_allclose(g0, dtheta, rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(g0, g1)
    np.testing.assert_allclose(s0, logz, rtol=1e-6)
    np.testing.assert_array_equal(s0, s1)


def test_shard_by_arcs_balances_and_is_deterministic():
    rng = np.random.default_rng(0)
    counts = rng.integers(100, 100_000, size=257).tolist()
    for world in (1, 2, 4, 8):
        bins = nd.shard_by_arcs(counts, world)
        assert sorted(i for b in bins for i in b) == list(range(257))
        loads = [sum(counts[i] for i in b) for b in bins]
        assert max(loads) - min(loads) <= max(counts)
        assert bins == nd.shard_by_arcs(counts, world)
    with pytest.raises(ValueError):
        nd.shard_by_arcs(counts, 0)
    assert nd.shard_by_arcs([], 2) == [[], []]
    assert nd.my_shard(counts) == list(range(257))  # no process group: everything is local
