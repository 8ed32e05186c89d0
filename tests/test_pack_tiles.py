"""Tile-stream layout (nfst_b200/tiles.py), host side: the packer's tiles are walked in numpy with the same
addressing the kernels use (tests/tile_replay.py) and the results compared with the float64 oracle -- beta /
logZ, posteriors through the conditional-probability flow, Viterbi with first-label ties -- for several warp
counts, tile sizes, ring sizes (far destinations in the far table) and heavy states."""
import numpy as np
import pytest
import torch

import nfst_b200 as nb
from nfst_b200 import synth
from nfst_b200 import tiles as T
from nfst_b200.pack import concat_packed, pack_arcs
from oracle import lattice_oracle as lo
from tests.tile_replay import replay, seg_arcs, tiles_of


def _np(t):
    return t.cpu().numpy().astype(np.int64)


def check_lattices(ab, p, w):
    assert p.has_tiles
    w = w.numpy()
    src_out, dst, lab = _np(p.src_out), _np(p.dst_out), _np(p.label_out)
    origin = _np(p.arc_origin)
    assert sorted(origin.tolist()) == list(range(ab.src.numel()))
    np.testing.assert_array_equal(lab, ab.label.numpy()[origin])
    state_off, arc_off = _np(p.state_off), _np(p.arc_off)
    for g in p.groups:
        if not g.tiles:
            continue
        nw = g.block_threads // 32
        for b in g.ids.tolist():
            a0, a1 = arc_off[b], arc_off[b + 1]
            info = _np(p.tile_lat_info)[b]
            assert info[1] % 32 == 0 and info[1] < info[2] <= g.tile_ring and info[2] - info[1] - 1 <= g.tile_far
            # every warp's tile list ascends by level; tiles respect the group's capacities
            for wi, tl in tiles_of(p, b, nw).items():
                assert [t.level for t in tl] == sorted(t.level for t in tl)
                for t in tl:
                    assert t.n_arcs <= g.tile_cap_arcs and t.n_bytes <= g.tile_cap_bytes and t.n_bytes % 16 == 0
                    assert a0 <= t.arc0 and t.arc0 + t.n_arcs <= a1
                    for seg in t.segs:
                        rows = {}
                        for (lane, k, a, c) in seg_arcs(t, seg):
                            assert src_out[a] == seg["state0"] + lane
                            rows.setdefault(lane, []).append(lab[a])
                        for labs in rows.values():
                            assert labs == sorted(labs), "columns follow label order (Viterbi tie-break)"
            beta, cond, post, delta, bp, seen = replay(p, w, b, nw)
            assert (seen[a0:a1] == 1).all() and seen.sum() == a1 - a0, "every arc is visited exactly once"
            ns = state_off[b + 1] - state_off[b]
            la = np.arange(a0, a1)
            ls, ld = src_out[la] - state_off[b], dst[la] - state_off[b]
            start = int(p.start_state[b]) - state_off[b]
            o_logz, o_alpha, o_beta, o_post = lo.forward_backward(ns, ls, ld, w[la].astype(np.float64), start=start)
            np.testing.assert_allclose(beta[state_off[b]:state_off[b + 1]], o_beta, rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(post[la], o_post, rtol=1e-9, atol=1e-300)
            v_score, v_arcs, _, _, _ = lo.viterbi_f32(ns, ls, lab[la], ld, w[la], start=start)
            assert np.float32(delta[int(p.start_state[b])]) == np.float32(v_score)
            path, s = [], int(p.start_state[b])
            while bp[s] >= 0:
                path.append(bp[s] - a0)
                s = dst[bp[s]]
            assert path == list(v_arcs)


@pytest.mark.parametrize("arcs,levels,seed,warps", [(1500, 6, 0, 0), (4000, 10, 1, 1), (4000, 10, 1, 4), (900, 3, 2, 2),
                                                   (6000, 12, 3, 8)])
def test_tile_layout_replays_to_the_oracle(arcs, levels, seed, warps, monkeypatch):
    monkeypatch.setattr(T, "TILE_WARPS", warps)
    ab = synth.random_dag_batch(3, arcs, levels=levels, seed=seed)
    p, w = ab.pack(tiles=True)
    assert all(g.tiles for g in p.groups)
    check_lattices(ab, p, w)


def test_tile_small_ring_keeps_far_destinations_in_the_far_table(monkeypatch):
    ab = synth.random_dag_batch(2, 5000, levels=20, seed=5)
    monkeypatch.setattr(T, "FORCE_RING_SLICES", 12)  # wider than any level, shorter than most arcs
    monkeypatch.setattr(T, "TILE_WARPS", 1)
    p, w = ab.pack(tiles=True)
    assert p.has_tiles and any(g.tile_far for g in p.groups)
    assert int(p.tile_lat_info[:, 1].max()) <= 32 * 12 and min(g.tile_far for g in p.groups) > 100
    check_lattices(ab, p, w)


def test_tile_heavy_states_are_cut_into_pieces(monkeypatch):
    monkeypatch.setattr(T, "TILE_ARCS", 64)
    monkeypatch.setattr(T, "TILE_WARPS", 2)
    ab = synth.random_dag_batch(2, 4000, levels=6, seed=9)  # the source has hundreds of arcs
    p, w = ab.pack(tiles=True)
    deg = np.diff(_np(p.out_ptr))
    assert deg.max() > 3 * 64
    check_lattices(ab, p, w)
    # a star: one heavy state over a wide level, plus states with 9..32 arcs (extension blocks)
    n = 300
    src = [torch.zeros(n, dtype=torch.int64)]
    dst = [torch.arange(1, n + 1)]
    lab = [torch.arange(n) % 250 + 4]
    for s in range(1, n + 1):  # state s gets 1 + s % 33 arcs into the next level (labels distinct per state)
        k = 1 + s % 33
        src.append(torch.full((k,), s)); dst.append(n + 1 + (torch.arange(k) * 7 + s) % n); lab.append(torch.arange(k) + 4)
    src.append(torch.arange(n + 1, 2 * n + 1)); dst.append(torch.full((n,), 2 * n + 1)); lab.append(torch.full((n,), 3))
    src, dst, lab = torch.cat(src), torch.cat(dst), torch.cat(lab)
    g = torch.Generator().manual_seed(0)
    sc = torch.randn(src.numel(), generator=g)
    ab2 = synth.ArcBatch(torch.zeros(src.numel(), dtype=torch.int64), src, dst, lab, sc, torch.tensor([2 * n + 2]), 256)
    p2, w2 = ab2.pack(tiles=True)
    check_lattices(ab2, p2, w2)


def test_tiles_are_per_lattice_and_survive_concat(monkeypatch):
    wide = synth.random_dag_batch(2, 3000, levels=5, seed=5)
    narrow = synth.transliteration_batch(3, seed=1)
    pw, ww = wide.pack(tiles=True)
    pn, wn = narrow.pack(tiles=True)
    assert pw.has_tiles and not pn.has_tiles
    both = concat_packed([pn, pw, pw])
    kinds = sorted((g.tiles, g.n) for g in both.groups)
    assert (True, 4) in kinds and sum(n for t, n in kinds if not t) == 3
    ab = synth.ArcBatch(torch.cat([narrow.arc_lattice, wide.arc_lattice + 3, wide.arc_lattice + 5]),
                        torch.cat([narrow.src, wide.src, wide.src]), torch.cat([narrow.dst, wide.dst, wide.dst]),
                        torch.cat([narrow.label, wide.label, wide.label]), torch.cat([narrow.scores, wide.scores, wide.scores]),
                        torch.cat([narrow.n_states, wide.n_states, wide.n_states]), wide.vocab)
    both.arc_origin = torch.cat([pn.arc_origin, pw.arc_origin + narrow.src.numel(),
                                 pw.arc_origin + narrow.src.numel() + wide.src.numel()])
    check_lattices(ab, both, torch.cat([wn, ww, ww]))
    # opting out gives the other layouts
    pc, _ = wide.pack(tiles=False, sell=False)
    assert not pc.has_tiles and not pc.has_sell


def test_tile_default_thresholds():
    assert not synth.transliteration_batch(2, seed=3).pack()[0].has_tiles  # the reference's own narrow lattices
    p, _ = synth.random_dag_batch(1, 40_000, levels=16, seed=1).pack()  # 625 states per level: 19 slices
    assert p.has_tiles and p.groups[0].block_threads == 128
    p, _ = synth.random_dag_batch(1, 100_000, levels=64, seed=1).pack()  # the bench lattices: 13 slices per level
    assert p.has_tiles and p.groups[0].block_threads == 128
    p, _ = synth.random_dag_batch(1, 10_000, levels=64, seed=1).pack()  # 40 states per level: two slices
    assert p.has_tiles and p.groups[0].block_threads == 64
    p, _ = synth.random_dag_batch(1, 300_000, levels=64, seed=1).pack()  # 37 slices per level, ring of 35 KB: still 4 warps
    assert p.has_tiles and p.groups[0].block_threads == 128
    p, _ = synth.random_dag_batch(1, 1_000_000, levels=64, seed=1).pack()  # ring > 100 KB: one block per SM, 16 warps
    assert p.has_tiles and p.groups[0].block_threads == 512


def test_deep_lattices_get_rings_sized_for_float64_state():
    # more than 96 levels -> float64 state by default (ops.resolve_state_dtype): 8 bytes per ring slot
    nw = torch.tensor([4, 16])
    shallow, deep = T.ring_cap_slots(nw, torch.tensor([64, 64])), T.ring_cap_slots(nw, torch.tensor([128, 400]))
    assert bool((deep <= shallow).all()) and bool((deep * 8 <= T.SMEM_BUDGET).all()) and bool((deep >= 64).all())
    p, _ = synth.random_dag_batch(1, 60_000, levels=400, seed=9).pack()  # narrow and deep: still tile-stream
    assert p.has_tiles and (p.groups[0].tile_ring + 32) * 8 < 200 * 1024


def test_in_order_arrays_are_built_on_demand_and_equal_the_eager_pack():
    # a batch of column-major lattices is packed without the arcs-by-destination arrays and the chunk lists (only the
    # CSR forward kernel reads them: alpha); ensure_in_order() builds exactly what an eager pack would have
    ab = synth.random_dag_batch(3, 4000, levels=8, seed=2)
    lazy, _ = ab.pack()
    assert lazy.has_tiles and not lazy.has_in_order and lazy.in2out.numel() == 0 and lazy.fwd_chunks.shape[0] == 0
    eager, _ = ab.pack(in_order=True)
    assert eager.has_in_order
    two = concat_packed([lazy, ab.pack()[0]])
    assert not two.has_in_order and two.n_arcs == 2 * lazy.n_arcs  # lazy parts: the batch stays lazy
    lazy.ensure_in_order()
    for f in ("in2out", "src_in", "label_in", "in_ptr", "fwd_chunk_off", "fwd_chunks", "bwd_chunk_off", "bwd_chunks", "fwd_gather",
              "bwd_order", "fwd_chunk_level", "bwd_chunk_level"):
        assert torch.equal(getattr(lazy, f), getattr(eager, f)), f
    # next to a CSR lattice the in-order arrays are needed: concat builds them for the lazy part
    narrow, _ = synth.transliteration_batch(2, seed=1).pack()
    mixed = concat_packed([narrow, ab.pack()[0]])
    assert mixed.has_in_order and mixed.in2out.numel() == mixed.n_arcs
    two.ensure_in_order()
    ref = concat_packed([eager, eager])
    for f in ("in2out", "src_in", "fwd_chunks", "bwd_order"):
        assert torch.equal(getattr(two, f), getattr(ref, f)), f
    with pytest.raises(ValueError, match="column-major"):
        synth.transliteration_batch(2, seed=1).pack(in_order=False)
