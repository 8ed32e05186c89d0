"""Tile-stream layout of wide lattices (kernels: nfst_b200/csrc/nfst_tiles.cu).

Same recurrence as everywhere else in this package (the reference's ``compute_beta_per_sample``,
``src/modules/scorers.py:692-751``, Wh = 0); what changes is how a lattice's arcs lie in HBM and how a
thread block walks them.  A lattice is dealt to the ``nw`` warps of its block AT PACK TIME:

* the states of a topological level are sorted by out-degree (descending) and cut into SLICES of 32 -- lane i
  of a warp owns state i of a slice; a state with more than ``TAILMAX`` arcs is HEAVY and forms a slice of its
  own whose arcs the whole warp strides over;
* the slices of a level are dealt to the warps in serpentine rounds and the states are numbered so that
  (level, warp, slice, lane) is the state order: what one warp does in one level is a contiguous run of states
  AND of arcs;
* the arcs of a slice are column-major without padding (column k = the k-th arc, in label order, of every
  state of the slice that has one); consecutive slices of one (level, warp) are grouped into TILES of about
  ``TILE_ARCS`` arcs -- the unit one warp fetches with ONE bulk copy (TMA) per array;
* per tile, the packer-owned byte stream holds a 16-byte tile header, one 16-byte header per slice (column
  counts, offsets, flags), 32-byte extension blocks for slices with more than 8 columns, and the arcs'
  destinations as 16-bit RING SLOTS: the DP values of recent states live in a shared-memory ring of ``W``
  slots (32 per slice, in slice order), slot ``W`` is a constant (beta = 0 / a dump for flow that nobody
  reads) used for arcs into the last level, and a destination that has already left the ring when one of its
  sources is processed additionally owns a slot of a small FAR TABLE behind the ring (``W+1 ...``: written once,
  never recycled) -- so every arc reads ``ring[code]``, whatever its length.

``tile_tab`` lists, per (lattice, warp), the warp's tiles in level order: {first canonical arc, stream offset,
sizes, level}.  Everything in the stream is lattice-relative, so independently packed batches concatenate
by offsetting ``tile_tab`` only.
"""
from __future__ import annotations

import os
from typing import Dict

import torch


def _phase(name, dev=None):
    from .pack import _phase as p

    p(name, dev)


TILES = int(os.environ.get("NFST_TILES", "1"))
# lattices whose levels average at least this many states take the tile-stream path
TILE_MIN_WIDTH = int(os.environ.get("NFST_TILE_MIN_WIDTH", "32"))
TILE_ARCS = int(os.environ.get("NFST_TILE_ARCS", "384"))  # target arcs per tile (a multiple of 32)
TILE_SLICES = int(os.environ.get("NFST_TILE_SLICES", "6"))  # slices per tile, at most
TILE_WARPS = int(os.environ.get("NFST_TILE_WARPS", "0"))  # 0 = from the level width
TILE_BLOCK_ARCS = int(os.environ.get("NFST_TILE_BLOCK_ARCS", "1024"))  # arcs in the tiles that a block's warps hold at a time
KU = 8  # columns of a slice the kernels hold in registers
# states with more arcs than this are heavy (a slice of their own).  Measured at 8 / 12 / 16 on B200: never faster than
# 32, and at 8 every 9+-arc state costs 32 ring slots
TAILMAX = int(os.environ.get("NFST_TILE_TAILMAX", "32"))
RING_MAX = int(os.environ.get("NFST_TILE_RING_MAX", "49152"))  # ring slots (float32: 192 KB)
NW_MAX = 32
FORCE_RING_SLICES = int(os.environ.get("NFST_TILE_FORCE_RING_SLICES", "0"))  # tests: a ring this short (more far destinations)

FLAG_FAR_IN, FLAG_HEAVY, FLAG_HEAVY_FIRST, FLAG_HEAVY_LAST = 2, 4, 8, 16


def _excl_cumsum(x: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(x.numel() + 1, dtype=torch.int64, device=x.device)
    torch.cumsum(x, 0, out=out[1:])
    return out


# shared memory of one block: the DP ring (4 bytes per slot; 8 with float64 state) and at least two stages per warp
SMEM_BUDGET = int(os.environ.get("NFST_TILE_SMEM_BUDGET", str(216 * 1024)))
NW_BASE = int(os.environ.get("NFST_TILE_WARPS_BASE", "4"))  # warps of a block when several blocks fit an SM
MIN_SM_WARPS = int(os.environ.get("NFST_TILE_MIN_SM_WARPS", "12"))  # wanted resident warps per SM (blocks x warps)


def stage_est(nw: torch.Tensor) -> torch.Tensor:
    """Bytes of one stage of an nw-warp block, roughly: the stream part of a tile (headers + 2 bytes per arc) and one
    staged 4-byte array (plus its alignment and over-read slack)."""
    return 7 * tile_arcs_for(nw) + 512


# lattices with more than this many levels run with float64 state vectors by default (ops.resolve_state_dtype): their
# ring slots are 8 bytes
F64_LEVELS = 96


def ring_cap_slots(nw: torch.Tensor, levels: torch.Tensor = None) -> torch.Tensor:
    """Largest DP ring (slots, far table included) that fits beside the stages of an nw-warp block; ``levels``: the
    lattices' level counts (deep lattices keep float64 state: half as many slots fit)."""
    elem = 4 if levels is None else torch.where(levels > F64_LEVELS, torch.full_like(nw, 8), torch.full_like(nw, 4))
    return torch.clamp(torch.div(SMEM_BUDGET - 2 * stage_est(nw) * nw, elem, rounding_mode="floor"), min=64, max=min(RING_MAX, 65000))


def warps_per_lattice(states: torch.Tensor, levels: torch.Tensor, span: torch.Tensor = None) -> torch.Tensor:
    """Warps per block of a tile-stream lattice (a power of two, 1..32).  Measured on B200 (tools/tile_tune.py): four
    warps per lattice win whenever several blocks fit an SM -- from 4 to 40 slices per level; what keeps HBM busy is
    the depth of the warps' bulk-copy rings, not the warp count -- so the default is NW_BASE, fewer only where a
    level has fewer slices than that.  A lattice whose DP ring leaves room for just one or two blocks per SM gets
    more warps (with smaller tiles: tile_arcs_for) until MIN_SM_WARPS warps are resident.  ``span``: how many states
    nearly all arcs of the lattice reach across (default: eight levels)."""
    if TILE_WARPS:
        return torch.full_like(states, TILE_WARPS)
    width = states.to(torch.float64) / torch.clamp(levels, min=1).to(torch.float64)
    slices = torch.clamp(torch.ceil(width / 32.0), min=1.0)
    base = torch.ones_like(states) << torch.clamp(torch.ceil(torch.log2(slices)).to(torch.int64), 0, int(NW_BASE).bit_length() - 1)
    elem = torch.where(levels > F64_LEVELS, 8.0, 4.0).to(torch.float64)  # deep lattices: float64 state
    ring = elem * (8 * width if span is None else 1.5 * span.to(torch.float64))  # bytes
    best = base
    done = torch.zeros_like(states, dtype=torch.bool)
    for lg in range(0, 6):
        nw = torch.full_like(states, 1 << lg)
        smem = ring + (2 * stage_est(nw) * nw).to(torch.float64)
        blocks = torch.clamp(torch.floor(SMEM_BUDGET / smem), max=float(2048 // (32 << lg)))
        blocks = torch.minimum(blocks, torch.full_like(blocks, 7.0))
        ok = (nw >= base) & (blocks >= 1) & ~done
        best = torch.where(ok, nw, best)  # the largest block that still fits, until one meets the target
        done = done | (ok & (blocks * nw.to(torch.float64) >= MIN_SM_WARPS))
    return best


def span_quantile(lat: torch.Tensor, span: torch.Tensor, n_lattices: int, q: float = 0.995) -> torch.Tensor:
    """Per-lattice q-quantile of the arcs' reach (in states), from a histogram of quarter octaves (at most 19 % high)."""
    bucket = torch.clamp(torch.ceil(4.0 * torch.log2(torch.clamp(span, min=1).to(torch.float64))).to(torch.int64), 0, 127)
    cum = torch.cumsum(torch.bincount(lat * 128 + bucket, minlength=n_lattices * 128).view(n_lattices, 128), 1)
    need = torch.ceil(cum[:, -1:].to(torch.float64) * q).to(torch.int64)
    return torch.pow(2.0, (cum < need).sum(1).to(torch.float64) / 4.0).to(torch.int64)


def tile_arcs_for(nw: torch.Tensor) -> torch.Tensor:
    """Target arcs per tile for lattices dealt to ``nw`` warps: TILE_ARCS for narrow blocks (few warps per SM: each
    bulk copy must be a kilobyte or two to keep HBM busy), smaller for wide blocks (many warps in flight, and
    nw x stages x tile bytes must fit shared memory next to the ring); a multiple of 32, at least 96.  Blocks of
    8 warps and more exist for lattices with a large DP ring (one block per SM): they get twice the budget."""
    budget = torch.where(nw >= 8, torch.full_like(nw, 2 * TILE_BLOCK_ARCS), torch.full_like(nw, TILE_BLOCK_ARCS))
    t = torch.clamp(torch.div(budget, torch.clamp(nw, min=1), rounding_mode="floor"), min=min(96, TILE_ARCS), max=TILE_ARCS)
    return torch.div(t, 32, rounding_mode="floor") * 32


def deal_order(order, lt, deg_out, slot_kept, level_start, tile_lat, tile_nw):
    """Refine the state order of tile-stream lattices.

    ``order`` sorts the kept states by (lattice, level, out-degree descending); returned is the final order
    (lattice, level, warp, slice, lane) plus, per state in that order, its warp, slice index in the level and lane.
    Other lattices keep their order (warp 0, lane 0)."""
    dev = order.device
    n = order.numel()
    lt_p, slot_p, deg_p = lt[order], slot_kept[order], deg_out[order]
    tile_p = tile_lat[lt_p]
    r = torch.arange(n, device=dev) - level_start[slot_p]
    heavy = tile_p & (deg_p > TAILMAX)
    h = torch.bincount(slot_p[heavy], minlength=int(level_start.numel()))[slot_p]
    j = torch.where(r < h, r, h + torch.div(r - h, 32, rounding_mode="floor"))
    lane = torch.where(r < h, torch.zeros_like(r), (r - h) % 32)
    nw = tile_nw[lt_p]
    rnd, ps = torch.div(j, nw, rounding_mode="floor"), j % nw
    w = torch.where(rnd % 2 == 0, ps, nw - 1 - ps)
    zero = torch.zeros_like(r)
    w, lane, j = torch.where(tile_p, w, zero), torch.where(tile_p, lane, zero), torch.where(tile_p, j, r)
    jmax = int(j.max()) + 1 if n else 1
    if int(level_start.numel()) * NW_MAX * jmax * 32 >= 2**62:
        raise ValueError("batch too large to deal into tiles; shard it")
    key = ((slot_p * NW_MAX + w) * jmax + j) * 32 + lane
    o2 = torch.argsort(key, stable=True)
    return order[o2], w[o2], j[o2], lane[o2]


def build(*, lt_s, slot, level_off, level_ptr, n_levels, st_w, st_j, out_ptr, out_deg, src_out, dst_out, tile_lat, tile_nw,
          state_off, n_lattices: int) -> Dict[str, torch.Tensor]:
    """Slices, ring slots, tiles and the byte stream of the tile-stream lattices of a batch (states and arcs
    already in their final order)."""
    dev = lt_s.device
    B = n_lattices
    i64 = dict(dtype=torch.int64, device=dev)
    assert TILE_ARCS % 32 == 0 and 32 <= TILE_ARCS <= 4096
    T_lat = tile_arcs_for(tile_nw)  # [B] target arcs per tile
    ts = torch.nonzero(tile_lat[lt_s]).squeeze(1)  # tile states, ascending
    empty = {
        "tile_stream": torch.zeros(16, dtype=torch.uint8, device=dev), "tile_tab": torch.zeros((0, 4), dtype=torch.int32, device=dev),
        "tile_lw_off": torch.zeros(1, dtype=torch.int32, device=dev), "tile_lat_info": torch.zeros((B, 4), dtype=torch.int32, device=dev),
        "stats": {"tile_ring": torch.zeros(B, dtype=torch.int64), "tile_far": torch.zeros(B, dtype=torch.int64),
                  "tile_cap_arcs": torch.zeros(B, dtype=torch.int64), "tile_cap_bytes": torch.zeros(B, dtype=torch.int64)},
    }
    if ts.numel() == 0:
        return empty
    S = int(lt_s.numel())

    _phase("tiles: slices", dev)
    # ---- slices ----
    jmax = int(st_j[ts].max()) + 1
    sl_key = (slot[ts] * NW_MAX + st_w[ts]) * jmax + st_j[ts]
    new = torch.ones(ts.numel(), dtype=torch.bool, device=dev)
    new[1:] = sl_key[1:] != sl_key[:-1]
    slice_of_ts = torch.cumsum(new.to(torch.int64), 0) - 1
    NSL = int(slice_of_ts[-1]) + 1
    sl_first = ts[new]
    sl_nst = torch.bincount(slice_of_ts, minlength=NSL)
    sl_lat, sl_slot, sl_w = lt_s[sl_first], slot[sl_first], st_w[sl_first]
    sl_level = sl_slot - level_off[sl_lat]
    deg_ts = out_deg[ts]
    sl_dmax = torch.zeros(NSL, **i64).scatter_reduce(0, slice_of_ts, deg_ts, reduce="amax")
    sl_heavy = sl_dmax > TAILMAX
    # n_k[s, k] = states of slice s with more than k arcs, k < 40 (columns 0..7 in the header, 8..39 in the extension)
    NK = 40
    hist = torch.bincount(slice_of_ts * (NK + 1) + torch.clamp(deg_ts, max=NK), minlength=NSL * (NK + 1)).view(NSL, NK + 1)
    n_k = sl_nst.unsqueeze(1) - torch.cumsum(hist, 1)[:, :NK]
    sl_arc0 = out_ptr[sl_first]
    sl_arcs = out_ptr[sl_first + sl_nst] - sl_arc0
    slice_of_state = torch.full((S,), -1, **i64)
    slice_of_state[ts] = slice_of_ts
    # degrees must descend inside a regular slice (column k = a prefix of the lanes)
    lane_ts = ts - sl_first[slice_of_ts]

    _phase("tiles: ring", dev)
    # ---- ring: 32 slots per slice, in slice order ----
    lat_slices = torch.bincount(sl_lat, minlength=B)
    lat_first_slice = _excl_cumsum(lat_slices)[:-1]
    sl_ord = torch.arange(NSL, device=dev) - lat_first_slice[sl_lat]
    n_slots_tot = int(level_off[-1])
    slot_first_ord = torch.full((n_slots_tot,), 2**40, **i64).scatter_reduce(0, sl_slot, sl_ord, reduce="amin")
    slot_slices = torch.bincount(sl_slot, minlength=n_slots_tot)
    lvl_slices_max = torch.zeros(B, **i64).scatter_reduce(0, sl_lat, slot_slices[sl_slot], reduce="amax")
    # arcs of tile lattices
    at = torch.nonzero(tile_lat[lt_s[src_out]]).squeeze(1)
    a_src, a_dst = src_out[at], dst_out[at]
    a_lat = lt_s[a_src]
    last_slot = level_off[a_lat] + n_levels[a_lat] - 1
    a_last = slot[a_dst] == last_slot  # destination in the last level: a constant, never the ring
    od = sl_ord[slice_of_state[a_dst]]
    need = od - slot_first_ord[slot[a_src]] + 1  # slices the ring must span for this arc
    need = torch.where(a_last, torch.zeros_like(need), need)
    cap_slots = ring_cap_slots(tile_nw, n_levels.to(torch.int64))  # ring + constant slot + far table, per lattice
    ring_cap = torch.div(cap_slots, 32, rounding_mode="floor") - 1  # slices
    if bool((lvl_slices_max[tile_lat] > ring_cap[tile_lat]).any()):
        raise ValueError("a level is wider than the largest DP ring; pack with tiles=False or fewer warps (NFST_TILE_WARPS)")
    # Ring size per lattice.  A destination that has left the ring when its source is processed is kept in a
    # small FAR TABLE behind the ring instead (slots W+1, W+2, ...: written once, never recycled), so the arc
    # reads it like any other slot.  n slices of ring cost 32 n slots plus one table slot per far destination:
    # take the n that minimises the sum (far arcs as a bound on far destinations), from a histogram of `need`.
    nmax = int(need.max()) + 1 if need.numel() else 1
    hist_need = torch.bincount(a_lat * (nmax + 1) + need, minlength=B * (nmax + 1)).view(B, nmax + 1)
    far_arcs_if = hist_need.sum(1, keepdim=True) - torch.cumsum(hist_need, 1)  # [B, n]: arcs with need > n
    cand = torch.arange(nmax + 1, device=dev).unsqueeze(0)
    cost = 32 * cand + far_arcs_if
    floor_n = torch.clamp(lvl_slices_max, min=1).unsqueeze(1)
    cost = torch.where((cand >= floor_n) & (cand <= ring_cap.unsqueeze(1)), cost, torch.full_like(cost, 2**40))
    ring_slices = torch.minimum(torch.maximum(torch.argmin(cost, 1), floor_n.squeeze(1)), torch.clamp(ring_cap, min=1))
    if FORCE_RING_SLICES:
        ring_slices = torch.maximum(torch.clamp(ring_slices, max=FORCE_RING_SLICES), floor_n.squeeze(1))
    best = torch.gather(cost, 1, ring_slices.unsqueeze(1)).squeeze(1)
    if bool((tile_lat & (best + 1 > cap_slots)).any()):
        raise ValueError("DP ring and far table of a lattice exceed shared memory (its arcs reach too far for its width); "
                         "pack with tiles=False or fewer warps (NFST_TILE_WARPS)")
    W = ring_slices * 32  # [B]
    Wa = W[a_lat]
    resident = (~a_last) & (need <= ring_slices[a_lat])
    far = (~a_last) & (~resident)
    # far table: one slot per far destination, numbered per lattice in state order
    far_slot = torch.zeros(S, **i64)  # 0 = none
    n_far = torch.zeros(B, **i64)
    sl_flags = torch.zeros(NSL, **i64)
    if bool(far.any()):
        fd = torch.unique(a_dst[far])  # sorted global state ids
        fd_lat = lt_s[fd]
        n_far = torch.bincount(fd_lat, minlength=B)
        rank = torch.arange(fd.numel(), device=dev) - _excl_cumsum(n_far)[:-1][fd_lat]
        far_slot[fd] = W[fd_lat] + 1 + rank
        f_in = torch.zeros(NSL, dtype=torch.bool, device=dev)
        f_in[slice_of_state[fd]] = True
        sl_flags = f_in.to(torch.int64) * FLAG_FAR_IN
    ring_total = W + 1 + n_far  # ring, the constant slot, the far table
    if bool((ring_total[tile_lat] > 65535).any()):
        raise ValueError("DP ring + far table exceed 16-bit slot numbers; lower NFST_TILE_RING_MAX")
    code = torch.where(resident, (32 * od + (a_dst - sl_first[slice_of_state[a_dst]])) % Wa,
                       torch.where(far, far_slot[a_dst], Wa))
    sl_vslot = (32 * sl_ord) % W[sl_lat]

    _phase("tiles: segments", dev)
    # ---- segments (one header each): regular slices, heavy slices cut into pieces of T arcs ----
    T = T_lat[sl_lat]  # per slice
    n_seg = torch.where(sl_heavy, torch.clamp((sl_arcs + T - 1) // T, min=1), torch.ones_like(sl_arcs))
    seg_start = _excl_cumsum(n_seg)
    NSEG = int(seg_start[-1])
    seg_slice = torch.repeat_interleave(torch.arange(NSL, device=dev), n_seg, output_size=NSEG)
    seg_piece = torch.arange(NSEG, device=dev) - seg_start[seg_slice]
    seg_heavy = sl_heavy[seg_slice]
    T = T[seg_slice]  # per segment from here on
    seg_arcs = torch.where(seg_heavy, torch.minimum(sl_arcs[seg_slice] - seg_piece * T, T), sl_arcs[seg_slice])
    seg_arc0 = sl_arc0[seg_slice] + seg_piece * T
    seg_flags = sl_flags[seg_slice] + seg_heavy.to(torch.int64) * (
        FLAG_HEAVY + (seg_piece == 0).to(torch.int64) * FLAG_HEAVY_FIRST
        + (seg_piece == n_seg[seg_slice] - 1).to(torch.int64) * FLAG_HEAVY_LAST)
    seg_ext = (~seg_heavy) & (sl_dmax[seg_slice] > KU)  # 32-byte block: n_8 .. n_39
    seg_farb = (~seg_heavy) & ((sl_flags[seg_slice] & FLAG_FAR_IN) > 0)  # 64-byte block: far-table slot per lane
    seg_ext_units = seg_ext.to(torch.int64) + 2 * seg_farb.to(torch.int64)  # in 32-byte units

    _phase("tiles: tiles", dev)
    # ---- tiles: consecutive segments of one (level, warp); heavy pieces stand alone ----
    g = (sl_slot * NW_MAX + sl_w)[seg_slice]
    g_new = torch.ones(NSEG, dtype=torch.bool, device=dev)
    g_new[1:] = g[1:] != g[:-1]
    g_id = torch.cumsum(g_new.to(torch.int64), 0) - 1
    g_first = torch.nonzero(g_new).squeeze(1)
    idx_in_g = torch.arange(NSEG, device=dev) - g_first[g_id]
    # greedy, in segment order: a segment joins the open tile unless that would exceed T arcs or TILE_SLICES
    # segments (so a tile holds at most max(T, its largest segment) arcs); heavy pieces stand alone.  One
    # vectorised step per position in the group (a group = what one warp does in one level: a handful of segments).
    NG = int(g_first.numel())
    acc = torch.zeros(NG, **i64)
    cnt = torch.zeros(NG, **i64)
    prev_heavy = torch.zeros(NG, dtype=torch.bool, device=dev)
    tile_new = torch.zeros(NSEG, dtype=torch.bool, device=dev)
    order_g = torch.argsort(idx_in_g, stable=True)  # segments by position in their group
    pos_counts = torch.bincount(idx_in_g)
    pos_off = _excl_cumsum(pos_counts)
    for j in range(int(pos_counts.numel())):
        sel = order_g[int(pos_off[j]): int(pos_off[j + 1])]  # the j-th segment of every group that has one
        gi = g_id[sel]
        a, hv, Tj = seg_arcs[sel], seg_heavy[sel], T[sel]
        start = (cnt[gi] == 0) | hv | prev_heavy[gi] | (acc[gi] + a > Tj) | (cnt[gi] >= TILE_SLICES)
        tile_new[sel] = start
        acc[gi] = torch.where(start, a, acc[gi] + a)
        cnt[gi] = torch.where(start, torch.ones_like(a), cnt[gi] + 1)
        prev_heavy[gi] = hv
    tile_of_seg = torch.cumsum(tile_new.to(torch.int64), 0) - 1
    NT = int(tile_of_seg[-1]) + 1
    t_first_seg = torch.nonzero(tile_new).squeeze(1)
    t_nseg = torch.bincount(tile_of_seg, minlength=NT)
    t_arc0 = seg_arc0[t_first_seg]
    t_arcs = torch.zeros(NT, **i64).index_add_(0, tile_of_seg, seg_arcs)
    t_next = torch.zeros(NT, **i64).index_add_(0, tile_of_seg, seg_ext_units)
    t_slice0 = seg_slice[t_first_seg]
    t_lat, t_w, t_level = sl_lat[t_slice0], sl_w[t_slice0], sl_level[t_slice0]
    t_dst_off = 16 + 16 * t_nseg + 32 * t_next
    t_bytes = (t_dst_off + 2 * t_arcs + 64 + 15) // 16 * 16
    if int(t_arcs.max()) >= 65536 or int(t_bytes.max()) >= 65536 or int(t_level.max()) >= 65536 or int(t_next.max()) >= 256:
        raise ValueError("a tile is too large for its 16-bit size fields")
    t_off = _excl_cumsum(t_bytes)
    n_bytes = int(t_off[-1])
    t_off = t_off[:-1]

    _phase("tiles: the byte stream, written as 16-bit words", dev)
    # ---- the byte stream, written as 16-bit words ----
    stream = torch.zeros(n_bytes // 2, dtype=torch.int32, device=dev)  # values 0..65535, narrowed at the end

    def put16(byte_off, val):
        stream[torch.div(byte_off, 2, rounding_mode="floor")] = (val & 0xFFFF).to(torch.int32)

    def put32(byte_off, val):
        put16(byte_off, val)
        put16(byte_off + 2, val >> 16)

    # ring-slot region (+ the 64-byte pad and the rounding) pre-filled with the lattice's constant slot W
    fill_n = torch.div(t_bytes - t_dst_off, 2, rounding_mode="floor")
    fill_tile = torch.repeat_interleave(torch.arange(NT, device=dev), fill_n)
    fill_pos = torch.arange(int(fill_n.sum()), device=dev) - _excl_cumsum(fill_n)[:-1][fill_tile]
    stream[torch.div(t_off + t_dst_off, 2, rounding_mode="floor")[fill_tile] + fill_pos] = W[t_lat][fill_tile].to(torch.int32)
    # tile header: state of the first slice (lattice-relative), its ring slot, offset of the ring-slot region
    put32(t_off, sl_first[t_slice0] - state_off[t_lat])
    put32(t_off + 4, t_arc0 - out_ptr[state_off[:-1]][t_lat])
    put16(t_off + 8, sl_vslot[t_slice0])
    put16(t_off + 10, t_dst_off)
    put16(t_off + 12, t_level)
    put16(t_off + 14, t_nseg)
    # segment headers
    seg_tile = tile_of_seg
    seg_idx = torch.arange(NSEG, device=dev) - t_first_seg[seg_tile]
    hdr = t_off[seg_tile] + 16 + 16 * seg_idx
    nk_seg = n_k[seg_slice]
    for q in range(4):
        lo, hi = nk_seg[:, 2 * q], nk_seg[:, 2 * q + 1]
        put16(hdr + 2 * q, torch.where(seg_heavy, torch.zeros_like(lo), lo | (hi << 8)))
    # heavy pieces: bytes 0-3 = arcs of the whole state, bytes 4-7 = arcs before this piece
    hv = torch.nonzero(seg_heavy).squeeze(1)
    if hv.numel():
        put32(hdr[hv], sl_arcs[seg_slice[hv]])
        put32(hdr[hv] + 4, seg_piece[hv] * T[hv])
    # arc offset in the tile; heavy pieces (always offset 0): the state's far-table slot instead
    put16(hdr + 8, torch.where(seg_heavy, far_slot[sl_first[seg_slice]], seg_arc0 - t_arc0[seg_tile]))
    put16(hdr + 10, sl_nst[seg_slice] | (torch.clamp(sl_dmax[seg_slice], max=255) << 8))
    ext_rank = torch.cumsum(seg_ext_units, 0) - seg_ext_units  # 32-byte units before this segment's region
    ext_idx = ext_rank - ext_rank[t_first_seg][seg_tile]
    ext_off = torch.where(seg_ext_units > 0, 16 + 16 * t_nseg[seg_tile] + 32 * ext_idx, torch.zeros_like(ext_idx))
    put16(hdr + 12, torch.where(seg_heavy, seg_arcs, ext_off))  # heavy pieces: arcs of the piece
    put16(hdr + 14, seg_flags)
    ex = torch.nonzero(seg_ext).squeeze(1)
    if ex.numel():
        base = t_off[seg_tile[ex]] + ext_off[ex]
        for q in range(16):
            put16(base + 2 * q, nk_seg[ex, KU + 2 * q] | (nk_seg[ex, KU + 2 * q + 1] << 8))
    fb = torch.nonzero(seg_farb).squeeze(1)
    if fb.numel():  # far-table slots of the segment's 32 lanes (0 = none), behind the n_k block if there is one
        base = t_off[seg_tile[fb]] + ext_off[fb] + 32 * seg_ext[fb].to(torch.int64)
        st0 = sl_first[seg_slice[fb]]
        nst = sl_nst[seg_slice[fb]]
        for ln in range(32):
            put16(base + 2 * ln, torch.where(ln < nst, far_slot[torch.clamp(st0 + ln, max=S - 1)], torch.zeros_like(st0)))
    # ring slots of the arcs
    a_tile = tile_of_seg[seg_start[slice_of_state[a_src]]
                         + torch.where(sl_heavy[slice_of_state[a_src]],
                                       torch.div(at - sl_arc0[slice_of_state[a_src]], T_lat[a_lat], rounding_mode="floor"),
                                       torch.zeros_like(at))]
    put16(t_off[a_tile] + t_dst_off[a_tile] + 2 * (at - t_arc0[a_tile]), code)
    tile_stream = stream.to(torch.int16).view(torch.uint8)  # little-endian 16-bit words

    _phase("tiles: tile table", dev)
    # ---- tile table: per (lattice, warp) in level order ----
    lmax = int(t_level.max()) + 1
    t_order = torch.argsort((t_lat * NW_MAX + t_w) * lmax + t_level, stable=True)
    nw_eff = torch.where(tile_lat, tile_nw, torch.zeros_like(tile_nw))
    lw_base = _excl_cumsum(nw_eff)
    lw_id = lw_base[:-1][t_lat] + t_w
    lw_off = _excl_cumsum(torch.bincount(lw_id, minlength=int(lw_base[-1])))
    tab = torch.stack([t_arc0, torch.div(t_off, 16, rounding_mode="floor"), t_arcs | (t_nseg << 16) | (t_next << 24),
                       t_level | (torch.div(t_bytes, 16, rounding_mode="floor") << 16)], dim=1)[t_order]
    tab = torch.where(tab >= 2**31, tab - 2**32, tab).to(torch.int32).contiguous()
    info = torch.stack([lw_base[:-1], W, ring_total, out_ptr[state_off[:-1]]], dim=1) * tile_lat.to(torch.int64).unsqueeze(1)
    info = info.to(torch.int32).contiguous()  # rows of other lattices stay zero

    cap_arcs = torch.zeros(B, **i64).scatter_reduce(0, t_lat, t_arcs, reduce="amax")
    cap_bytes = torch.zeros(B, **i64).scatter_reduce(0, t_lat, t_bytes, reduce="amax")
    # sanity of the column-major layout: inside a regular slice the degrees descend
    reg = ~sl_heavy[slice_of_ts]
    if reg.numel() > 1:
        same = (slice_of_ts[1:] == slice_of_ts[:-1]) & reg[1:]
        if bool((deg_ts[1:][same] > deg_ts[:-1][same]).any()):
            raise AssertionError("tile slices: out-degrees must descend inside a slice")
    del lane_ts
    return {
        "tile_stream": tile_stream, "tile_tab": tab, "tile_lw_off": lw_off.to(torch.int32).contiguous(), "tile_lat_info": info,
        "stats": {"tile_ring": (ring_total * tile_lat).cpu(), "tile_far": (n_far * tile_lat).cpu(), "tile_cap_arcs": cap_arcs.cpu(),
                  "tile_cap_bytes": cap_bytes.cpu()},
    }
