"""numpy replay of the tile-stream layout (nfst_b200/tiles.py) with the SAME addressing the kernels of
nfst_tiles.cu use: per (lattice, warp) tile lists from ``tile_tab``, tile / slice headers and 16-bit ring slots
from ``tile_stream``, a DP ring of ``W`` slots (+ the constant slot ``W``), far arcs through ``dst_out``.
Test infrastructure only."""
import numpy as np

from nfst_b200 import tiles as T


def _np(t):
    return t.cpu().numpy()


class Tile:
    def __init__(self, p, row, b):
        tab = _np(p.tile_tab)[row].astype(np.int64) & 0xFFFFFFFF
        self.arc0 = int(tab[0])
        self.off = int(tab[1]) * 16
        self.n_arcs = int(tab[2] & 0xFFFF)
        self.n_seg = int((tab[2] >> 16) & 0xFF)
        self.n_ext = int((tab[2] >> 24) & 0xFF)
        self.level = int(tab[3] & 0xFFFF)
        self.n_bytes = int(tab[3] >> 16) * 16
        raw = _np(p.tile_stream)[self.off:self.off + self.n_bytes]
        self.raw = raw
        h, h16 = raw[:16].view(np.int32), raw[:16].view(np.uint16)
        s_base = int(_np(p.state_off)[b])
        self.state0 = int(h[0]) + s_base
        assert int(_np(p.tile_lat_info)[b, 3]) == int(_np(p.out_ptr)[s_base])
        assert int(h[1]) + int(_np(p.tile_lat_info)[b, 3]) == self.arc0
        self.vslot0, self.dst_off = int(h16[4]), int(h16[5])
        assert int(h16[6]) == self.level and int(h16[7]) == self.n_seg
        assert self.dst_off == 16 + 16 * self.n_seg + 32 * self.n_ext
        self.codes = raw[self.dst_off:].view(np.uint16)
        self.segs = []
        state, vslot = self.state0, self.vslot0
        W = int(_np(p.tile_lat_info)[b, 1])
        for i in range(self.n_seg):
            s = raw[16 + 16 * i: 32 + 16 * i]
            u16 = s.view(np.uint16)
            seg = {"arc_rel": int(u16[4]), "n_states": int(s[10]), "dmax": int(s[11]), "ext": int(u16[6]), "flags": int(u16[7]),
                   "state0": state, "vslot": vslot}
            if seg["flags"] & T.FLAG_HEAVY:
                seg["state_arcs"], seg["before"] = int(s[:4].view(np.uint32)[0]), int(s[4:8].view(np.uint32)[0])
                seg["n_k"] = None
                seg["far_slots"] = [seg["arc_rel"]]  # heavy pieces: the arc-offset field holds the state's far-table slot
                seg["arc_rel"] = 0
            else:
                nk = list(s[:8])
                pos = seg["ext"]
                if seg["dmax"] > T.KU:
                    nk += list(raw[pos: pos + 32])
                    pos += 32
                seg["n_k"] = [int(x) for x in nk]
                seg["far_slots"] = [0] * 32
                if seg["flags"] & T.FLAG_FAR_IN:
                    seg["far_slots"] = [int(x) for x in raw[pos: pos + 64].view(np.uint16)]
            self.segs.append(seg)
            # the next segment of the tile: the next 32 states / ring slots (heavy pieces stand alone in their tile)
            state += 32
            vslot = (vslot + 32) % W


def tiles_of(p, b, nw):
    """{warp: [Tile...]} of lattice b in level order."""
    info = _np(p.tile_lat_info)[b]
    lw = _np(p.tile_lw_off)
    return {w: [Tile(p, r, b) for r in range(lw[info[0] + w], lw[info[0] + w + 1])] for w in range(nw)}


def seg_arcs(tile, seg):
    """[(lane, k, arc id, ring code)] of one segment in the order the kernels address it."""
    out = []
    pos = seg["arc_rel"]
    if seg["flags"] & T.FLAG_HEAVY:
        n = seg["ext"]  # heavy pieces: the extension field holds the arcs of the piece
        assert n == tile.n_arcs and n <= seg["state_arcs"] - seg["before"]
        for i in range(n):
            out.append((0, seg["before"] + i, tile.arc0 + pos + i, int(tile.codes[pos + i])))
        return out
    for k, n in enumerate(seg["n_k"]):
        if k >= seg["dmax"]:
            assert n == 0
            continue
        for lane in range(n):
            out.append((lane, k, tile.arc0 + pos + lane, int(tile.codes[pos + lane])))
        pos += n
    return out


def replay(p, w, b, nw):
    """(beta, cond, post, delta, backptr, visits) of lattice b computed by walking the tile stream like the kernels do."""
    info = _np(p.tile_lat_info)[b]
    W, total = int(info[1]), int(info[2])
    dst = _np(p.dst_out).astype(np.int64)
    S, A = p.n_states, p.n_arcs
    tl = tiles_of(p, b, nw)
    L = int(_np(p.n_levels)[b])
    beta = np.full(S, np.nan)
    delta = np.full(S, np.nan, dtype=np.float32)
    bp = np.full(S, -2, dtype=np.int64)
    cond = np.zeros(A)
    seen = np.zeros(A, dtype=np.int64)
    ring = np.full(total, np.nan)
    ringd = np.full(total, np.nan, dtype=np.float32)
    ring[W], ringd[W] = 0.0, 0.0  # the constant slot: the last level

    for l in range(L - 1, -1, -1):
        writes = []
        for wi in range(nw):
            hv = None
            for t in tl[wi]:
                if t.level != l:
                    continue
                for seg in t.segs:
                    arcs = seg_arcs(t, seg)
                    for (_, _, a, _) in arcs:
                        seen[a] += 1
                    if seg["flags"] & T.FLAG_HEAVY:
                        if seg["flags"] & T.FLAG_HEAVY_FIRST:
                            hv = []
                        hv += [(a, c) for (_, _, a, c) in arcs]
                        if not seg["flags"] & T.FLAG_HEAVY_LAST:
                            continue
                        rows = {0: hv}
                        hv = None
                    else:
                        rows = {i: [] for i in range(seg["n_states"])}
                        for (lane, k, a, c) in arcs:
                            assert len(rows[lane]) == k
                            rows[lane].append((a, c))
                    for lane, ac in rows.items():
                        s = seg["state0"] + lane
                        slots = [seg["vslot"] + lane] + ([seg["far_slots"][lane]] if seg["far_slots"][lane] else [])
                        if not ac:
                            writes.append((s, slots, 0.0, np.float32(0.0), -1))
                            continue
                        ids = [a for a, _ in ac]
                        assert all(c < total for _, c in ac)
                        tt = np.array([w[a].astype(np.float64) + ring[c] for a, c in ac])
                        assert not np.isnan(tt).any(), "an arc read a ring slot its destination does not own (any more)"
                        m = tt.max()
                        bs = m + np.log(np.exp(tt - m).sum()) if np.isfinite(m) else -np.inf
                        cond[ids] = np.exp(tt - bs) if np.isfinite(bs) else 0.0
                        cc = np.array([np.float32(w[a]) + ringd[c] for a, c in ac], dtype=np.float32)
                        j = int(np.argmax(cc))
                        writes.append((s, slots, bs, cc[j], ids[j]))
        for (s, slots, bv, dv, arg) in writes:  # the level barrier: values become visible to the next level
            beta[s], delta[s], bp[s] = bv, dv, arg
            assert slots[0] < W and all(W < x < total for x in slots[1:])
            for x in slots:
                ring[x], ringd[x] = bv, dv
    # every ring read must have hit the slot of its true destination: recompute beta from dst_out
    # flow: start level first; ring slots are zeroed when their owner is consumed
    gring = np.zeros(total)
    post = np.zeros(A)
    start = int(_np(p.start_state)[b])
    first_seg = tl[0][0].segs[0] if tl[0] else None
    assert first_seg is not None and first_seg["state0"] == start and first_seg["vslot"] == 0
    gring[0] = 1.0
    for l in range(L):
        pushes = []
        for wi in range(nw):
            for t in tl[wi]:
                if t.level != l:
                    continue
                for seg in t.segs:
                    heavy = bool(seg["flags"] & T.FLAG_HEAVY)
                    n_st = 1 if heavy else seg["n_states"]
                    if not heavy or seg["flags"] & T.FLAG_HEAVY_FIRST:
                        gam = {}
                        for lane in range(n_st):
                            slot, fs = seg["vslot"] + lane, seg["far_slots"][lane]
                            gam[lane] = gring[slot] + (gring[fs] if fs else 0.0)
                            gring[slot] = 0.0
                        if heavy:
                            hv_g = gam[0]
                    for (lane, _, a, c) in seg_arcs(t, seg):
                        pr = (hv_g if heavy else gam[lane]) * cond[a]
                        post[a] = pr
                        pushes.append((c, pr))
        for (c, pr) in pushes:
            gring[c] += pr
    return beta, cond, post, delta, bp, seen
