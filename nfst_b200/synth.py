"""Seeded synthetic lattices with the shapes of the BASELINE.json configs.

The reference ships no data, so every workload is generated: edit lattices with the
topology of the transliteration machine (``src/fsm/tr.py:321-390``), char x tag grids for
SNIPS (``src/fsm/snips.py``), cipher trellises, and layered random DAGs.  Generators return
arc lists (``ArcBatch``) that ``pack_arcs`` turns into the packed layout; nothing here is
on the timed path.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np
import torch

from .pack import PackedLattices, pack_arcs

BOS, EOS, PAD = 1, 2, 3


@dataclasses.dataclass
class ArcBatch:
    arc_lattice: torch.Tensor  # int64 [A]
    src: torch.Tensor  # int64 [A] local state ids
    dst: torch.Tensor
    label: torch.Tensor
    scores: torch.Tensor  # float32 [A]
    n_states: torch.Tensor  # int64 [B]
    vocab: int

    def pack(self, **kw) -> "tuple[PackedLattices, torch.Tensor]":
        """(packed lattices, arc scores in canonical order)."""
        p = pack_arcs(self.arc_lattice, self.src, self.dst, self.label, self.n_states, self.vocab, **kw)
        return p, self.scores[p.arc_origin].contiguous()

    def select(self, lattice_ids) -> "ArcBatch":
        """The sub-batch of the given lattices (ascending ids), renumbered 0..n-1; arc order is kept."""
        ids = torch.as_tensor(lattice_ids, dtype=torch.int64, device=self.src.device)
        remap = torch.full((int(self.n_states.numel()),), -1, dtype=torch.int64, device=self.src.device)
        remap[ids] = torch.arange(ids.numel(), device=self.src.device)
        new_lat = remap[self.arc_lattice]
        keep = new_lat >= 0
        return ArcBatch(new_lat[keep], self.src[keep], self.dst[keep], self.label[keep], self.scores[keep], self.n_states[ids], self.vocab)

    def to(self, device) -> "ArcBatch":
        return ArcBatch(*(getattr(self, f.name).to(device) if isinstance(getattr(self, f.name), torch.Tensor)
                          else getattr(self, f.name) for f in dataclasses.fields(self)))


def _from_lists(lat, src, dst, lab, sc, ns, vocab) -> ArcBatch:
    cat = lambda xs, dt: torch.from_numpy(np.concatenate(xs).astype(dt))  # noqa: E731
    return ArcBatch(cat(lat, np.int64), cat(src, np.int64), cat(dst, np.int64), cat(lab, np.int64),
                    cat(sc, np.float32), torch.tensor(ns, dtype=torch.int64), vocab)


# --------------------------------------------------------------------------------------
# config 1 / 5: transliteration edit lattices ("sub" machine, fsm/tr.py:321-390)
# --------------------------------------------------------------------------------------
def _edit_lattice(rng: np.random.Generator, n: int, m: int, vocab: int):
    """Grid (n+1)x(m+1); per cell delete = 2-arc chain, insert = 2-arc chain, substitute =
    5-arc chain (mark strings of the sub machine), plus bos arc, eos arc and sink."""
    grid = lambda i, j: 1 + i * (m + 1) + j  # noqa: E731  state 0 = start
    n_grid = (n + 1) * (m + 1)
    x = rng.integers(8, 8 + 60, size=n)  # input symbol marks
    y = rng.integers(70, 70 + 60, size=m)  # output symbol marks
    nxt = 1 + n_grid
    src, dst, lab = [0], [grid(0, 0)], [BOS]
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(m + 1), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    # delete x_i : (i,j) -x_i-> t -DEL-> (i+1,j)
    sel = ii < n
    k = int(sel.sum())
    t = nxt + np.arange(k); nxt += k
    src += [grid(ii[sel], jj[sel]), t]; dst += [t, grid(ii[sel] + 1, jj[sel])]
    lab += [x[ii[sel]], np.full(k, 4)]
    # insert y_j : (i,j) -INS-> t -y_j-> (i,j+1)
    sel = jj < m
    k = int(sel.sum())
    t = nxt + np.arange(k); nxt += k
    src += [grid(ii[sel], jj[sel]), t]; dst += [t, grid(ii[sel], jj[sel] + 1)]
    lab += [np.full(k, 5), y[jj[sel]]]
    # substitute x_i:y_j : 5-arc chain (i,j) -SUB-> t1 -x_i-> t2 -IN-> t3 -y_j-> t4 -OUT-> (i+1,j+1)
    sel = (ii < n) & (jj < m)
    k = int(sel.sum())
    t1 = nxt + np.arange(k); t2 = t1 + k; t3 = t2 + k; t4 = t3 + k; nxt += 4 * k
    g0, g1 = grid(ii[sel], jj[sel]), grid(ii[sel] + 1, jj[sel] + 1)
    src += [g0, t1, t2, t3, t4]; dst += [t1, t2, t3, t4, g1]
    lab += [np.full(k, 6), x[ii[sel]], np.full(k, 7), y[jj[sel]], np.full(k, 135)]
    sink = nxt
    src.append(grid(n, m)); dst.append(sink); lab.append(EOS)
    flat = lambda xs: np.concatenate([np.atleast_1d(np.asarray(v)) for v in xs])  # noqa: E731
    return flat(src), flat(dst), flat(lab), sink + 1


def transliteration_batch(B: int = 32, seed: int = 0, vocab: int = 256, lo: int = 4, hi: int = 12,
                          integer_scores: bool = False) -> ArcBatch:
    """Config 1 (B=32) and config 5 (B=4096): |x|,|y| ~ U{lo..hi}; theta ~ N(0,1) per label,
    or integer scores in {-2,-1,0} (forces exact Viterbi ties)."""
    rng = np.random.default_rng(seed)
    theta = rng.normal(size=vocab) if not integer_scores else rng.integers(-2, 1, size=vocab).astype(np.float64)
    lat, src, dst, lab, sc, ns = [], [], [], [], [], []
    for b in range(B):
        n, m = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
        s, d, l, S = _edit_lattice(rng, n, m, vocab)
        lat.append(np.full(len(s), b)); src.append(s); dst.append(d); lab.append(l); ns.append(S)
        sc.append(theta[l])
    return _from_lists(lat, src, dst, lab, sc, ns, vocab)


# --------------------------------------------------------------------------------------
# config 2: SNIPS slot tagging -- char x tag grids with an intent fan-out
# --------------------------------------------------------------------------------------
def snips_batch(B: int = 256, seed: int = 1, vocab: int = 256) -> ArcBatch:
    """|x| ~ U{20..60} characters, |y| ~ U{4..12} tag positions; per cell one char arc and
    B/I/O tag arcs; 7-way intent fan-out at the end (fsm/snips.py:20-28)."""
    rng = np.random.default_rng(seed)
    theta = rng.normal(size=vocab)
    lat, src, dst, lab, sc, ns = [], [], [], [], [], []
    for b in range(B):
        n, m = int(rng.integers(20, 61)), int(rng.integers(4, 13))
        chars = rng.integers(8, 8 + 100, size=n)
        node = lambda i, t: 1 + i * m + t  # noqa: E731  (i chars consumed, tag slot t)
        n_node = (n + 1) * m
        s_, d_, l_ = [np.array([0])], [np.array([node(0, 0)])], [np.array([BOS])]
        ii, tt = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
        ii, tt = ii.ravel(), tt.ravel()
        k = len(ii)
        mid = 1 + n_node + np.arange(k)  # after the char arc, before the tag arc
        s_.append(node(ii, tt)); d_.append(mid); l_.append(chars[ii])
        # tag arcs: I (stay in slot) ...
        s_.append(mid); d_.append(node(ii + 1, tt)); l_.append(np.full(k, 120))
        # ... O (stay, outside tag) with a distinct label
        s_.append(mid); d_.append(node(ii + 1, tt)); l_.append(np.full(k, 121))
        # ... B (advance to the next slot)
        adv = tt < m - 1
        s_.append(mid[adv]); d_.append(node(ii[adv] + 1, tt[adv] + 1)); l_.append(122 + tt[adv])
        nxt = 1 + n_node + k
        intents = nxt + np.arange(7); sink = nxt + 7
        for t in range(m):  # every tag slot may end the utterance
            s_.append(np.full(7, node(n, t))); d_.append(intents); l_.append(140 + np.arange(7))
        s_.append(intents); d_.append(np.full(7, sink)); l_.append(np.full(7, EOS))
        s, d, l = np.concatenate(s_), np.concatenate(d_), np.concatenate(l_)
        lat.append(np.full(len(s), b)); src.append(s); dst.append(d); lab.append(l); ns.append(sink + 1)
        sc.append(theta[l])
    return _from_lists(lat, src, dst, lab, sc, ns, vocab)


# --------------------------------------------------------------------------------------
# config 3: substitution cipher trellises
# --------------------------------------------------------------------------------------
def cipher_batch(B: int = 64, T: int = 1000, bigram: bool = True, seed: int = 2, alphabet: int = 26,
                 device="cpu") -> ArcBatch:
    """Ciphertext of length T over `alphabet` symbols.  bigram=False: T+1 chain states
    with `alphabet` parallel arcs per position (depth T, width 1).  bigram=True: `alphabet`
    states per position, alphabet^2 arcs per position.  Arc score = log P(c_t | p) +
    log P(p | p') from seeded random stochastic matrices."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    K = alphabet
    chan = torch.log_softmax(torch.randn(K, K, generator=g), dim=1)  # log P(c | p)
    lm2 = torch.log_softmax(torch.randn(K, K, generator=g), dim=1)  # log P(p | p')
    lm1 = torch.log_softmax(torch.randn(K, generator=g), dim=0)
    cipher = torch.randint(0, K, (B, T), generator=g)
    vocab = 8 + K
    dev = torch.device(device)
    chan, lm2, lm1, cipher = chan.to(dev), lm2.to(dev), lm1.to(dev), cipher.to(dev)
    ar = lambda n: torch.arange(n, device=dev)  # noqa: E731
    if not bigram:
        # states: 0 start(bos) -> 1 .. T+1 chain -> sink T+2
        t = ar(T).repeat_interleave(K)
        p = ar(K).repeat(T)
        src1 = (1 + t).repeat(B)
        dst1 = (2 + t).repeat(B)
        lab1 = (8 + p).repeat(B)
        c = cipher[:, :, None].expand(B, T, K).reshape(-1)
        sc1 = chan[p.repeat(B), c] + lm1[p.repeat(B)]
        lat1 = ar(B).repeat_interleave(T * K)
        S = T + 3
        src = torch.cat([torch.zeros(B, dtype=torch.int64, device=dev), src1, torch.full((B,), T + 1, device=dev)])
        dst = torch.cat([torch.ones(B, dtype=torch.int64, device=dev), dst1, torch.full((B,), T + 2, device=dev)])
        lab = torch.cat([torch.full((B,), BOS, device=dev), lab1, torch.full((B,), EOS, device=dev)])
        sc = torch.cat([torch.zeros(B, device=dev), sc1, torch.zeros(B, device=dev)])
        lat = torch.cat([ar(B), lat1, ar(B)])
    else:
        # states: 0 start, 1 hub (after bos), 2 + t*K + p, pre-final 2+T*K, sink 3+T*K
        node = lambda t, p: 2 + t * K + p  # noqa: E731
        S = T * K + 4
        pieces = []
        # bos
        pieces.append((ar(B), torch.zeros(B, dtype=torch.int64, device=dev), torch.ones(B, dtype=torch.int64, device=dev),
                       torch.full((B,), BOS, device=dev), torch.zeros(B, device=dev)))
        # hub -> (0, p)
        p0 = ar(K).repeat(B)
        latp = ar(B).repeat_interleave(K)
        pieces.append((latp, torch.ones(B * K, dtype=torch.int64, device=dev), node(0, p0), 8 + p0,
                       chan[p0, cipher[latp, 0]] + lm1[p0]))
        # (t-1, p') -> (t, p)
        tt = (1 + ar(T - 1)).repeat_interleave(K * K)
        pp_prev = ar(K).repeat_interleave(K).repeat(T - 1)
        pp = ar(K).repeat(K * (T - 1))
        n1 = tt.numel()
        latt = ar(B).repeat_interleave(n1)
        tt_b, pprev_b, pp_b = tt.repeat(B), pp_prev.repeat(B), pp.repeat(B)
        pieces.append((latt, node(tt_b - 1, pprev_b), node(tt_b, pp_b), 8 + pp_b,
                       chan[pp_b, cipher[latt, tt_b]] + lm2[pprev_b, pp_b]))
        # (T-1, p) -> pre-final (label = a per-letter closing mark keeps the FSA deterministic)
        pieces.append((latp, node(T - 1, p0), torch.full((B * K,), 2 + T * K, device=dev), torch.full((B * K,), 4, device=dev),
                       torch.zeros(B * K, device=dev)))
        # eos
        pieces.append((ar(B), torch.full((B,), 2 + T * K, device=dev), torch.full((B,), 3 + T * K, device=dev),
                       torch.full((B,), EOS, device=dev), torch.zeros(B, device=dev)))
        lat, src, dst, lab, sc = (torch.cat([pc[i] for pc in pieces]) for i in range(5))
    return ArcBatch(lat.to(torch.int64), src.to(torch.int64), dst.to(torch.int64), lab.to(torch.int64),
                    sc.to(torch.float32), torch.full((B,), S, dtype=torch.int64, device=dev), vocab)


# --------------------------------------------------------------------------------------
# config 4: layered random DAGs
# --------------------------------------------------------------------------------------
def random_dag_batch(B: int, arcs_per_lattice: int, levels: int = 64, seed: int = 3, vocab: int = 256,
                     device="cpu") -> ArcBatch:
    """S = A/4 states in `levels` equal-width levels (single source, single sink); every
    non-source state draws 1 + Poisson(3) predecessors uniformly from the previous 1-3
    levels; dead ends are wired to the sink; scores ~ U(-1, 0)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    S = max(arcs_per_lattice // 4, levels)
    width = max((S - 2) // (levels - 2), 1)
    S = width * (levels - 2) + 2
    sink = S - 1
    # state s in 1..S-2 sits at level 1 + (s-1)//width; level 0 = {0}; last level = {sink}
    inner = torch.arange(1, S - 1, device=dev)
    lvl = 1 + (inner - 1) // width
    n_inner = inner.numel()
    deg = 1 + torch.poisson(torch.full((B, n_inner), 3.0, device=dev), generator=g).to(torch.int64)
    deg[:, lvl == 1] = 1  # only the source precedes level 1
    deg = deg.reshape(-1)
    dst_state = inner.repeat(B)
    dst_lvl = lvl.repeat(B)
    lat_state = torch.arange(B, device=dev).repeat_interleave(n_inner)
    dst = dst_state.repeat_interleave(deg)
    dl = dst_lvl.repeat_interleave(deg)
    lat = lat_state.repeat_interleave(deg)
    n = dst.numel()
    back = torch.randint(1, 4, (n,), generator=g, device=dev)  # 1..3 levels back
    sl = torch.clamp(dl - back, min=0)
    pos = torch.randint(0, width, (n,), generator=g, device=dev)
    src = torch.where(sl == 0, torch.zeros_like(pos), 1 + (sl - 1) * width + pos)
    # sink: fed by the last inner level and by every dead end
    gid = lat * S + src
    has_out = torch.zeros(B * S, dtype=torch.bool, device=dev)
    has_out[gid] = True
    has_out = has_out.view(B, S)
    has_out[:, sink] = True
    dead_b, dead_s = torch.nonzero(~has_out, as_tuple=True)
    lat = torch.cat([lat, dead_b])
    src = torch.cat([src, dead_s])
    dst = torch.cat([dst, torch.full_like(dead_s, sink)])
    A = lat.numel()
    label = torch.randint(8, vocab, (A,), generator=g, device=dev)
    scores = -torch.rand(A, generator=g, device=dev)
    return ArcBatch(lat, src, dst, label, scores, torch.full((B,), S, dtype=torch.int64, device=dev), vocab)
