"""Lattice packing: dense tables / arc lists -> level-sorted CSR (the HBM layout).

Boundary of the hot path.  The reference hands its lattice DP two dense tables per
example, ``emission[S, V]`` and ``transition[S, V]`` (``scorers.py:995-1035``), batched and
padded by ``collate`` (``util/dataset_reader.py:175-186``), and rebuilds the graph from
them with Python ``.item()`` loops on every call (``scorers.py:704-716``, ``:764-776``).  Here
the graph is built once per batch (cacheable per example) with the same edge rule

    cell (i, j) is an arc  i --j--> t   iff   t != 0 and t != i

and stored as int32 CSR by outgoing state ("out" order == canonical arc id == the
reference's scan order: state, then label) and by incoming state ("in" order), with the
states of each lattice renumbered by topological level so that a level is a contiguous
range.  States that are unreachable from the start state (row 0, ``scorers.py:1005``) --
which includes every row added by ``collate`` padding (quirk Q5) -- are trimmed; their
beta is reported as 0.  That equals the batched reference for the padding rows (no path to
the sink: it leaves 0 there).  A REAL state that the start cannot reach but that can reach
the sink is a documented difference: ``compute_beta_parallel`` runs its recurrence on every
row and leaves a non-zero beta there, this package reports 0.  The sampler never reads such
a state (it only gathers beta at successors of reachable states) and OpenFst's ``connect``
removes them from the machines the reference builds, so the golden lattices have none.

Two packers produce this layout: ``pack_small_device`` (the library's device packer,
``nfst_pack.cu``: three launches, for lattices that fit one SM's shared memory -- the size
nFST builds) and the tensor-op packer below (any size, any device; it also builds the chunk
lists, the sliced-column descriptors and the tile stream of wide lattices).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import os
from typing import List, Optional

import torch

from . import _lib
from . import tiles as tiles_mod


PACK_TIMING = int(os.environ.get("NFST_PACK_TIMING", "0"))
phase_ms = {}  # NFST_PACK_TIMING=1: milliseconds per packer phase (device-synchronised), accumulated
_phase_state = [None, 0.0]


def _phase(name, dev=None):
    """Timing hook (off by default): closes the running phase and opens ``name``."""
    if not PACK_TIMING:
        return
    import time

    if dev is not None and dev.type == "cuda":
        torch.cuda.synchronize(dev)
    now = time.perf_counter()
    if _phase_state[0] is not None:
        phase_ms[_phase_state[0]] = phase_ms.get(_phase_state[0], 0.0) + 1e3 * (now - _phase_state[1])
    if name is None:
        _phase_state[0] = None
        return
    _phase_state[0], _phase_state[1] = name, now


# shared-memory window of the most recent per-state DP values, in bytes per state vector
WINDOW_BYTES_MAX = int(os.environ.get("NFST_WINDOW_BYTES", str(64 * 1024)))
# chunk geometry (tunable): a chunk targets ARCS_PER_THREAD arcs per thread of a block of at
# most BLOCK_MAX threads; a state with more than target/HEAVY_DIV arcs gets a chunk of its own
BLOCK_MAX = int(os.environ.get("NFST_BLOCK_MAX", "256"))
ARCS_PER_THREAD = int(os.environ.get("NFST_ARCS_PER_THREAD", "4"))
HEAVY_DIV = 4
# lattices whose widest level has at least this many arcs run level-major (one launch per
# topological level over all their chunks) instead of one block per lattice
LEVEL_MODE_MIN_ARCS = int(os.environ.get("NFST_LEVEL_MODE_MIN", "4096"))
# lattices whose whole working set fits this much shared memory run in one piece (nfst_small_kernel)
SMALL_SMEM_BYTES = int(os.environ.get("NFST_SMALL_SMEM_BYTES", str(96 * 1024)))
SMALL_BLOCK_MIN = int(os.environ.get("NFST_SMALL_BLOCK", "32"))


def small_footprint_bytes(states, arcs, levels, vocab):
    """Upper bound of nfst_small_kernel's shared memory for one lattice (fp64 state, both
    semirings, labels and a dtheta histogram; state / arc indices are 16-bit)."""
    return 4 * (6 * states + 2 * (levels + 1) + (states + 2) + 4 * arcs + min(vocab, 4096) + 64)
DEGREE_SORT = int(os.environ.get("NFST_DEGREE_SORT", "1"))
# Sliced-column execution (nfst_sell.cu): lattices whose levels average at least SELL_MIN_WIDTH states and
# whose DP ring (SELL_WINDOW_QUANTILE of the arc spans, and the widest level) fits SELL_WINDOW_MAX states; the
# few longer arcs go through global memory.  Measured on B200 against the CSR kernels (config 4, B = 1024,
# fwd+bwd, sliced columns vs CSR): 10k arcs (40 states per level: two slices for four warps) 0.33 vs 0.25 ms;
# 30k arcs (119 per level) 0.45 vs 0.57 ms, Viterbi 0.16 vs 0.26 ms; 100k: 1.02 vs 1.42 ms; 300k: 3.60 vs
# 5.12 ms; lattices whose ring would not fit stay CSR (1M arcs per lattice).  The threshold sits where a level
# fills the four warps of a block (three full slices and a partial one).
SELL = int(os.environ.get("NFST_SELL", "1"))
SELL_MIN_WIDTH = int(os.environ.get("NFST_SELL_MIN_WIDTH", "96"))
SELL_WINDOW_MAX = int(os.environ.get("NFST_SELL_WINDOW_MAX", "16384"))
SELL_WINDOW_QUANTILE = float(os.environ.get("NFST_SELL_WINDOW_QUANTILE", "0.995"))
SELL_THREADS = int(os.environ.get("NFST_SELL_THREADS", "0"))  # 0 = from the level width (128, or 256 from 1024 states per level)


def chunk_geometry(block_threads: int):
    """(target, heavy threshold, stage capacity) in arcs for a block size."""
    target = ARCS_PER_THREAD * block_threads
    heavy = max(target // HEAVY_DIV, 8)
    cap = (target + heavy + 7) // 8 * 8
    return target, heavy, cap


@dataclasses.dataclass
class LaunchGroup:
    """Lattices that share one kernel launch (one thread block per lattice)."""

    ids: torch.Tensor  # int32 [n], lattice indices, heaviest first
    n: int
    block_threads: int  # 32 / 64 / 128 / 256; the lattices' chunks were cut for this size
    max_states: int  # largest lattice of the group, in states
    max_reach: int  # longest arc of the group, in packed state ids (sizes the shared-memory window)
    n_arcs: int
    chunk_cap: int = 0  # stage capacity the lattices' chunks were cut for
    # level-major execution (wide lattices): the group's chunks sorted by level
    n_levels: int = 0
    fwd_level_chunks: Optional[torch.Tensor] = None  # int32 [n, 4] device
    fwd_level_off: Optional[torch.Tensor] = None  # int32 [n_levels + 1] HOST
    bwd_level_chunks: Optional[torch.Tensor] = None
    bwd_level_off: Optional[torch.Tensor] = None
    bwd_level_lat: Optional[torch.Tensor] = None  # int32 [n] device
    # small-lattice execution: every lattice of the group fits in shared memory
    small_max_states: int = 0
    small_max_arcs: int = 0
    small_max_levels: int = 0
    # sliced-column execution (nfst_sell.cu): one block per lattice, arcs column-major per 32-state slice
    sell: bool = False
    sell_window: int = 0  # states in the shared-memory DP ring (a power of two)
    sell_far: bool = False  # some arc spans more than the ring: (end of dst's level - start of src's level) > window
    csr_block_threads: int = 0  # sliced-column groups: block size their in-order (CSR by destination) chunks were cut for
    # tile-stream execution (nfst_tiles.cu): block_threads = 32 * warps the lattices were dealt to
    tiles: bool = False
    tile_ring: int = 0  # largest DP ring of the group in slots: ring + constant slot + far table
    tile_far: int = 0  # largest far table of the group (destinations that outlive their ring slot)
    tile_cap_arcs: int = 0  # largest tile, in arcs
    tile_cap_bytes: int = 0  # largest tile, in stream bytes

    def to(self, device, non_blocking: bool = False) -> "LaunchGroup":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)  # noqa: E731
        return dataclasses.replace(self, ids=mv(self.ids), fwd_level_chunks=mv(self.fwd_level_chunks),
                                   bwd_level_chunks=mv(self.bwd_level_chunks), bwd_level_lat=mv(self.bwd_level_lat))

    def window_states(self, state_bytes: int = 4, window_bytes_max: int = WINDOW_BYTES_MAX) -> int:
        """Power-of-two window: covers every arc of the group if that fits the byte budget
        (then no neighbour value is ever re-read from global memory)."""
        cap = max(32, window_bytes_max // state_bytes)
        return max(32, min(_pow2_ceil(min(self.max_reach, self.max_states)), _pow2_floor(cap)))


class PackedLattices:
    """A batch of lattices in the layout ``nfst_packed_lattices_t`` describes."""

    _INT_FIELDS = (
        "state_off", "level_off", "level_ptr", "start_state", "sink_off", "sinks",
        "in_ptr", "src_in", "label_in", "in2out", "out_ptr", "dst_out", "label_out",
        "fwd_chunk_off", "fwd_chunks", "bwd_chunk_off", "bwd_chunks", "fwd_gather",
        "fwd_chunk_level", "bwd_chunk_level", "bwd_order", "sell_desc", "sell_lvl_slice",
        "tile_tab", "tile_lw_off", "tile_lat_info", "out_arc",
    )

    # arrays the kernels stage with 16-byte copies: kept zero-padded by PAD elements
    _ARC_FIELDS = ("src_in", "label_in", "in2out", "dst_out", "label_out", "in_ptr", "out_ptr", "out_deg8")
    PAD = 4

    @staticmethod
    def _padded(t: torch.Tensor) -> torch.Tensor:
        """Same values, but backed by storage with PAD trailing zeros (so that a 128-bit load
        that straddles the end reads zeros -- valid indices -- and never garbage)."""
        n = t.numel()
        room = t.untyped_storage().nbytes() // t.element_size() - t.storage_offset() - n
        if room >= PackedLattices.PAD and getattr(t, "_nfst_padded", False):
            return t
        buf = torch.zeros(n + PackedLattices.PAD, dtype=t.dtype, device=t.device)
        buf[:n] = t
        out = buf[:n]
        out._nfst_padded = True
        return out

    @staticmethod
    def alloc_padded(n: int, dtype, device) -> torch.Tensor:
        """An uninitialised length-n tensor with zeroed PAD slack behind it (for callers that
        refill the arc arrays every step, e.g. from pinned host memory)."""
        buf = torch.zeros(n + PackedLattices.PAD, dtype=dtype, device=device)
        out = buf[:n]
        out._nfst_padded = True
        return out

    def __init__(self, **kw):
        for f in self._ARC_FIELDS:
            kw[f] = self._padded(kw[f])
        self.n_lattices: int = kw["n_lattices"]
        self.n_states: int = kw["n_states"]
        self.n_arcs: int = kw["n_arcs"]
        self.vocab: int = kw["vocab"]
        for f in self._INT_FIELDS:
            setattr(self, f, kw[f])
        self.lanes_in_log2: torch.Tensor = kw["lanes_in_log2"]
        self.lanes_out_log2: torch.Tensor = kw["lanes_out_log2"]
        self.out_deg8: torch.Tensor = kw["out_deg8"]  # uint8 [S] min(out-degree, 255)
        self.tile_stream: torch.Tensor = kw["tile_stream"]  # uint8: tile headers, slice headers, 16-bit ring slots
        self.src_out: torch.Tensor = kw["src_out"]  # int32 [A] source state of each canonical arc (host-side view)
        self.orig_state: torch.Tensor = kw["orig_state"]  # int32 [S] original local state id
        self.arc_origin: torch.Tensor = kw["arc_origin"]  # int64 [A] index into the caller's arc list / dense cells
        self.arc_off: torch.Tensor = kw["arc_off"]  # int32 [B+1] canonical arc range per lattice
        self.n_levels: torch.Tensor = kw["n_levels"]  # int32 [B]
        self.static_scores: Optional[torch.Tensor] = kw.get("static_scores")  # float32 [A] (weighted tables)
        self.dense_shape = kw.get("dense_shape")  # (B, S, V) when packed from dense tables
        self.groups: List[LaunchGroup] = kw["groups"]
        self.max_levels: int = kw["max_levels"]
        self.stats = kw["stats"]  # per-lattice CPU tensors: arcs, states, levels, block_class
        self._c = None

    # ---- plumbing ----------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.state_off.device

    @property
    def has_sell(self) -> bool:
        return any(g.sell for g in self.groups)

    @property
    def has_tiles(self) -> bool:
        return any(g.tiles for g in self.groups)

    @property
    def has_columns(self) -> bool:
        """some lattices are stored column-major (sliced-column or tile-stream groups): their arcs are not CSR"""
        return any(g.sell or g.tiles for g in self.groups)

    @property
    def has_in_order(self) -> bool:
        return self.n_arcs == 0 or int(self.in2out.numel()) == self.n_arcs

    def ensure_in_order(self) -> "PackedLattices":
        """Build the arcs-by-destination arrays and the chunk lists of a batch that was packed without them (all of
        its lattices are column-major): what the CSR forward kernel needs to produce alpha for such lattices."""
        if self.has_in_order:
            return self
        dev, B, S = self.device, self.n_lattices, self.n_states
        state_off, level_ptr, level_off = self.state_off.to(torch.int64), self.level_ptr.to(torch.int64), self.level_off.to(torch.int64)
        lt_s = torch.repeat_interleave(torch.arange(B, device=dev), state_off[1:] - state_off[:-1])
        # level slot of every state: the last entry of level_ptr that is <= the state (a lattice's spare entry equals
        # the first entry of the next lattice; right=True picks the latter)
        slot = torch.searchsorted(level_ptr, torch.arange(S, device=dev), right=True) - 1
        csr = _in_order_and_chunks(
            in_ptr=self.in_ptr.to(torch.int64), out_ptr=self.out_ptr.to(torch.int64), src_out=self.src_out.to(torch.int64),
            dst_out=self.dst_out.to(torch.int64), label_out=self.label_out.to(torch.int64), slot=slot,
            lvl_first_state=level_ptr[slot], lt_s=lt_s, level_of_state=slot - level_off[lt_s],
            block_class=self.stats["block_class"].to(dev), n_lattices=B)
        for k, v in csr.items():
            setattr(self, k, self._padded(v) if k in self._ARC_FIELDS else v)
        self._c = None
        return self

    def tensors(self):
        names = list(self._INT_FIELDS) + ["lanes_in_log2", "lanes_out_log2", "out_deg8", "tile_stream", "src_out",
                                          "orig_state", "arc_origin", "arc_off", "n_levels"]
        if self.static_scores is not None:
            names.append("static_scores")
        return names

    def to(self, device, non_blocking: bool = False) -> "PackedLattices":
        kw = {k: getattr(self, k) for k in ("n_lattices", "n_states", "n_arcs", "vocab", "dense_shape", "max_levels")}
        for name in self.tensors():
            kw[name] = getattr(self, name).to(device, non_blocking=non_blocking)
        kw.setdefault("static_scores", None)
        kw["groups"] = [g.to(device, non_blocking) for g in self.groups]
        kw["stats"] = self.stats
        return PackedLattices(**kw)

    def pin_memory(self) -> "PackedLattices":
        for name in self.tensors():
            setattr(self, name, getattr(self, name).pin_memory())
        for g in self.groups:
            g.ids = g.ids.pin_memory()
        return self

    def nbytes(self) -> int:
        return sum(getattr(self, n).numel() * getattr(self, n).element_size() for n in self.tensors())

    def c_struct(self) -> "_lib.PackedLatticesC":
        """The C view of this batch (device pointers; valid while ``self`` is alive)."""
        if self._c is None:
            if self.device.type != "cuda":
                raise RuntimeError("nfst_b200 kernels need the packed lattices on a CUDA device (no CPU fallback)")
            c = _lib.PackedLatticesC()
            c.n_lattices, c.n_states, c.n_arcs, c.vocab = self.n_lattices, self.n_states, self.n_arcs, self.vocab
            for f in self._INT_FIELDS:
                if f.endswith("_chunk_level"):
                    continue  # host-side bookkeeping, not part of nfst_packed_lattices_t
                t = getattr(self, f)
                assert t.dtype == torch.int32 and t.is_contiguous() and t.data_ptr() % 16 == 0
                setattr(c, f, t.data_ptr() if t.numel() else None)
            c.lanes_in_log2 = self.lanes_in_log2.data_ptr()
            c.lanes_out_log2 = self.lanes_out_log2.data_ptr()
            c.out_deg8 = self.out_deg8.data_ptr()
            assert self.tile_stream.dtype == torch.uint8 and self.tile_stream.data_ptr() % 16 == 0
            c.tile_stream = self.tile_stream.data_ptr()
            self._c = c
        return self._c

    # ---- algorithmic byte counts ----------------------------
    def algorithmic_bytes_fwd_bwd(self) -> int:
        return 20 * self.n_arcs + 20 * self.n_states

    def algorithmic_bytes_viterbi(self, path_len_total: int = 0) -> int:
        return 8 * self.n_arcs + 12 * self.n_states + 8 * path_len_total


def _excl_cumsum(x: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(x.numel() + 1, dtype=torch.int64, device=x.device)
    torch.cumsum(x, 0, out=out[1:])
    return out


def _pow2_ceil(x: int) -> int:
    return 1 if x <= 1 else 1 << (x - 1).bit_length()


def _pow2_floor(x: int) -> int:
    return 1 << (max(int(x), 1).bit_length() - 1)


def _build_chunks(ptr, slot, level_first, lat_of_state, target, heavy_thr, n_lattices, descending):
    """Cut every level into chunks of consecutive states whose CSR segments END in the same
    `target`-sized window of the level's arc range (so a chunk has fewer than 2*target arcs
    unless it is a single heavy state, which gets a chunk of its own).

    ptr [S+1] CSR row pointer; slot [S] global level index of each state; level_first [S]
    first state of the state's level; target [S] per-state chunk target (its lattice's).
    Returns (chunk_off [B+1], chunks [NC, 4] = arc_begin, arc_end, state_begin, state_end).
    """
    S = ptr.numel() - 1
    dev = ptr.device
    deg = ptr[1:] - ptr[:-1]
    rel_end = ptr[1:] - ptr[level_first]
    ck = torch.clamp(torch.div(rel_end - 1, target, rounding_mode="floor"), min=0)
    # ... and at most `target` states (levels full of arc-less states, e.g. many sinks)
    ck = ck + torch.div(torch.arange(S, device=dev) - level_first, target, rounding_mode="floor")
    heavy = (deg > heavy_thr).to(torch.int64)
    hk = 2 * torch.cumsum(heavy, 0) - heavy
    new = torch.ones(S, dtype=torch.bool, device=dev)
    if S > 1:
        new[1:] = (slot[1:] != slot[:-1]) | (ck[1:] != ck[:-1]) | (hk[1:] != hk[:-1])
    sb = torch.nonzero(new).squeeze(1)
    se = torch.cat([sb[1:], torch.tensor([S], device=dev, dtype=sb.dtype)])
    chunks = torch.stack([ptr[sb], ptr[se], sb, se], dim=1)
    # processing order inside a chunk: its states sorted by degree, so that the 32 states a
    # warp reduces have equal trip counts even where the state numbering cannot provide it
    cid = torch.cumsum(new.to(torch.int64), 0) - 1
    dmax = int(deg.max()) + 1 if S else 1
    order = torch.argsort(cid * dmax + deg, stable=True)
    clat = lat_of_state[sb]
    chunk_off = _excl_cumsum(torch.bincount(clat, minlength=n_lattices))
    if descending:
        n = chunks.shape[0]
        pos = torch.arange(n, device=dev)
        rev = chunk_off[clat] + (chunk_off[clat + 1] - 1 - pos)
        chunks = chunks[rev]
    return chunk_off, chunks.to(torch.int32).contiguous(), order.to(torch.int32).contiguous()


def _in_order_and_chunks(*, in_ptr, out_ptr, src_out, dst_out, label_out, slot, lvl_first_state, lt_s, level_of_state,
                         block_class, n_lattices: int):
    """The arrays only the CSR kernels read: arcs by destination (``in2out``, ``src_in``, ``label_in``), the forward /
    backward chunk lists and the score-gather ranges.  Column-major batches build them on demand
    (``PackedLattices.ensure_in_order``): their own kernels read the out order only."""
    dev = out_ptr.device
    B, A = n_lattices, int(src_out.numel())
    in2out = torch.argsort(dst_out, stable=True)
    src_in, label_in = src_out[in2out], label_out[in2out]
    bmax = int(math.log2(BLOCK_MAX))
    geo_t = torch.tensor([chunk_geometry(1 << k) if k >= 5 else (0, 0, 0) for k in range(bmax + 1)], device=dev)
    target_state, heavy_state = geo_t[block_class, 0][lt_s], geo_t[block_class, 1][lt_s]
    fwd_chunk_off, fwd_chunks, _ = _build_chunks(in_ptr, slot, lvl_first_state, lt_s, target_state, heavy_state, B, False)
    bwd_chunk_off, bwd_chunks, bwd_order = _build_chunks(out_ptr, slot, lvl_first_state, lt_s, target_state, heavy_state, B, True)
    # canonical-id range [lo, hi) that the arcs of each forward chunk gather their scores from
    # (the kernel prefetches it into L2 several chunks ahead)
    nfc = int(fwd_chunks.shape[0])
    fc = fwd_chunks.to(torch.int64)
    n_arc_c = fc[:, 1] - fc[:, 0]
    g_lo = torch.full((nfc,), A, dtype=torch.int64, device=dev)
    g_hi = torch.zeros(nfc, dtype=torch.int64, device=dev)
    if A:
        cid = torch.repeat_interleave(torch.arange(nfc, device=dev), n_arc_c)  # fwd chunks tile the in-order arcs
        g_lo = g_lo.scatter_reduce(0, cid, in2out, reduce="amin")
        g_hi = g_hi.scatter_reduce(0, cid, in2out + 1, reduce="amax")
    g_lo = torch.where(n_arc_c > 0, g_lo, torch.zeros_like(g_lo))
    fwd_gather = torch.stack([g_lo, torch.maximum(g_hi, g_lo)], dim=1)
    i32 = lambda t: t.to(torch.int32).contiguous()  # noqa: E731
    return {
        "in2out": i32(in2out), "src_in": i32(src_in), "label_in": i32(label_in),
        "fwd_chunk_off": i32(fwd_chunk_off), "fwd_chunks": fwd_chunks, "bwd_chunk_off": i32(bwd_chunk_off),
        "bwd_chunks": bwd_chunks, "fwd_gather": i32(fwd_gather), "bwd_order": bwd_order,
        "fwd_chunk_level": i32(level_of_state[fwd_chunks[:, 2].to(torch.int64)]),
        "bwd_chunk_level": i32(level_of_state[bwd_chunks[:, 2].to(torch.int64)]),
    }


def _sell_block_log2(states: torch.Tensor, levels: torch.Tensor) -> torch.Tensor:
    """log2 of the warps per block of a sliced-column lattice: 4 warps, 8 from 1024 states per level
    (measured: 128 threads best at 400 states per level, 256 at 1200)."""
    if SELL_THREADS:
        return torch.full_like(states, max(int(math.log2(max(SELL_THREADS // 32, 1))), 0))
    width = states.to(torch.float64) / torch.clamp(levels, min=1).to(torch.float64)
    return torch.where(width >= 1024, torch.full_like(states, 3), torch.full_like(states, 2))


def build_groups(stats, dev, chunk_info=None) -> List[LaunchGroup]:
    """Partition the batch into launches: lattices cut for the same block size share a
    launch; heaviest lattices first (longest-processing-time order).  Lattices with wide
    levels form level-major groups (see LaunchGroup / nfst_launch_t)."""
    wide = (stats["width_arcs"] >= LEVEL_MODE_MIN_ARCS).to(torch.int64)
    foot = small_footprint_bytes(stats["states"], stats["arcs"], stats["levels"], int(stats["vocab"][0]))
    small = ((foot <= SMALL_SMEM_BYTES) & (stats["states"] < 65536) & (stats["arcs"] < 65536)).to(torch.int64) * (1 - wide)
    sell = stats["sell"].to(torch.int64)
    tile = stats["tile"].to(torch.int64)
    wide, small = wide * (1 - sell) * (1 - tile), small * (1 - sell) * (1 - tile)
    # sliced-column lattices: one group per block size (a power of two of warps); tile-stream lattices likewise
    gkey = torch.where(sell > 0, (stats["sell_block"] * 2) * 2 + 4096, (stats["block_class"] * 2 + wide) * 2 + small)
    # (deep tile-stream lattices -- float64 DP rings by default -- do not share a launch with shallow ones)
    deep = (stats["levels"] > tiles_mod.F64_LEVELS).to(torch.int64)
    gkey = torch.where(tile > 0, stats["tile_warps_log2"] + 8192 + 64 * deep, gkey)
    groups: List[LaunchGroup] = []
    B = int(gkey.numel())
    for key in sorted(set(gkey.tolist()), reverse=True):
        members = torch.nonzero(gkey == key).squeeze(1)
        members = members[torch.argsort(stats["arcs"][members], descending=True, stable=True)]
        if key >= 8192:
            groups.append(LaunchGroup(
                ids=members.to(torch.int32).to(dev), n=int(members.numel()), block_threads=32 << ((key - 8192) & 63),
                max_states=int(stats["states"][members].max()), max_reach=int(stats["reach"][members].max()),
                n_arcs=int(stats["arcs"][members].sum()), n_levels=int(stats["levels"][members].max()),
                chunk_cap=int(stats["chunk_cap"][members].max()), csr_block_threads=1 << int(stats["block_class"][members].max()),
                tiles=True, tile_ring=int(stats["tile_ring"][members].max()), tile_far=int(stats["tile_far"][members].max()),
                tile_cap_arcs=int(stats["tile_cap_arcs"][members].max()), tile_cap_bytes=int(stats["tile_cap_bytes"][members].max()),
            ))
            continue
        if key >= 4096:
            groups.append(LaunchGroup(
                ids=members.to(torch.int32).to(dev), n=int(members.numel()),
                block_threads=32 << ((key - 4096) >> 2),
                max_states=int(stats["states"][members].max()), max_reach=int(stats["reach"][members].max()),
                n_arcs=int(stats["arcs"][members].sum()), n_levels=int(stats["levels"][members].max()),
                chunk_cap=int(stats["chunk_cap"][members].max()),  # for the CSR forward kernel (exact alpha)
                csr_block_threads=1 << int(stats["block_class"][members].max()),
                sell=True, sell_window=int(stats["sell_window"][members].max()),
                sell_far=bool((stats["sell_bound"][members] > int(stats["sell_window"][members].max())).any()),
            ))
            continue
        g = LaunchGroup(
            ids=members.to(torch.int32).to(dev),
            n=int(members.numel()),
            block_threads=1 << (key >> 2),
            max_states=int(stats["states"][members].max()),
            max_reach=int(stats["reach"][members].max()),
            n_arcs=int(stats["arcs"][members].sum()),
            chunk_cap=int(stats["chunk_cap"][members].max()),
        )
        if key & 1:
            g.block_threads = max(g.block_threads, SMALL_BLOCK_MIN)
            g.small_max_states = int(stats["states"][members].max())
            g.small_max_arcs = max(int(stats["arcs"][members].max()), 1)
            g.small_max_levels = int(stats["levels"][members].max())
        if (key & 2) and chunk_info is not None:
            member_mask = torch.zeros(B, dtype=torch.bool, device=dev)
            member_mask[members.to(dev)] = True
            n_levels = int(stats["levels"][members].max())
            g.n_levels = n_levels
            for direction in ("fwd", "bwd"):
                off, chunks, level = chunk_info[direction]
                clat = torch.repeat_interleave(torch.arange(B, device=dev), (off[1:] - off[:-1]).to(torch.int64))
                sel = torch.nonzero(member_mask[clat]).squeeze(1)
                lv = level[sel].to(torch.int64)
                order = torch.argsort(lv, stable=True)
                lvl_off = _excl_cumsum(torch.bincount(lv, minlength=n_levels)).to(torch.int32).cpu().contiguous()
                setattr(g, f"{direction}_level_chunks", chunks[sel][order].contiguous())
                setattr(g, f"{direction}_level_off", lvl_off)
                if direction == "bwd":
                    g.bwd_level_lat = clat[sel][order].to(torch.int32).contiguous()
        groups.append(g)
    return groups


def concat_packed(parts: List["PackedLattices"]) -> "PackedLattices":
    """Collate independently packed lattices (e.g. cached per example) into one batch.

    This is the per-step collate when packs are cached per example, so it is written as a fixed
    number of tensor ops (one concatenation per array plus one gathered offset add), independent
    of the number of parts."""
    if not parts:
        raise ValueError("empty batch")
    dev = parts[0].device
    vocab = parts[0].vocab
    if any(p.vocab != vocab for p in parts):
        raise ValueError("all parts must share one vocabulary")
    P = len(parts)
    # parts packed without their in-order arrays (column-major lattices): the batch keeps it that way when all parts
    # agree, otherwise the lazy parts build theirs now
    lazy = not any(p.has_in_order and p.n_arcs for p in parts)
    if not lazy:
        for p in parts:
            p.ensure_in_order()
    # element counts of every index space, per part (host integers: no device sync)
    n_state = [p.n_states for p in parts]
    n_arc = [p.n_arcs for p in parts]
    n_lat = [p.n_lattices for p in parts]
    n_lvl = [int(p.level_ptr.numel()) for p in parts]
    n_sink = [int(p.sinks.numel()) for p in parts]
    n_fc = [int(p.fwd_chunks.shape[0]) for p in parts]
    n_bc = [int(p.bwd_chunks.shape[0]) for p in parts]
    n_sl = [int(p.sell_desc.shape[0]) for p in parts]
    n_tt = [int(p.tile_tab.shape[0]) for p in parts]
    n_lw = [int(p.tile_lw_off.numel()) - 1 for p in parts]
    n_ts = [int(p.tile_stream.numel()) for p in parts]  # multiples of 16 bytes

    def starts(counts):
        out, t = [], 0
        for c in counts:
            out.append(t)
            t += c
        return out, t

    (oS, S), (oA, A), (oB, B), (oL, Lv), (oK, n_sinks), (oF, nfc), (oC, nbc), (oQ, nsl) = (
        starts(c) for c in (n_state, n_arc, n_lat, n_lvl, n_sink, n_fc, n_bc, n_sl))
    if S >= 2**31 or A >= 2**31:
        raise ValueError("batch too large for int32 indices; shard it")
    table = torch.tensor([oS, oA, oL, oK, oF, oC, n_state, n_arc, n_lat, n_lvl, n_sink, n_fc, n_bc, oQ, n_sl],
                         dtype=torch.int64).to(dev, non_blocking=True)
    offS, offA, offL, offK, offF, offC = (table[i].to(torch.int32) for i in range(6))
    pid = torch.arange(P, device=dev)
    rep = lambda row, total: torch.repeat_interleave(pid, table[row], output_size=total)  # noqa: E731
    by_state, by_arc, by_lat, by_lvl, by_sink, by_fc, by_bc, by_sl = (
        rep(6, S), rep(7, A), rep(8, B), rep(9, Lv), rep(10, n_sinks), rep(11, nfc), rep(12, nbc), rep(14, nsl))
    offQ = table[13].to(torch.int32)

    def cat(name, cut=False):
        ts = [getattr(p, name) for p in parts]
        return torch.cat([t[:-1] for t in ts] if cut else ts)

    def closed(name, space, off, total):
        """an offset array [n+1] per part -> [N+1] for the batch"""
        body = cat(name, cut=True) + off[space]
        return torch.cat([body, torch.tensor([total], dtype=body.dtype, device=dev)])

    kw = {
        "state_off": closed("state_off", by_lat, offS, S),
        "level_off": closed("level_off", by_lat, offL, Lv),
        "level_ptr": cat("level_ptr") + offS[by_lvl],
        "start_state": cat("start_state") + offS[by_lat],
        "sink_off": closed("sink_off", by_lat, offK, n_sinks),
        "sinks": cat("sinks") + offS[by_sink],
        "in_ptr": closed("in_ptr", by_state, offA, A),
        "src_in": cat("src_in") + (0 if lazy else offS[by_arc]),
        "label_in": cat("label_in"),
        "in2out": cat("in2out") + (0 if lazy else offA[by_arc]),
        "out_ptr": closed("out_ptr", by_state, offA, A),
        "dst_out": cat("dst_out") + offS[by_arc],
        "label_out": cat("label_out"),
        "fwd_chunk_off": closed("fwd_chunk_off", by_lat, offF, nfc),
        "bwd_chunk_off": closed("bwd_chunk_off", by_lat, offC, nbc),
        "bwd_order": cat("bwd_order") + (0 if lazy else offS[by_state]),
        "fwd_chunk_level": cat("fwd_chunk_level"),
        "bwd_chunk_level": cat("bwd_chunk_level"),
        "lanes_in_log2": cat("lanes_in_log2"),
        "lanes_out_log2": cat("lanes_out_log2"),
        "out_deg8": cat("out_deg8"),
        "sell_lvl_slice": cat("sell_lvl_slice") + offQ[by_lvl],
        "src_out": cat("src_out") + offS[by_arc],
        "orig_state": cat("orig_state"),
        "arc_origin": cat("arc_origin"),
        "arc_off": closed("arc_off", by_lat, offA, A),
        "n_levels": cat("n_levels"),
    }
    for name, space, off_c in (("fwd_chunks", by_fc, offF), ("bwd_chunks", by_bc, offC)):
        shift = torch.stack([offA[space], offA[space], offS[space], offS[space]], dim=1)
        kw[name] = cat(name) + shift
    # tile-stream arrays: the stream is lattice-relative, only the tables shift
    oT, oW, oY = starts(n_tt)[0], starts(n_lw)[0], starts(n_ts)[0]
    tabs, lws, infos = [], [], []
    for i, p in enumerate(parts):
        t = p.tile_tab.clone()
        if t.shape[0]:
            t[:, 0] += oA[i]
            t[:, 1] += oY[i] // 16
        tabs.append(t)
        lws.append(p.tile_lw_off[:-1] + oT[i])
        inf = p.tile_lat_info.clone()
        is_tile = (inf[:, 1] > 0).to(inf.dtype)  # rows of other lattices stay zero
        inf[:, 0] += oW[i] * is_tile
        inf[:, 3] += oA[i] * is_tile
        infos.append(inf)
    kw["tile_tab"] = torch.cat(tabs)
    kw["tile_lw_off"] = torch.cat(lws + [torch.tensor([sum(n_tt)], dtype=torch.int32, device=dev)])
    kw["tile_lat_info"] = torch.cat(infos)
    kw["tile_stream"] = torch.cat([p.tile_stream for p in parts])
    if any(p.out_arc.numel() for p in parts):  # identity for the parts without column-major lattices
        kw["out_arc"] = torch.cat([(p.out_arc if p.out_arc.numel() else torch.arange(p.n_arcs, dtype=torch.int32, device=dev)) + oA[i]
                                   for i, p in enumerate(parts)])
    else:
        kw["out_arc"] = torch.zeros(0, dtype=torch.int32, device=dev)
    sd = cat("sell_desc")
    kw["sell_desc"] = sd + torch.stack([offA[by_sl], offA[by_sl], torch.zeros_like(offA[by_sl]), torch.zeros_like(offA[by_sl])], dim=1)
    fg = cat("fwd_gather")
    nonempty = (fg[:, 1] > fg[:, 0]).to(torch.int32).unsqueeze(1)  # empty chunks keep [0, 0)
    kw["fwd_gather"] = fg + offA[by_fc].unsqueeze(1) * nonempty
    kw = {k: v.contiguous() for k, v in kw.items()}
    static = [p.static_scores for p in parts if p.static_scores is not None]
    if static and len(static) != P:
        raise ValueError("either all or none of the parts may carry static scores")
    stats = {k: torch.cat([p.stats[k] for p in parts]) for k in parts[0].stats}
    ci = {d: (kw[f"{d}_chunk_off"].to(torch.int64), kw[f"{d}_chunks"], kw[f"{d}_chunk_level"]) for d in ("fwd", "bwd")}
    return PackedLattices(
        n_lattices=B, n_states=S, n_arcs=A, vocab=vocab, static_scores=torch.cat(static) if static else None,
        dense_shape=None, groups=build_groups(stats, dev, ci), max_levels=max(p.max_levels for p in parts),
        stats=stats, **kw,
    )


def pack_arcs(
    arc_lattice: torch.Tensor,
    src: torch.Tensor,
    dst: torch.Tensor,
    label: torch.Tensor,
    n_states: torch.Tensor,
    vocab: int,
    *,
    start_state: int = 0,
    static_scores: Optional[torch.Tensor] = None,
    dense_shape=None,
    sell: Optional[bool] = None,
    tiles: Optional[bool] = None,
    in_order: Optional[bool] = None,
) -> PackedLattices:
    """Pack an arc list.  ``arc_lattice/src/dst/label`` are [A0] integer tensors (local
    state ids), ``n_states`` is [B].  Raises ``ValueError`` for cyclic lattices (the
    reference's denominator / base machines, which it never feeds to the DP either).
    ``sell`` / ``tiles``: allow the sliced-column / tile-stream layouts for wide lattices (defaults: the NFST_SELL
    and NFST_TILES knobs; tiles win where both apply).  ``in_order``: build the arcs-by-destination arrays and the chunk
    lists now (default: only when some lattice runs on the CSR kernels; column-major batches build them on the first
    call that needs alpha, ``PackedLattices.ensure_in_order``)."""
    dev = src.device
    n_states = n_states.to(device=dev, dtype=torch.int64)
    B = int(n_states.numel())
    if B == 0:
        raise ValueError("empty batch")
    arc_lattice = arc_lattice.to(torch.int64)
    label = label.to(torch.int64)
    so = _excl_cumsum(n_states)  # original global state offsets
    # ---- small lattices on a GPU: the library's device packer (nfst_pack.cu); it checks the arc endpoints itself ----
    if DEVICE_PACK and dev.type == "cuda" and sell is None and tiles is None and in_order is not False and src.numel():
        cnt = torch.bincount(arc_lattice, minlength=B)
        head = torch.stack([so[-1], n_states.max(), cnt.max(), label.min(), label.max()]).cpu().tolist()  # one host read
        S0, smax, amax = head[0], head[1], head[2]
        if S0 >= 2**31 or src.numel() >= 2**31:
            raise ValueError("batch too large for int32 indices; shard it")
        if head[3] < 0 or head[4] >= vocab:
            raise ValueError("label out of range")
        if smax <= 65535 and B * smax * vocab < 2**62:
            perm = torch.argsort((arc_lattice * smax + src.to(torch.int64)) * vocab + label, stable=True)  # (lattice, source, label)
            i32c = lambda t: t[perm].to(torch.int32)  # noqa: E731
            packed = pack_small_device(so.to(torch.int32), _excl_cumsum(cnt).to(torch.int32), i32c(src), i32c(dst), i32c(label),
                                       vocab, src_is_global=False, max_states=smax, max_arcs=amax, n_states_raw=S0,
                                       start_state=start_state, dense_shape=dense_shape)
            if packed is not None:
                packed.arc_origin = perm[packed.arc_origin]
                if static_scores is not None:
                    packed.static_scores = static_scores[packed.arc_origin].to(torch.float32).contiguous()
                return packed
    S0 = int(so[-1])
    if S0 >= 2**31 or src.numel() >= 2**31:
        raise ValueError("batch too large for int32 indices; shard it")
    gsrc = so[arc_lattice] + src.to(torch.int64)
    gdst = so[arc_lattice] + dst.to(torch.int64)
    if src.numel() and (int(src.min()) < 0 or int(dst.min()) < 0 or bool((src >= n_states[arc_lattice]).any())
                        or bool((dst >= n_states[arc_lattice]).any())):
        raise ValueError("arc endpoint out of range")
    if label.numel() and (int(label.min()) < 0 or int(label.max()) >= vocab):
        raise ValueError("label out of range")

    _phase("pack: levels = longest distance from the start", dev)
    # ---- levels = longest distance from the start state; -1 = unreachable (trimmed) ----
    level = torch.full((S0,), -1, dtype=torch.int64, device=dev)
    level[so[:-1] + start_state] = 0
    max_iter = int(n_states.max()) if B else 0
    it = 0
    if dev.type == "cuda" and gsrc.numel():
        # the library's sweep kernel: in place, several levels settle per sweep; one flag read per round of 4 sweeps
        lib = _lib.load()
        level32 = level.to(torch.int32)
        changed = torch.zeros(1, dtype=torch.int32, device=dev)
        gs, gd = gsrc.contiguous(), gdst.contiguous()
        st = torch.cuda.current_stream(dev).cuda_stream
        while True:
            with torch.cuda.device(dev):
                _lib.check(lib.nfst_level_sweeps(gs.data_ptr(), gd.data_ptr(), gs.numel(), level32.data_ptr(), changed.data_ptr(), 4, st))
            if int(changed) == 0:
                break
            changed.zero_()
            it += 4
            if it > max_iter + 8:
                raise ValueError("lattice is cyclic: the DP is defined for acyclic lattices only")
        level = level32.to(torch.int64)
    while dev.type != "cuda" and gsrc.numel():
        before = level
        for _ in range(8):  # relaxation sweeps between convergence checks (each check is a host sync)
            ls = level[gsrc]
            level = level.scatter_reduce(0, gdst, torch.where(ls >= 0, ls + 1, ls), reduce="amax", include_self=True)
        if torch.equal(before, level):
            break
        it += 8
        if it > max_iter + 8:
            raise ValueError("lattice is cyclic: the DP is defined for acyclic lattices only")

    _phase("pack: state renumbering", dev)
    # ---- state renumbering: (lattice, level, degree key, original id) ----
    lat_of_state = torch.repeat_interleave(torch.arange(B, device=dev), n_states)
    kept = torch.nonzero(level >= 0).squeeze(1)
    lv = level[kept]
    lt = lat_of_state[kept]
    lmax = int(lv.max()) + 1 if kept.numel() else 1
    live = level[gsrc] >= 0
    deg_in = torch.bincount(gdst[live], minlength=S0)[kept]
    deg_out = torch.bincount(gsrc[live], minlength=S0)[kept]
    dmax = int(max(deg_in.max(), deg_out.max())) + 1 if kept.numel() else 1
    # level bookkeeping does not depend on the order inside a level
    n_levels = torch.zeros(B, dtype=torch.int64, device=dev).scatter_reduce(0, lt, lv + 1, reduce="amax")
    level_off = _excl_cumsum(n_levels + 1)
    n_slots = int(level_off[-1])
    slot_kept = level_off[lt] + lv
    counts = torch.bincount(slot_kept, minlength=n_slots)
    level_ptr = torch.cumsum(counts, 0) - counts  # exclusive; the spare slot of lattice b lands on state_off[b+1]
    S_b0 = torch.bincount(lt, minlength=B)

    _phase("pack: sliced-column eligibility", dev)
    # tile-stream eligibility (nfst_tiles.cu): levels at least a slice wide; every level must fit the largest ring
    tile_lat = torch.zeros(B, dtype=torch.bool, device=dev)
    tile_nw = torch.ones(B, dtype=torch.int64, device=dev)
    if (tiles_mod.TILES if tiles is None else tiles) and kept.numel():
        slot_lat1 = torch.repeat_interleave(torch.arange(B, device=dev), n_levels + 1)
        lvl_width1 = torch.zeros(B, dtype=torch.int64, device=dev).scatter_reduce(0, slot_lat1, counts, reduce="amax")
        # how far the arcs reach, in states (arcs into the last level read a constant): sizes the DP ring, and
        # with it the number of warps whose stages fit beside it
        slot_of1 = torch.zeros(S0, dtype=torch.int64, device=dev)
        slot_of1[kept] = slot_kept
        inner1 = live & (level[gdst] < n_levels[arc_lattice] - 1)
        la1 = arc_lattice[inner1]
        sd1, ss1 = slot_of1[gdst[inner1]], slot_of1[gsrc[inner1]]
        span1 = level_ptr[sd1] + counts[sd1] - level_ptr[ss1]
        tile_nw = tiles_mod.warps_per_lattice(S_b0, n_levels, tiles_mod.span_quantile(la1, span1, B))
        cap1 = tiles_mod.ring_cap_slots(tile_nw, n_levels)
        # slices are per (level, warp): every level an arc crosses may add a partial slice per warp
        est1 = span1 + 32 * tile_nw[la1] * (sd1 - ss1 + 1)
        far1 = torch.bincount(la1[est1 > (cap1[la1] * 3) // 4], minlength=B)
        # the ring must hold the widest level (every warp's share rounded up to whole slices) with room to spare, and
        # the destinations that outlive it must fit the far table
        tile_lat = (S_b0 >= tiles_mod.TILE_MIN_WIDTH * n_levels) & (lvl_width1 + 32 * tile_nw + 512 <= cap1) & (far1 <= cap1 // 8)
    # ---- sliced-column eligibility (nfst_sell.cu): wide levels, out-degree <= 255, bounded arc span ----
    sell_lat = torch.zeros(B, dtype=torch.bool, device=dev)
    sell_bound = torch.zeros(B, dtype=torch.int64, device=dev)
    sell_win = torch.full((B,), 32, dtype=torch.int64, device=dev)
    # (tile-stream lattices win where both layouts apply: when they take the whole batch this block is skipped)
    if (SELL if sell is None else sell) and kept.numel() and not bool(tile_lat.all()):
        slot_of = torch.full((S0,), 0, dtype=torch.int64, device=dev)
        slot_of[kept] = slot_kept
        # arcs into the LAST level need no ring: its states are final (beta = 0, delta = 0), the kernels
        # know them by their id
        inner = live & (level[gdst] < n_levels[arc_lattice] - 1)
        la = arc_lattice[inner]
        span = (level_ptr[slot_of[gdst[inner]]] + counts[slot_of[gdst[inner]]]) - level_ptr[slot_of[gsrc[inner]]]
        sell_bound = sell_bound.scatter_reduce(0, la, span, reduce="amax")
        sell_lat = S_b0 >= SELL_MIN_WIDTH * n_levels
        # ring size: covers SELL_WINDOW_QUANTILE of the lattice's arcs (log2 histogram of the spans); the few
        # longer ones -- e.g. dead ends wired to the sink -- go through global memory
        lgs = torch.ceil(torch.log2(span.to(torch.float64))).to(torch.int64).clamp_(0, 31)
        cum = torch.cumsum(torch.bincount(la * 32 + lgs, minlength=B * 32).view(B, 32), 1)
        need = torch.ceil(cum[:, -1:].to(torch.float64) * SELL_WINDOW_QUANTILE).to(torch.int64)
        sell_win = torch.clamp(torch.ones(B, dtype=torch.int64, device=dev) << (cum < need).sum(1), min=32)
        # ... and a whole level: two states of one level must never share a ring slot
        slot_lat0 = torch.repeat_interleave(torch.arange(B, device=dev), n_levels + 1)
        lvl_width = torch.zeros(B, dtype=torch.int64, device=dev).scatter_reduce(0, slot_lat0, counts, reduce="amax")
        pow2_width = torch.ones(B, dtype=torch.int64, device=dev) << torch.ceil(
            torch.log2(torch.clamp(lvl_width, min=1).to(torch.float64))).to(torch.int64)
        sell_win = torch.maximum(sell_win, pow2_width)
        sell_lat = sell_lat & (sell_win <= SELL_WINDOW_MAX) & ~tile_lat
    col_lat = sell_lat | tile_lat
    # inside a level states are ordered by (in-degree, out-degree): the 32 states a warp reduces then
    # have (nearly) equal segment lengths in both passes; column-major lattices by out-degree,
    # descending, so that the states of a slice that own a k-th arc are a prefix of the slice
    if DEGREE_SORT and lmax * dmax * dmax < 2**62 // max(B, 1):
        dkey = torch.where(col_lat[lt], (dmax - 1 - deg_out) * dmax, deg_in * dmax + deg_out)
        order = torch.argsort((lt * lmax + lv) * dmax * dmax + dkey, stable=True)
    else:  # pathological degrees: fall back to level order only
        sell_lat = torch.zeros_like(sell_lat)
        tile_lat = torch.zeros_like(tile_lat)
        col_lat = sell_lat
        order = torch.argsort(lt * lmax + lv, stable=True)
    st_w = st_j = st_lane = None
    if bool(tile_lat.any()):
        # tile-stream lattices: the slices of a level are dealt to the block's warps; (level, warp, slice, lane) order
        order, st_w, st_j, st_lane = tiles_mod.deal_order(order, lt, deg_out, slot_kept, level_ptr, tile_lat, tile_nw)
    kept_sorted = kept[order]
    lt_s, lv_s = lt[order], lv[order]
    S = int(kept_sorted.numel())
    new_id = torch.full((S0,), -1, dtype=torch.int64, device=dev)
    new_id[kept_sorted] = torch.arange(S, device=dev)
    state_off = _excl_cumsum(S_b0)
    slot = level_off[lt_s] + lv_s
    orig_state = kept_sorted - so[lt_s]
    start_packed = new_id[so[:-1] + start_state]

    _phase("pack: arcs", dev)
    # ---- arcs: canonical (out) order and in order ----
    origin = torch.nonzero(level[gsrc] >= 0).squeeze(1)
    ns, nd, lb = new_id[gsrc[origin]], new_id[gdst[origin]], label[origin]
    perm = torch.argsort(ns * vocab + lb, stable=True)
    src_out, dst_out, label_out, origin = ns[perm], nd[perm], lb[perm], origin[perm]
    A = int(src_out.numel())
    out_deg = torch.bincount(src_out, minlength=S)
    out_ptr = _excl_cumsum(out_deg)
    out_arc = torch.zeros(0, dtype=torch.int64, device=dev)  # empty = identity (no column-major lattice)
    if bool(col_lat.any()) and A:
        # column-major inside every slice: key (first arc of the slice, k, lane)
        pos = torch.arange(A, device=dev)
        k_in_state = pos - out_ptr[src_out]
        rel = src_out - level_ptr[slot[src_out]]
        lane = rel % 32
        if st_lane is not None:  # tile-stream lattices: slices are per (level, warp); heavy states are slices of their own
            lane = torch.where(tile_lat[lt_s[src_out]], st_lane[src_out], lane)
        is_sell = col_lat[lt_s[src_out]]
        zero = torch.zeros_like(pos)
        kmul = int(k_in_state[is_sell].max()) + 1 if bool(is_sell.any()) else 1
        if A * kmul * 32 >= 2**62:
            raise ValueError("batch too large to sort into slices; shard it")
        key = (torch.where(is_sell, out_ptr[src_out - lane], pos) * kmul + torch.where(is_sell, k_in_state, zero)) * 32 \
            + torch.where(is_sell, lane, zero)
        perm2 = torch.argsort(key, stable=True)
        src_out, dst_out, label_out, origin = src_out[perm2], dst_out[perm2], label_out[perm2], origin[perm2]
        # the arcs of ONE state in label order, for the consumers that walk a single state (sampler, beta-hat):
        # CSR position out_ptr[s] + k -> canonical id
        out_arc = torch.empty(A, dtype=torch.int64, device=dev)
        out_arc[perm2] = pos
    in_ptr = _excl_cumsum(torch.bincount(dst_out, minlength=S))
    sinks = torch.nonzero(out_deg == 0).squeeze(1)
    sink_lat = torch.searchsorted(state_off, sinks, right=True) - 1
    sink_off = _excl_cumsum(torch.bincount(sink_lat, minlength=B))
    arc_off = out_ptr[state_off]

    _phase("pack: slice descriptors of the sliced-column l", dev)
    # ---- slice descriptors of the sliced-column lattices (nfst_packed_lattices_t.sell_desc) ----
    sell_slot = torch.repeat_interleave(sell_lat, n_levels + 1)
    nsl_slot = torch.where(sell_slot, (counts + 31) // 32, torch.zeros_like(counts))
    slice_start = _excl_cumsum(nsl_slot)
    NS = int(slice_start[-1])
    sell_lvl_slice = slice_start[:-1]
    sell_desc = torch.zeros((NS, 4), dtype=torch.int64, device=dev)
    if NS:
        st_ids = torch.nonzero(sell_lat[lt_s]).squeeze(1)
        sid = slice_start[slot[st_ids]] + (st_ids - level_ptr[slot[st_ids]]) // 32
        first = torch.full((NS,), S, dtype=torch.int64, device=dev).scatter_reduce(0, sid, st_ids, reduce="amin")
        cnt = torch.bincount(sid, minlength=NS)
        dg = out_deg[st_ids]
        n_k = torch.stack([torch.bincount(sid[dg > k], minlength=NS) for k in range(7)], dim=1)
        cs = torch.cumsum(n_k, 1)  # start of column 1..7 (<= 224)
        dmax8 = torch.clamp(torch.zeros(NS, dtype=torch.int64, device=dev).scatter_reduce(0, sid, dg, reduce="amax"), max=255)
        sell_desc[:, 0] = out_ptr[first]
        sell_desc[:, 1] = out_ptr[first + cnt]
        sell_desc[:, 2] = cs[:, 0] | (cs[:, 1] << 8) | (cs[:, 2] << 16) | (cs[:, 3] << 24)
        sell_desc[:, 3] = cs[:, 4] | (cs[:, 5] << 8) | (cs[:, 6] << 16) | (dmax8 << 24)
        sell_desc = torch.where(sell_desc >= 2**31, sell_desc - 2**32, sell_desc)  # bit pattern as int32

    _phase("pack: tile-stream lattices", dev)
    # ---- tile-stream lattices: slices, ring slots, tiles, byte stream ----
    if st_w is None:
        st_w = st_j = torch.zeros(S, dtype=torch.int64, device=dev)
    tile_data = tiles_mod.build(lt_s=lt_s, slot=slot, level_off=level_off, level_ptr=level_ptr, n_levels=n_levels, st_w=st_w,
                                st_j=st_j, out_ptr=out_ptr, out_deg=out_deg, src_out=src_out, dst_out=dst_out,
                                tile_lat=tile_lat, tile_nw=tile_nw, state_off=state_off, n_lattices=B)

    _phase("pack: per-lattice shape statistics -> lanes pe", dev)
    # ---- per-lattice shape statistics -> lanes per state, block size, launch groups ----
    A_b = (arc_off[1:] - arc_off[:-1]).to(torch.float64)
    S_b = state_off[1:] - state_off[:-1]
    n_sinks_b = sink_off[1:] - sink_off[:-1]
    avg_in = A_b / torch.clamp(S_b - 1, min=1)
    avg_out = A_b / torch.clamp(S_b - n_sinks_b, min=1)

    def lanes_log2(avg):
        return torch.clamp(torch.round(torch.log2(torch.clamp(avg * 0.75, min=1.0))), 0, 5).to(torch.int64)

    lg_in, lg_out = lanes_log2(avg_in), lanes_log2(avg_out)
    # block size: 4 arcs per thread over the widest level (in either direction), 32..256 threads
    lvl_first_state = level_ptr[slot]  # [S] first state of each state's level
    slot_lat = torch.repeat_interleave(torch.arange(B, device=dev), n_levels + 1)
    lvl_end = torch.cat([level_ptr[1:], level_ptr.new_tensor([S])])
    lvl_arcs = torch.maximum(in_ptr[lvl_end] - in_ptr[level_ptr], out_ptr[lvl_end] - out_ptr[level_ptr])
    width_arcs = torch.zeros(B, dtype=torch.int64, device=dev).scatter_reduce(0, slot_lat, lvl_arcs, reduce="amax")
    bmax = int(math.log2(BLOCK_MAX))
    block_class = torch.clamp(torch.ceil(torch.log2(torch.clamp(
        width_arcs.to(torch.float64) / ARCS_PER_THREAD, min=32.0))), 5, bmax).to(torch.int64)
    geo = torch.tensor([chunk_geometry(1 << k) if k >= 5 else (0, 0, 0) for k in range(bmax + 1)], device=dev)
    # the in-order arrays and the chunk lists: read by the CSR kernels only -- a batch whose lattices are all
    # column-major leaves them to PackedLattices.ensure_in_order() (alpha of such a batch is asked for rarely)
    lazy = bool(col_lat.all()) if in_order is None else not in_order
    if lazy and not bool(col_lat.all()):
        raise ValueError("in_order=False needs a batch of column-major lattices only (the CSR kernels read the in-order arrays)")
    e32 = torch.zeros(0, dtype=torch.int32, device=dev)
    if lazy:
        csr = {"in2out": e32, "src_in": e32, "label_in": e32, "fwd_chunk_off": torch.zeros(B + 1, dtype=torch.int32, device=dev),
               "fwd_chunks": torch.zeros((0, 4), dtype=torch.int32, device=dev),
               "bwd_chunk_off": torch.zeros(B + 1, dtype=torch.int32, device=dev),
               "bwd_chunks": torch.zeros((0, 4), dtype=torch.int32, device=dev), "fwd_gather": torch.zeros((0, 2), dtype=torch.int32, device=dev),
               "bwd_order": e32, "fwd_chunk_level": e32, "bwd_chunk_level": e32}
    else:
        csr = _in_order_and_chunks(in_ptr=in_ptr, out_ptr=out_ptr, src_out=src_out, dst_out=dst_out, label_out=label_out,
                                   slot=slot, lvl_first_state=lvl_first_state, lt_s=lt_s, level_of_state=lv_s,
                                   block_class=block_class, n_lattices=B)
    # how far back (in packed state ids) an arc reaches: sizes the shared-memory window
    arc_lat = torch.repeat_interleave(torch.arange(B, device=dev), (arc_off[1:] - arc_off[:-1]))
    # (99% quantile over the lattice's arcs, rounded up to a power of two, from a log2
    # histogram -- a few long arcs, e.g. dead ends wired to the sink, take the global path)
    reach = torch.full((B,), 32, dtype=torch.int64, device=dev)
    if A:
        lg = torch.ceil(torch.log2((dst_out - src_out).to(torch.float64))).to(torch.int64).clamp_(0, 31)
        hist = torch.bincount(arc_lat * 32 + lg, minlength=B * 32).view(B, 32)
        cum = torch.cumsum(hist, 1)
        need = torch.ceil(cum[:, -1:].to(torch.float64) * 0.99).to(torch.int64)
        reach = torch.ones_like(reach) << (cum < need).sum(1)
    stats = {
        "width_arcs": width_arcs.cpu(),
        "arcs": A_b.to(torch.int64).cpu(),
        "states": S_b.cpu(),
        "levels": n_levels.cpu(),
        "block_class": block_class.cpu(),
        "vocab": torch.full((B,), int(vocab), dtype=torch.int64),
        "chunk_cap": geo[block_class, 2].cpu(),
        "reach": reach.cpu(),
        "sell": sell_lat.cpu(),
        "sell_bound": sell_bound.cpu(),
        "sell_window": sell_win.cpu(),
        "sell_block": _sell_block_log2(S_b.cpu(), n_levels.cpu()),
        "tile": tile_lat.cpu(),
        "tile_warps_log2": torch.round(torch.log2(tile_nw.to(torch.float64))).to(torch.int64).cpu(),
        **tile_data["stats"],
    }
    groups = build_groups(stats, dev, None if lazy else {
        "fwd": (csr["fwd_chunk_off"].to(torch.int64), csr["fwd_chunks"], csr["fwd_chunk_level"].to(torch.int64)),
        "bwd": (csr["bwd_chunk_off"].to(torch.int64), csr["bwd_chunks"], csr["bwd_chunk_level"].to(torch.int64))})

    _phase("pack: final int32 conversion", dev)
    i32 = lambda t: t.to(torch.int32).contiguous()  # noqa: E731
    packed_result = PackedLattices(
        n_lattices=B, n_states=S, n_arcs=A, vocab=int(vocab),
        state_off=i32(state_off), level_off=i32(level_off), level_ptr=i32(level_ptr), start_state=i32(start_packed),
        sink_off=i32(sink_off), sinks=i32(sinks), in_ptr=i32(in_ptr), src_in=csr["src_in"], label_in=csr["label_in"],
        in2out=csr["in2out"], out_ptr=i32(out_ptr), dst_out=i32(dst_out), label_out=i32(label_out),
        fwd_chunk_off=csr["fwd_chunk_off"], fwd_chunks=csr["fwd_chunks"], bwd_chunk_off=csr["bwd_chunk_off"],
        bwd_chunks=csr["bwd_chunks"], fwd_gather=csr["fwd_gather"], fwd_chunk_level=csr["fwd_chunk_level"],
        bwd_chunk_level=csr["bwd_chunk_level"], bwd_order=csr["bwd_order"], sell_desc=i32(sell_desc),
        sell_lvl_slice=i32(sell_lvl_slice),
        tile_stream=tile_data["tile_stream"], tile_tab=tile_data["tile_tab"], tile_lw_off=tile_data["tile_lw_off"],
        tile_lat_info=tile_data["tile_lat_info"], out_arc=i32(out_arc),
        lanes_in_log2=lg_in.to(torch.uint8).contiguous(), lanes_out_log2=lg_out.to(torch.uint8).contiguous(),
        out_deg8=torch.clamp(out_deg, max=255).to(torch.uint8).contiguous(), src_out=i32(src_out),
        orig_state=i32(orig_state), arc_origin=origin.contiguous(), arc_off=i32(arc_off), n_levels=i32(n_levels),
        static_scores=None if static_scores is None else static_scores[origin].to(torch.float32).contiguous(),
        dense_shape=dense_shape, groups=groups, max_levels=int(n_levels.max()) if B else 0, stats=stats,
    )
    _phase(None, dev)
    return packed_result


def _dense_arcs_cuda(transition: torch.Tensor):
    """(row, label, dst) int32 in scan order plus row_start int64 [B*S + 1], through the library's kernels."""
    B, S, V = transition.shape
    lib = _lib.load()
    tr = transition.to(torch.int64).contiguous()
    n_rows = B * S
    counts = torch.empty(n_rows, dtype=torch.int32, device=tr.device)
    st = torch.cuda.current_stream(tr.device).cuda_stream
    with torch.cuda.device(tr.device):
        _lib.check(lib.nfst_dense_count_arcs(tr.data_ptr(), n_rows, S, V, counts.data_ptr(), st))
        row_start = _excl_cumsum(counts.to(torch.int64))
        lat_off = row_start[::S]  # [B + 1]: raw arcs before lattice b
        head = torch.stack([row_start[-1], (lat_off[1:] - lat_off[:-1]).max()]).cpu()  # the one sync: sizes
        A0, max_arcs = int(head[0]), int(head[1])
        row = torch.empty(A0, dtype=torch.int32, device=tr.device)
        lab = torch.empty(A0, dtype=torch.int32, device=tr.device)
        dst = torch.empty(A0, dtype=torch.int32, device=tr.device)
        _lib.check(lib.nfst_dense_extract_arcs(tr.data_ptr(), n_rows, S, V, row_start.data_ptr(), A0, row.data_ptr(),
                                               lab.data_ptr(), dst.data_ptr(), st))
    return row, lab, dst, lat_off, max_arcs


def dense_arcs(transition: torch.Tensor):
    """Apply the reference's edge rule (``scorers.py:704-716``) to ``transition[B, S, V]``.

    Returns (row, label, dst) int tensors in scan order, ``row = b * S + s``.  CUDA inputs
    go through the library's kernels (one warp per table row); CPU inputs through
    ``torch.nonzero`` (host-side preprocessing only -- the DP itself has no CPU path).
    """
    if transition.dim() != 3:
        raise ValueError("transition must be [B, S, V]")  # scorers.py:878-879
    B, S, V = transition.shape
    if transition.device.type == "cuda":
        row, lab, dst, _, _ = _dense_arcs_cuda(transition)
        return row.to(torch.int64), lab.to(torch.int64), dst.to(torch.int64)
    rows = torch.arange(S, device=transition.device).view(1, S, 1)
    keep = (transition != 0) & (transition != rows)
    b, s, l = torch.nonzero(keep, as_tuple=True)
    return b * S + s, l, transition[b, s, l].to(torch.int64)


# pack small lattices on the device (nfst_pack.cu) instead of with the tensor-op packer below
DEVICE_PACK = int(os.environ.get("NFST_DEVICE_PACK", "1"))
launch_count = 0  # launches of the library's own pack kernels (diagnostics)


def _carve(pool: torch.Tensor, sizes):
    """Views into ``pool`` (int32 words), each starting on a 16-byte boundary; ``sizes`` = [(name, words)]."""
    out, at = {}, 0
    for name, n in sizes:
        out[name] = pool[at: at + n]
        at += (n + 3) & ~3
    return out, at


def _pool_words(sizes) -> int:
    return sum((n + 3) & ~3 for _, n in sizes)


def pack_small_device(raw_state_off: torch.Tensor, raw_arc_off: torch.Tensor, src: torch.Tensor, dst: torch.Tensor,
                      label: torch.Tensor, vocab: int, *, src_is_global: bool, max_states: int, max_arcs: int,
                      n_states_raw: int, start_state: int = 0, dense_shape=None) -> Optional[PackedLattices]:
    """Pack a batch of small lattices with the library's device packer (``nfst_pack_small``: three launches, one
    host read of the sizes between the count and the build phase, so every output has its exact size).
    ``src / dst / label``: int32 raw arcs grouped by lattice and sorted by (source, label) -- the scan order of the
    dense tables; ``dst`` local ids, ``src`` local or (``src_is_global``) global rows.  Returns None when the batch
    is not one for this packer (a lattice too large for one SM's shared memory, or one that would not run on the
    small-lattice kernels): the caller then uses the tensor-op packer.  ``arc_origin`` of the result indexes the raw
    arc list.  Raises ``ValueError`` for cyclic lattices and out-of-range arcs, like ``pack_arcs``."""
    global launch_count
    lib = _lib.load()
    dev = src.device
    B = int(raw_state_off.numel()) - 1
    if max_states > 65535:
        return None
    # max_arcs bounds the arcs a lattice KEEPS (the raw list may be far longer: collate() padding adds V arcs per pad
    # row, all unreachable); shared memory caps it -- a lattice that keeps more makes the pack return None
    room = (200 * 1024 - 32 * int(max_states) - 64) // 8
    if room < 256:
        return None
    max_raw = int(max_arcs)
    max_arcs = max(min(max_raw, 65535, room), 1)
    if B > 65535:
        return None
    S0, A0 = int(n_states_raw), int(src.numel())
    PAD = PackedLattices.PAD
    # ---- phase 1: levels, trimming, per-lattice counts and offsets (two launches) ----
    small = [("state_off", B + 1), ("arc_off", B + 1), ("level_off", B + 1), ("sink_off", B + 1), ("n_levels", B),
             ("start_state", B), ("lanes", (2 * B + 3) // 4), ("lattice_stats", 8 * B + 8)]
    o, _ = _carve(torch.zeros(_pool_words(small), dtype=torch.int32, device=dev), small)
    o["totals"] = o["lattice_stats"][8 * B:]
    out = _lib.PackOutC()
    for k in ("state_off", "arc_off", "level_off", "sink_off", "n_levels", "start_state", "lattice_stats", "totals"):
        setattr(out, k, o[k].data_ptr())
    ws_bytes = int(lib.nfst_pack_small_workspace_bytes(S0, A0))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    args = (B, raw_state_off.data_ptr(), raw_arc_off.data_ptr(), src.data_ptr(), dst.data_ptr(), label.data_ptr(),
            int(src_is_global), int(start_state), int(max_states))
    with torch.cuda.device(dev):
        _lib.check(lib.nfst_pack_small(*args, int(max_arcs), max_raw, C.byref(out), ws.data_ptr(), ws_bytes, S0, A0, 1, st))
    launch_count += 3
    host = o["lattice_stats"].cpu()  # the pack's host synchronisation
    totals, per = host[B * 8:].tolist(), host[:B * 8].view(B, 8).to(torch.int64)
    if totals[4] == 1:
        raise ValueError(f"lattice is cyclic: the DP is defined for acyclic lattices only (lattice {totals[5]})")
    if totals[4] == 2:
        raise ValueError(f"arc endpoint out of range (lattice {totals[5]})")
    if totals[4] == 4:
        return None  # a lattice keeps more arcs than one SM's shared memory holds
    if totals[4]:
        raise ValueError("batch too large for int32 indices; shard it")
    S, A, n_lp, n_sink = totals[0], totals[1], totals[2], totals[3]
    st_b, ar_b, lv_b = per[:, 0], per[:, 1], per[:, 2]
    # do these lattices belong on the small-lattice kernels?
    foot = small_footprint_bytes(st_b, ar_b, lv_b, int(vocab))
    other = (foot > SMALL_SMEM_BYTES) | (per[:, 4] >= LEVEL_MODE_MIN_ARCS)
    if tiles_mod.TILES:
        other |= st_b >= tiles_mod.TILE_MIN_WIDTH * lv_b
    if SELL:
        other |= st_b >= SELL_MIN_WIDTH * lv_b
    if bool(other.any()):
        return None  # some lattice belongs on the chunked / column-major kernels: those layouts come from pack_arcs
    # ---- phase 2: the packed arrays at their exact sizes (one launch) ----
    big = [("level_ptr", n_lp), ("sinks", n_sink), ("orig_state", S), ("in_ptr", S + 1 + PAD), ("out_ptr", S + 1 + PAD),
           ("src_in", A + PAD), ("label_in", A + PAD), ("in2out", A + PAD), ("dst_out", A + PAD), ("label_out", A + PAD),
           ("src_out", A), ("out_deg8", (S + PAD + 3) // 4), ("sell_lvl_slice", n_lp)]
    ob, _ = _carve(torch.zeros(_pool_words(big), dtype=torch.int32, device=dev), big)
    arc_origin = torch.empty(A, dtype=torch.int64, device=dev)
    deg8 = ob["out_deg8"].view(torch.uint8)
    for k in ("level_ptr", "sinks", "orig_state", "in_ptr", "out_ptr", "src_in", "label_in", "in2out", "dst_out", "label_out",
              "src_out"):
        setattr(out, k, ob[k].data_ptr())
    out.arc_origin, out.out_deg8 = arc_origin.data_ptr(), deg8.data_ptr()
    kept_max = max(int(ar_b.max()), 1)
    with torch.cuda.device(dev):
        _lib.check(lib.nfst_pack_small(*args, kept_max, max_raw, C.byref(out), ws.data_ptr(), ws_bytes, S0, A0, 2, st))
    launch_count += 1
    width_arcs = per[:, 4]  # arcs of the widest level: sizes the thread block of the DP kernels
    bmax = int(math.log2(BLOCK_MAX))
    block_class = torch.clamp(torch.ceil(torch.log2(torch.clamp(width_arcs.to(torch.float64) / ARCS_PER_THREAD, min=32.0))), 5, bmax).to(torch.int64)
    zeros = torch.zeros(B, dtype=torch.int64)
    stats = {
        "width_arcs": width_arcs, "arcs": ar_b.clone(), "states": st_b.clone(), "levels": lv_b.clone(),
        "block_class": block_class, "vocab": torch.full((B,), int(vocab), dtype=torch.int64),
        "chunk_cap": torch.tensor([chunk_geometry(1 << int(k))[2] for k in block_class.tolist()], dtype=torch.int64),
        "reach": st_b.clone(), "sell": torch.zeros(B, dtype=torch.bool), "sell_bound": zeros, "sell_window": zeros + 32,
        "sell_block": zeros + 2, "tile": torch.zeros(B, dtype=torch.bool), "tile_warps_log2": zeros, "tile_ring": zeros,
        "tile_far": zeros, "tile_cap_arcs": zeros, "tile_cap_bytes": zeros,
    }
    groups = build_groups(stats, dev)

    def arc_view(name, n):  # allocated with PAD zeroed elements behind the view
        t = ob[name][:n]
        t._nfst_padded = True
        return t

    i32 = dict(dtype=torch.int32, device=dev)
    e32, e4, lanes = torch.zeros(0, **i32), torch.zeros((0, 4), **i32), o["lanes"].view(torch.uint8)
    d8 = deg8[:S]
    d8._nfst_padded = True
    return PackedLattices(
        n_lattices=B, n_states=S, n_arcs=A, vocab=int(vocab),
        state_off=o["state_off"], level_off=o["level_off"], level_ptr=ob["level_ptr"], start_state=o["start_state"],
        sink_off=o["sink_off"], sinks=ob["sinks"], in_ptr=arc_view("in_ptr", S + 1), src_in=arc_view("src_in", A),
        label_in=arc_view("label_in", A), in2out=arc_view("in2out", A), out_ptr=arc_view("out_ptr", S + 1),
        dst_out=arc_view("dst_out", A), label_out=arc_view("label_out", A),
        fwd_chunk_off=torch.zeros(B + 1, **i32), fwd_chunks=e4, bwd_chunk_off=torch.zeros(B + 1, **i32), bwd_chunks=e4,
        fwd_gather=torch.zeros((0, 2), **i32), fwd_chunk_level=e32, bwd_chunk_level=e32,
        bwd_order=torch.arange(S, **i32), sell_desc=e4, sell_lvl_slice=ob["sell_lvl_slice"],
        tile_stream=torch.zeros(16, dtype=torch.uint8, device=dev), tile_tab=e4, tile_lw_off=torch.zeros(1, **i32),
        tile_lat_info=torch.zeros((B, 4), **i32), out_arc=e32, lanes_in_log2=lanes[:B], lanes_out_log2=lanes[B:2 * B],
        out_deg8=d8, src_out=ob["src_out"], orig_state=ob["orig_state"], arc_origin=arc_origin, arc_off=o["arc_off"],
        n_levels=o["n_levels"], static_scores=None, dense_shape=dense_shape, groups=groups, max_levels=int(totals[6]), stats=stats,
    )


def pack_dense(emission: Optional[torch.Tensor], transition: torch.Tensor, *,
               weighted: Optional[bool] = None, sell: Optional[bool] = None, tiles: Optional[bool] = None) -> PackedLattices:
    """Pack collate()-style dense tables ``emission[B, S, V]`` / ``transition[B, S, V]``.

    ``weighted``: treat ``emission`` as float log-weights (``scorers.py:1011-1013,1026-1027``)
    that become static arc scores; default = emission is a floating tensor.  ``sell`` / ``tiles``: as in ``pack_arcs``.
    """
    if transition.dim() != 3 or (emission is not None and emission.shape != transition.shape):
        raise ValueError("emission and transition must both be [B, S, V]")
    B, S, V = transition.shape
    if weighted is None:
        weighted = emission is not None and emission.is_floating_point()
    if transition.device.type == "cuda" and DEVICE_PACK and B * S < 2**31 and (sell is None and tiles is None):
        # device packer (nfst_pack.cu): the tables are scanned, levelled and laid out in a handful of launches
        row32, lab32, dst32, lat_off, max_arcs = _dense_arcs_cuda(transition)
        if row32.numel() < 2**31:
            raw_state_off = torch.arange(B + 1, dtype=torch.int32, device=transition.device) * S
            try:
                packed = pack_small_device(raw_state_off, lat_off.to(torch.int32), row32, dst32, lab32, V, src_is_global=True,
                                           max_states=S, max_arcs=max_arcs, n_states_raw=B * S, dense_shape=(B, S, V))
            except ValueError as e:
                raise ValueError("transition points outside the table" if "out of range" in str(e) else str(e)) from None
            if packed is not None:
                ao = packed.arc_origin  # raw arc index of every kept arc -> its dense cell (b*S+s)*V+l
                cell = row32[ao].to(torch.int64) * V + lab32[ao].to(torch.int64)
                if weighted:
                    packed.static_scores = emission.reshape(-1)[cell].to(torch.float32).contiguous()
                packed.arc_origin = cell.contiguous()
                return packed
        row, lab, dst = row32.to(torch.int64), lab32.to(torch.int64), dst32.to(torch.int64)
    else:
        row, lab, dst = dense_arcs(transition)
    if dst.numel() and int(dst.max()) >= S:
        raise ValueError("transition points outside the table")
    static = None
    if weighted:
        static = emission.reshape(-1)[row * V + lab].to(torch.float32)
    n_states = torch.full((B,), S, dtype=torch.int64, device=transition.device)
    packed = pack_arcs(row // S, row % S, dst, lab, n_states, V, static_scores=static, dense_shape=(B, S, V), sell=sell,
                       tiles=tiles)
    # arc_origin indexes the arc list; turn it into the dense cell index (b*S+s)*V+l
    packed.arc_origin = (row * V + lab)[packed.arc_origin].contiguous()
    return packed
