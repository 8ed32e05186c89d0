"""PyTorch operator surface over the C ABI (include/nfst_b200.h).

Mirrors what the reference's modules ask of the lattice:

* ``LatticeLogPartition`` / ``lattice_log_partition`` -- exact log-marginal ``[B]`` of an
  arc-factored model, the quantity ``Estimators.iwae`` estimates by importance sampling
  (``src/modules/estimatros.py:32-44`` via ``lightning.py:408-440``); its autograd
  gradient w.r.t. the arc scores is the arc posterior (which the reference's own DP cannot
  provide: autograd through ``scorers.py:744-747`` raises, quirk Q7).
* ``lattice_forward_backward`` -- (logZ, alpha, beta, posteriors).
* ``lattice_viterbi`` -- best path; replaces best-of-k-samples (``lightning.py:474-479``).
* ``compute_beta`` -- real-space ``beta[B*k, S]`` with the layout of
  ``FSAGRUScorer.compute_beta`` (``scorers.py:854,858-875``).

Torch is used for device memory and streams only; all arithmetic runs in the library's
sm_100a kernels.  CPU tensors are rejected -- there is no fallback.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _lib
from . import pack as pack_mod
from . import tiles as tiles_mod
from .pack import LaunchGroup, PackedLattices, pack_dense

# number of library kernels launched since import (bench.py reports the per-step count)
launch_count = 0


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# per-state-vector shared-memory window (bytes); tests shrink it to exercise the
# global-memory path behind the window
WINDOW_BYTES_MAX = pack_mod.WINDOW_BYTES_MAX
# lattices deeper than this default to float64 state vectors (see resolve_state_dtype): column-major lattices
# (their kernels form every arc term as a float32 OFFSET from a reference arc) / all others (plain float32 log-values)
F64_DEPTH = tiles_mod.F64_LEVELS  # 96: the tile-stream packer sizes the DP rings of such lattices for 8-byte slots
# shared memory of all SMs that a small-lattice group's one-launch forward+backward may claim (148 SMs x ~200 KB)
SMALL_RESIDENT_BYTES = int(os.environ.get("NFST_SMALL_RESIDENT_BYTES", str(148 * 200 * 1024)))
F64_DEPTH_PLAIN = int(os.environ.get("NFST_F64_DEPTH_PLAIN", "32"))
# depth of the tile-stream kernels' stage rings; 0 = chosen by the library from the shared memory per block
TILE_STAGES = int(os.environ.get("NFST_TILE_STAGES", "0"))
# width of the tile-stream flow pass's fixed-point accumulator (see nfst_launch_t.tile_flow_bits): 32 or 64
FLOW_BITS = int(os.environ.get("NFST_FLOW_BITS", "32"))


def resolve_state_dtype(packed: PackedLattices, state_dtype="auto") -> torch.dtype:
    """float32 or float64 for alpha / beta / logZ.

    An fp32 log-value x is only known to ulp(|x|)/2 ~ 6e-8*|x| and that rounding is
    committed at every level, so posteriors of deep lattices (|alpha| in the hundreds or
    thousands) cannot be 1e-5-accurate with fp32 state.  "auto" therefore uses float64 for
    batches with a lattice deeper than F64_DEPTH = 96 levels -- F64_DEPTH_PLAIN = 32 for lattices outside the
    column-major layouts, whose kernels form posteriors as exp(alpha + w + beta - logZ) from float32 log-values
    rounded at every level: random batches measured up to 2e-5 on 40-63-level edit lattices and 3e-5 on 43-level
    cipher lattices (|alpha| + |beta| = 260) with float32 state (tests/fuzz/fuzz_gpu.py) -- and float32 otherwise.  The
    column-major kernels take posteriors from per-state conditionals and a fixed-point flow instead, which does not
    accumulate that rounding: 5e-6 at 64 levels.  float64 state costs the small-lattice kernel 17 % at B = 32 and
    50 % at B = 4096 (shared memory per lattice); ``state_dtype=torch.float32`` buys that back where 2e-5 is enough.
    """
    if state_dtype == "auto" or state_dtype is None:
        if packed.max_levels > F64_DEPTH:
            return torch.float64
        if packed.max_levels > F64_DEPTH_PLAIN:
            lv, st = packed.stats.get("levels"), packed.stats
            if lv is None or "tile" not in st or "sell" not in st:
                return torch.float64
            plain = ~(st["tile"].bool() | st["sell"].bool())
            if bool((plain & (lv > F64_DEPTH_PLAIN)).any()):
                return torch.float64
        return torch.float32
    if state_dtype in (torch.float32, torch.float64):
        return state_dtype
    raise ValueError("state_dtype must be 'auto', torch.float32 or torch.float64")


def _check_f32(name: str, t: Optional[torch.Tensor], n: int, dev: torch.device) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.device != dev:
        raise RuntimeError(f"{name} must live on {dev} (nfst_b200 has no CPU fallback), got {t.device}")
    if t.numel() != n:
        raise ValueError(f"{name} must have {n} elements, got {t.numel()}")
    t = t.detach().to(torch.float32).contiguous()
    if t.data_ptr() % 16:
        t = t.clone()  # the kernels read scores with 128-bit loads (values past the end are never used)
    return t


def _launch(g: LaunchGroup, st_dtype: torch.dtype) -> "_lib.LaunchC":
    c = _lib.LaunchC()
    c.lattice_ids = g.ids.data_ptr()
    c.n_ids = g.n
    c.block_threads = g.block_threads
    f64 = st_dtype == torch.float64
    c.window_states = g.window_states(8 if f64 else 4, WINDOW_BYTES_MAX)
    c.state_f64 = int(f64)
    c.chunk_cap = g.chunk_cap
    c.small_max_states, c.small_max_arcs, c.small_max_levels = g.small_max_states, g.small_max_arcs, g.small_max_levels
    if g.sell:  # sliced-column group: the ring must cover every arc (no global-memory path behind it)
        c.sell = 1
        c.sell_far = int(g.sell_far)
        c.window_states = g.sell_window
        c.n_levels = g.n_levels
    if g.tiles:  # tile-stream group: ring slots and tile sizes were fixed at pack time
        c.tiles = 1
        c.tile_ring, c.tile_far = g.tile_ring, int(g.tile_far)
        c.tile_cap_arcs, c.tile_cap_bytes = g.tile_cap_arcs, g.tile_cap_bytes
        c.tile_stages = TILE_STAGES
        c.tile_flow_bits = FLOW_BITS
        c.n_levels = g.n_levels
    if g.fwd_level_chunks is not None:  # level-major group: one launch per topological level
        c.n_levels = g.n_levels
        c.fwd_level_chunks = g.fwd_level_chunks.data_ptr()
        c.fwd_level_off = g.fwd_level_off.data_ptr()  # host array
        c.bwd_level_chunks = g.bwd_level_chunks.data_ptr()
        c.bwd_level_off = g.bwd_level_off.data_ptr()
        c.bwd_level_lat = g.bwd_level_lat.data_ptr()
    return c


def _launch_csr_forward(g: LaunchGroup, st_dtype: torch.dtype) -> "_lib.LaunchC":
    """A sliced-column group as the CSR forward kernel sees it (its in-order arrays are ordinary CSR by
    destination): used where alpha itself is asked for -- the sliced-column passes never compute it."""
    c = _lib.LaunchC()
    c.lattice_ids = g.ids.data_ptr()
    c.n_ids = g.n
    c.block_threads = g.csr_block_threads
    f64 = st_dtype == torch.float64
    c.window_states = g.window_states(8 if f64 else 4, WINDOW_BYTES_MAX)
    c.state_f64 = int(f64)
    c.chunk_cap = g.chunk_cap
    return c


def _scores(packed: PackedLattices, arc_scores, theta):
    dev = packed.device
    if dev.type != "cuda":
        raise RuntimeError("nfst_b200 kernels need the packed lattices on a CUDA device (no CPU fallback)")
    a = _check_f32("arc_scores", arc_scores, packed.n_arcs, dev)
    if packed.static_scores is not None:
        a = packed.static_scores if a is None else (a + packed.static_scores)
    if packed.n_arcs == 0:
        a = None
    t = _check_f32("theta", theta, packed.vocab, dev)
    if a is not None and t is not None and _two_arrays_do_not_fit(packed):
        # per-arc scores AND theta: the tile-stream pull pass stages two arrays per tile; next to a DP ring that fills
        # the shared memory (1M-arc lattices) they do not fit, so the same sum w = score + theta[label] (one float32
        # add, as in the kernels) is formed once here and the pass runs on per-arc scores alone
        a = a + t[packed.label_out.long()]
        t = None
    c = _lib.ScoresC()
    c.arc_scores = _ptr(a)
    c.theta = _ptr(t)
    return c, (a, t)  # keep the tensors alive for the duration of the call


def _two_arrays_do_not_fit(packed: PackedLattices) -> bool:
    """some tile-stream group of the batch has no room for two staged per-arc arrays (cached on the batch)"""
    cached = getattr(packed, "_two_arrays_overflow", None)
    if cached is None:
        lib = _lib.load()
        cached = False
        for g in packed.groups:
            if g.tiles:
                for dt in (torch.float32, torch.float64):
                    need = lib.nfst_tile_smem_bytes(_launch(g, dt), packed.vocab, 0, 2, 1, 0)
                    cached = cached or need > 227 * 1024
        packed._two_arrays_overflow = cached
    return cached


def _gamma_far(packed: PackedLattices, alpha: bool = False) -> Optional[torch.Tensor]:
    """zero-filled float32 [S] for the flow pass of groups whose ring does not cover every arc (and of every
    sliced-column group when alpha is wanted: the flow into the last level is only needed for it)"""
    if any(g.sell and (g.sell_far or alpha) for g in packed.groups):
        return torch.zeros(packed.n_states, dtype=torch.float32, device=packed.device)
    return None


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


# Launch groups of one call are independent (disjoint lattices), and the small / deep ones are
# latency-bound: they run concurrently, group 0 on the caller's stream and the others on side
# streams that fork from it and are joined before the call returns.
PARALLEL_GROUPS = int(os.environ.get("NFST_PARALLEL_GROUPS", "1"))
_side_streams = {}


class _GroupStreams:
    def __init__(self, dev, n: int):
        self.main = torch.cuda.current_stream(dev)
        self.streams = [self.main]
        if n > 1 and PARALLEL_GROUPS:
            pool = _side_streams.setdefault(dev.index, [])
            while len(pool) < n - 1:
                pool.append(torch.cuda.Stream(dev))
            for s in pool[: n - 1]:
                s.wait_stream(self.main)
                self.streams.append(s)

    def __getitem__(self, i: int) -> int:
        return self.streams[i % len(self.streams)].cuda_stream

    def join(self):
        for s in self.streams[1:]:
            self.main.wait_stream(s)


def lattice_forward(packed: PackedLattices, arc_scores=None, theta=None, *, state_dtype="auto"
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """alpha[S] (log space, packed state order) and logZ[B], in the state dtype."""
    global launch_count
    lib = _lib.load()
    dev = packed.device
    sc, keep = _scores(packed, arc_scores, theta)
    st = resolve_state_dtype(packed, state_dtype)
    alpha = torch.empty(packed.n_states, dtype=st, device=dev)
    logz = torch.empty(packed.n_lattices, dtype=st, device=dev)
    packed.ensure_in_order()  # column-major batches are packed without their arcs-by-destination arrays
    with torch.cuda.device(dev):
        streams = _GroupStreams(dev, len(packed.groups))
        for i, g in enumerate(packed.groups):
            # sliced-column groups: their in-order arrays are plain CSR by destination, the forward kernel runs on them
            lc = _launch_csr_forward(g, st) if (g.sell or g.tiles) else _launch(g, st)
            _lib.check(lib.nfst_fwd_f32(packed.c_struct(), lc, sc, alpha.data_ptr(), logz.data_ptr(), streams[i]))
            launch_count += 1
        streams.join()
    del keep
    return alpha, logz


def lattice_pull(packed: PackedLattices, arc_scores=None, theta=None, *, state_dtype="auto",
                 beta_out: Optional[torch.Tensor] = None):
    """First half of the exact log-marginal: ``logZ[B]`` plus what the gradient pass needs.

    Returns ``(logz, alpha, cond)``: ``alpha[S]`` (forward pass) for the CSR launch groups, ``cond[A]`` --
    the probability of every arc given its source state, ``exp(w + beta[dst] - beta[src])`` -- for the
    sliced-column groups (beta pass; their posteriors then need no second logsumexp).  Either may be None.
    ``beta_out[S]`` (state dtype) additionally receives beta of the sliced-column groups."""
    global launch_count
    lib = _lib.load()
    dev = packed.device
    sc, keep = _scores(packed, arc_scores, theta)
    st = resolve_state_dtype(packed, state_dtype)
    any_csr = any(not (g.sell or g.tiles) for g in packed.groups)
    alpha = torch.empty(packed.n_states, dtype=st, device=dev) if any_csr else None
    cond = torch.empty(packed.n_arcs, dtype=torch.float32, device=dev) if packed.has_columns else None
    logz = torch.empty(packed.n_lattices, dtype=st, device=dev)
    if beta_out is not None and (beta_out.dtype != st or beta_out.numel() != packed.n_states or beta_out.device != dev):
        raise ValueError("beta_out must be [S] in the state dtype on the lattices' device")
    if beta_out is None and any(g.sell and g.sell_far for g in packed.groups):
        beta_out = torch.empty(packed.n_states, dtype=st, device=dev)  # arcs longer than the ring re-read beta
    with torch.cuda.device(dev):
        streams = _GroupStreams(dev, len(packed.groups))
        for i, g in enumerate(packed.groups):
            lc = _launch(g, st)
            if g.tiles:
                _lib.check(lib.nfst_tile_pull_f32(packed.c_struct(), lc, sc, _ptr(beta_out), logz.data_ptr(),
                                                  cond.data_ptr(), None, None, None, streams[i]))
            elif g.sell:
                _lib.check(lib.nfst_sell_pull_f32(packed.c_struct(), lc, sc, _ptr(beta_out), logz.data_ptr(),
                                                  cond.data_ptr(), None, None, None, streams[i]))
            else:
                _lib.check(lib.nfst_fwd_f32(packed.c_struct(), lc, sc, alpha.data_ptr(), logz.data_ptr(), streams[i]))
            launch_count += 1
        streams.join()
    del keep
    return logz, alpha, cond


def lattice_backward(
    packed: PackedLattices,
    arc_scores=None,
    theta=None,
    *,
    alpha: Optional[torch.Tensor] = None,
    logz: Optional[torch.Tensor] = None,
    grad_logz: Optional[torch.Tensor] = None,
    cond: Optional[torch.Tensor] = None,
    want_beta: bool = True,
    want_post: bool = False,
    want_dtheta: bool = False,
    want_viterbi: bool = False,
    state_dtype="auto",
):
    """Fused backward pass.  Returns a dict with the requested outputs among
    ``beta[S]``, ``logz_bwd[B]`` (state dtype), ``post[A]``, ``dtheta[V]`` (float32),
    ``delta[S]``, ``backptr[S]``, ``vit_score[B]`` (float32 / int32).

    Posteriors of CSR launch groups need ``alpha`` / ``logz`` (``lattice_forward`` or ``lattice_pull``);
    sliced-column groups compute them from ``cond`` (``lattice_pull``; recomputed here when not given)."""
    global launch_count
    lib = _lib.load()
    dev = packed.device
    sc, keep = _scores(packed, arc_scores, theta)
    S, A, B, V = packed.n_states, packed.n_arcs, packed.n_lattices, packed.vocab
    logs = want_beta or want_post or want_dtheta
    any_csr = any(not (g.sell or g.tiles) for g in packed.groups)
    if (want_post or want_dtheta) and any_csr and (alpha is None or logz is None):
        raise ValueError("posteriors need alpha and logz from lattice_forward")
    if alpha is not None and logz is None:
        raise ValueError("alpha needs the logz of the same lattice_forward / lattice_pull call")
    st = alpha.dtype if alpha is not None else (logz.dtype if logz is not None else resolve_state_dtype(packed, state_dtype))
    if alpha is not None and (logz.dtype != st or alpha.numel() != S or logz.numel() != B):
        raise ValueError("alpha / logz must come from lattice_forward on the same packed batch")
    out = {}
    f32 = dict(dtype=torch.float32, device=dev)
    beta = torch.empty(S, dtype=st, device=dev) if logs else None
    # column-major groups only run their log-semiring pull pass when beta or the conditionals are missing: their
    # entries of logz_bwd then come from the caller's logz
    logz_bwd = (logz.to(st).clone() if logz is not None else torch.empty(B, dtype=st, device=dev)) if logs else None
    post = torch.empty(A, **f32) if want_post else None
    dtheta = torch.zeros(V, **f32) if want_dtheta else None
    delta = torch.empty(S, **f32) if want_viterbi else None
    backptr = torch.empty(S, dtype=torch.int32, device=dev) if want_viterbi else None
    vit = torch.empty(B, **f32) if want_viterbi else None
    g32 = _check_f32("grad_logz", grad_logz, B, dev)
    flow = (want_post or want_dtheta) and packed.has_columns
    have_cond = cond is not None
    gfar = _gamma_far(packed) if flow else None
    if flow and not have_cond:
        cond = post if want_post else torch.empty(A, **f32)  # the flow pass may run in place
    with torch.cuda.device(dev):
        streams = _GroupStreams(dev, len(packed.groups))
        for i, g in enumerate(packed.groups):
            if g.sell or g.tiles:
                lc = _launch(g, st)
                far = g.sell and g.sell_far  # tile-stream groups keep far destinations in shared memory
                # the log-semiring pull pass runs for beta / logZ, and for the conditionals when the caller has none
                log_pull = want_beta or (flow and not have_cond)
                if log_pull or want_viterbi:
                    pull = lib.nfst_tile_pull_f32 if g.tiles else lib.nfst_sell_pull_f32
                    _lib.check(pull(
                        packed.c_struct(), lc, sc, _ptr(beta) if log_pull and (want_beta or far) else None,
                        _ptr(logz_bwd) if log_pull else None,
                        _ptr(cond) if (flow and not have_cond) else None, _ptr(delta), _ptr(backptr), _ptr(vit), streams[i]))
                    launch_count += int(log_pull) + int(want_viterbi)
                if flow:
                    if g.tiles:  # without want_post the tile-stream flow pass writes nothing per arc (dtheta only)
                        _lib.check(lib.nfst_tile_flow_f32(packed.c_struct(), lc, cond.data_ptr(), _ptr(g32), _ptr(post),
                                                          _ptr(dtheta), streams[i]))
                    else:
                        dst = post if want_post else (torch.empty(A, **f32) if have_cond else cond)
                        _lib.check(lib.nfst_sell_flow_f32(packed.c_struct(), lc, cond.data_ptr(), _ptr(g32), dst.data_ptr(),
                                                          None, None, None, _ptr(dtheta), _ptr(gfar), streams[i]))
                    launch_count += 1
                continue
            _lib.check(
                lib.nfst_bwd_fused_f32(
                    packed.c_struct(), _launch(g, st), sc, _ptr(alpha), _ptr(logz), _ptr(g32), _ptr(beta),
                    _ptr(logz_bwd), _ptr(post), _ptr(dtheta), _ptr(delta), _ptr(backptr), _ptr(vit), streams[i],
                )
            )
            launch_count += 1
        streams.join()
    del keep
    for k, v in (("beta", beta), ("logz_bwd", logz_bwd), ("post", post), ("dtheta", dtheta), ("delta", delta),
                 ("backptr", backptr), ("vit_score", vit)):
        if v is not None:
            out[k] = v
    return out


def _small_fused_fits(g: LaunchGroup, vocab: int, st: torch.dtype) -> bool:
    """The one-launch forward+backward of a small-lattice group keeps both passes' arrays in shared memory: it wins
    while all the group's lattices are resident at once (config 1, B = 32: 69 vs 84 us for two launches) and loses once
    they are not (config 5, B = 4096: 206 vs 153 us -- the two kernels each need about half the footprint, so twice as
    many lattices run per wave)."""
    foot = pack_mod.small_footprint_bytes(g.small_max_states, g.small_max_arcs, g.small_max_levels, vocab)
    if st != torch.float64:
        foot = foot * 3 // 4  # the bound assumes float64 state vectors
    return g.n * (foot + 1024) <= SMALL_RESIDENT_BYTES


def lattice_forward_backward(packed: PackedLattices, arc_scores=None, theta=None, *, want_dtheta: bool = False,
                             state_dtype="auto"):
    """(logZ[B], alpha[S], beta[S], post[A]) -- plus dtheta[V] when ``want_dtheta``.

    logZ is ``beta[start]`` (the reference's definition, ``log beta[0]``); posteriors are
    normalised with the forward logZ and agree to round-off.  Groups of small lattices run
    forward + backward in ONE launch (``nfst_fwd_bwd_small_f32``: the lattice lives in shared
    memory, alpha never leaves the SM); the others run the forward and the fused backward
    kernels back to back."""
    global launch_count
    lib = _lib.load()
    dev = packed.device
    sc, keep = _scores(packed, arc_scores, theta)
    st = resolve_state_dtype(packed, state_dtype)
    S, A, B, V = packed.n_states, packed.n_arcs, packed.n_lattices, packed.vocab
    alpha = torch.empty(S, dtype=st, device=dev)
    beta = torch.empty(S, dtype=st, device=dev)
    logz = torch.empty(B, dtype=st, device=dev)
    logz_bwd = torch.empty(B, dtype=st, device=dev)
    post = torch.empty(A, dtype=torch.float32, device=dev)
    dtheta = torch.zeros(V, dtype=torch.float32, device=dev) if want_dtheta else None
    gfar = _gamma_far(packed)
    packed.ensure_in_order()  # alpha of column-major lattices comes from the CSR forward kernel
    with torch.cuda.device(dev):
        streams = _GroupStreams(dev, len(packed.groups))
        for i, g in enumerate(packed.groups):
            lc = _launch(g, st)
            stream = streams[i]
            if g.tiles:  # as below, through the tile-stream kernels
                _lib.check(lib.nfst_tile_pull_f32(packed.c_struct(), lc, sc, beta.data_ptr(), logz_bwd.data_ptr(),
                                                  post.data_ptr(), None, None, None, stream))
                _lib.check(lib.nfst_tile_flow_f32(packed.c_struct(), lc, post.data_ptr(), None, post.data_ptr(),
                                                  _ptr(dtheta), stream))
                _lib.check(lib.nfst_fwd_f32(packed.c_struct(), _launch_csr_forward(g, st), sc, alpha.data_ptr(),
                                            logz.data_ptr(), stream))
                launch_count += 3
            elif g.sell:  # beta pass (cond written into post), then the flow pass in place; alpha by the CSR forward kernel
                _lib.check(lib.nfst_sell_pull_f32(packed.c_struct(), lc, sc, beta.data_ptr(), logz_bwd.data_ptr(),
                                                  post.data_ptr(), None, None, None, stream))
                _lib.check(lib.nfst_sell_flow_f32(packed.c_struct(), lc, post.data_ptr(), None, post.data_ptr(),
                                                  None, None, None, _ptr(dtheta), _ptr(gfar), stream))
                _lib.check(lib.nfst_fwd_f32(packed.c_struct(), _launch_csr_forward(g, st), sc, alpha.data_ptr(),
                                            logz.data_ptr(), stream))
                launch_count += 3
            elif g.small_max_arcs > 0 and _small_fused_fits(g, packed.vocab, st):
                _lib.check(lib.nfst_fwd_bwd_small_f32(packed.c_struct(), lc, sc, None, alpha.data_ptr(), logz.data_ptr(),
                                                      beta.data_ptr(), logz_bwd.data_ptr(), post.data_ptr(),
                                                      _ptr(dtheta), stream))
                launch_count += 1
            else:
                _lib.check(lib.nfst_fwd_f32(packed.c_struct(), lc, sc, alpha.data_ptr(), logz.data_ptr(), stream))
                _lib.check(lib.nfst_bwd_fused_f32(packed.c_struct(), lc, sc, alpha.data_ptr(), logz.data_ptr(), None,
                                                  beta.data_ptr(), logz_bwd.data_ptr(), post.data_ptr(), _ptr(dtheta),
                                                  None, None, None, stream))
                launch_count += 2
        streams.join()
    del keep
    if want_dtheta:
        return logz_bwd, alpha, beta, post, dtheta
    return logz_bwd, alpha, beta, post


class CapturedForwardBackward:
    """``lattice_forward_backward`` of one packed batch captured in a CUDA graph.

    The operators only enqueue kernels (no host synchronisation; buffers come from torch's
    capture pool), so a step over a FIXED lattice structure can be replayed with new scores:
    ``run(scores)`` copies them into the captured input and replays.  Worth it where a step is
    many launches -- the level-major groups issue one kernel per topological level: -22 % at
    64 x 300k-arc lattices, -15 % at 32 x 1M (tools/graph_timing.py); nothing for the two-launch
    groups.  Outputs are the graph's static tensors (overwritten by the next ``run``)."""

    def __init__(self, packed: PackedLattices, *, theta_mode: bool = False, want_dtheta: bool = False,
                 state_dtype="auto"):
        dev = packed.device
        n = packed.vocab if theta_mode else packed.n_arcs
        self._in = torch.zeros(n, dtype=torch.float32, device=dev)
        kw = dict(theta=self._in) if theta_mode else dict(arc_scores=self._in)
        run = lambda: lattice_forward_backward(packed, want_dtheta=want_dtheta, state_dtype=state_dtype, **kw)  # noqa: E731
        run()  # warm-up outside the capture: library load, shared-memory opt-in, side streams
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = run()

    def run(self, scores: torch.Tensor):
        self._in.copy_(scores, non_blocking=True)
        self.graph.replay()
        return self.outputs


class LatticeLogPartition(torch.autograd.Function):
    """logZ[B] = log sum over start->sink paths of exp(sum of arc scores).

    ``arc_scores`` ([A], canonical arc order of ``packed``) and/or ``theta`` ([V], score of
    an arc = theta[label], the reference's WFSTScorer parametrisation, scorers.py:1672-1675).
    backward: d logZ_b / d arc_scores[a] = posterior of arc a;  d / d theta[l] = sum of
    posteriors of the arcs labelled l.
    """

    @staticmethod
    def forward(ctx, arc_scores, theta, packed: PackedLattices, state_dtype="auto"):
        logz, alpha, cond = lattice_pull(packed, arc_scores, theta, state_dtype=state_dtype)
        ctx.packed = packed
        none = torch.empty(0)
        ctx.save_for_backward(alpha if alpha is not None else none, logz, cond if cond is not None else none,
                              arc_scores if arc_scores is not None else none, theta if theta is not None else none)
        ctx.has = (arc_scores is not None, theta is not None, alpha is not None, cond is not None)
        return logz

    @staticmethod
    def backward(ctx, grad_logz):
        alpha, logz, cond, a, t = ctx.saved_tensors
        has_a, has_t, has_alpha, has_cond = ctx.has
        need_a = has_a and ctx.needs_input_grad[0]
        need_t = has_t and ctx.needs_input_grad[1]
        if not (need_a or need_t):
            return None, None, None, None
        r = lattice_backward(
            ctx.packed, a if has_a else None, t if has_t else None, alpha=alpha if has_alpha else None, logz=logz,
            cond=cond if has_cond else None, grad_logz=grad_logz.contiguous(), want_beta=False, want_post=need_a,
            want_dtheta=need_t,
        )
        return (r.get("post") if need_a else None), (r.get("dtheta") if need_t else None), None, None


def lattice_log_partition(packed: PackedLattices, arc_scores=None, theta=None, state_dtype="auto") -> torch.Tensor:
    """logZ[B] with autograd (state dtype: float32, or float64 for deep lattices)."""
    return LatticeLogPartition.apply(arc_scores, theta, packed, state_dtype)


def _path_slots(packed: PackedLattices) -> torch.Tensor:
    """int32 [B+1]: every lattice's slot in the path buffer (a path has at most L_b - 1 arcs); cached."""
    cached = getattr(packed, "_path_off", None)
    if cached is None:
        cap = torch.clamp(packed.n_levels.to(torch.int64) - 1, min=0)
        off = torch.zeros(packed.n_lattices + 1, dtype=torch.int64, device=packed.device)
        torch.cumsum(cap, 0, out=off[1:])
        cached = packed._path_off = off.to(torch.int32)
    return cached


def lattice_viterbi(packed: PackedLattices, arc_scores=None, theta=None):
    """Best path per lattice under SURVEY.md section 8(c)'s rule (bit-exact, first-label
    tie-break).  Returns (score[B], path_offsets[B+1] int64, path_arcs int32 canonical arc
    ids, path_labels int32); labels run start->sink, eos included, the sink's pad loop
    excluded (cf. samplers.py:304-307)."""
    global launch_count
    lib = _lib.load()
    dev = packed.device
    sc, keep = _scores(packed, arc_scores, theta)
    S, B = packed.n_states, packed.n_lattices
    i32 = dict(dtype=torch.int32, device=dev)
    delta = torch.empty(S, dtype=torch.float32, device=dev)
    backptr = torch.empty(S, **i32)
    vit = torch.empty(B, dtype=torch.float32, device=dev)
    path_off = _path_slots(packed)
    path_buf = torch.empty(max(S, 1), **i32)  # >= the sum of the slots; avoids a host sync
    path_len = torch.empty(B, **i32)
    with torch.cuda.device(dev):
        streams = _GroupStreams(dev, len(packed.groups))
        for i, g in enumerate(packed.groups):
            _lib.check(lib.nfst_viterbi_paths_f32(packed.c_struct(), _launch(g, torch.float32), sc, delta.data_ptr(),
                                                  backptr.data_ptr(), vit.data_ptr(), path_off.data_ptr(),
                                                  path_buf.data_ptr(), path_len.data_ptr(), streams[i]))
            launch_count += 1 if g.small_max_arcs > 0 else 2
        streams.join()
        # ragged result: one cumsum, one host read of the total, one kernel
        offsets = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        torch.cumsum(path_len, 0, out=offsets[1:])
        n = int(offsets[-1])
        path_arcs = torch.empty(n, **i32)
        path_labels = torch.empty(n, **i32)
        _lib.check(lib.nfst_compact_paths(packed.c_struct(), path_off.data_ptr(), path_len.data_ptr(), path_buf.data_ptr(),
                                          offsets.data_ptr(), path_arcs.data_ptr(), path_labels.data_ptr(), _stream(dev)))
        launch_count += 1
    del keep
    return vit, offsets, path_arcs, path_labels


def lattice_viterbi_padded(packed: PackedLattices, arc_scores=None, theta=None, *, pad_label: int, skip_label: int = -1):
    """``lattice_viterbi`` with the result in the reference's sample layout and NO host synchronisation:
    ``(score[B], labels[B, T] int64 padded with pad_label, lengths[B] int32)``, ``T = max_levels - 1`` (the longest
    possible path).  A leading ``skip_label`` (the reference's samples never hold ``bos``, ``scorers.py:230-231``) is
    dropped.  What ``JointProb.forward(return_samples=True)`` returns as ``best_sample`` (``lightning.py:474-479``)."""
    global launch_count
    lib = _lib.load()
    dev = packed.device
    sc, keep = _scores(packed, arc_scores, theta)
    S, B = packed.n_states, packed.n_lattices
    i32 = dict(dtype=torch.int32, device=dev)
    delta = torch.empty(S, dtype=torch.float32, device=dev)
    backptr = torch.empty(S, **i32)
    vit = torch.empty(B, dtype=torch.float32, device=dev)
    path_off = _path_slots(packed)
    path_buf = torch.empty(max(S, 1), **i32)
    path_len = torch.empty(B, **i32)
    T = max(packed.max_levels - 1, 0)
    labels = torch.empty((B, T), dtype=torch.int64, device=dev)
    lengths = torch.empty(B, **i32)
    with torch.cuda.device(dev):
        streams = _GroupStreams(dev, len(packed.groups))
        for i, g in enumerate(packed.groups):
            _lib.check(lib.nfst_viterbi_paths_f32(packed.c_struct(), _launch(g, torch.float32), sc, delta.data_ptr(),
                                                  backptr.data_ptr(), vit.data_ptr(), path_off.data_ptr(),
                                                  path_buf.data_ptr(), path_len.data_ptr(), streams[i]))
            launch_count += 1 if g.small_max_arcs > 0 else 2
        streams.join()
        _lib.check(lib.nfst_pad_paths(packed.c_struct(), path_off.data_ptr(), path_len.data_ptr(), path_buf.data_ptr(),
                                      int(skip_label), int(pad_label), T, labels.data_ptr(), lengths.data_ptr(), _stream(dev)))
        launch_count += 1
    del keep
    return vit, labels, lengths


def _level_state_lists(packed: PackedLattices):
    """(states[S] int32 sorted by topological level, host offsets[L+1]) -- cached on `packed`."""
    cached = getattr(packed, "_level_lists", None)
    if cached is not None:
        return cached
    dev = packed.device
    lp = packed.level_ptr.to(torch.int64)  # per lattice: first state of each level, then the end
    lo = packed.level_off.to(torch.int64)
    n_lev = packed.n_levels.to(torch.int64)
    B = packed.n_lattices
    # global slot index of every (lattice, level); its size = next pointer - this pointer
    slot_lat = torch.repeat_interleave(torch.arange(B, device=dev), n_lev)
    slot_first = lo[:-1][slot_lat]
    slot_level = torch.arange(int(n_lev.sum()), device=dev) - torch.repeat_interleave(
        torch.cumsum(n_lev, 0) - n_lev, n_lev)
    begin = lp[slot_first + slot_level]
    size = lp[slot_first + slot_level + 1] - begin
    level_of_state = torch.repeat_interleave(slot_level, size)  # states are numbered level by level
    order = torch.argsort(level_of_state, stable=True).to(torch.int32)
    counts = torch.bincount(level_of_state, minlength=max(packed.max_levels, 1))
    off = [0] + torch.cumsum(counts, 0).cpu().tolist()
    packed._level_lists = (order, off)
    return packed._level_lists


def lattice_beta_hat(packed: PackedLattices, label_proj: torch.Tensor, Wh: torch.Tensor, W: torch.Tensor
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's full beta recurrence (``scorers.py:692-751``, Wh != 0):
    ``m_hat = tanh(label_proj[label] + Wh beta_hat[dst])``, ``m = exp(W.m_hat) beta[dst]``,
    ``beta[c] = sum m``, ``beta_hat[c] = sum (m/beta[c]) m_hat``.  ``label_proj[V, H]`` is
    ``Wx e_l + b`` for every label.  Returns (log beta[S] float32 in packed state order,
    beta_hat[S, H]).  One launch per topological level, deepest level first."""
    global launch_count
    lib = _lib.load()
    dev = packed.device
    if dev.type != "cuda":
        raise RuntimeError("nfst_b200 kernels need the packed lattices on a CUDA device (no CPU fallback)")
    V, H = label_proj.shape
    if V != packed.vocab or tuple(Wh.shape) != (H, H) or W.numel() != H:
        raise ValueError("label_proj must be [vocab, H], Wh [H, H], W [H]")
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
    proj, wh_t, w = f32(label_proj), f32(Wh.t()), f32(W.reshape(-1))
    S = packed.n_states
    log_beta = torch.empty(S, dtype=torch.float32, device=dev)
    beta_hat = torch.empty(S, H, dtype=torch.float32, device=dev)
    h_proj = torch.empty(S, H, dtype=torch.float32, device=dev)
    order, off = _level_state_lists(packed)
    with torch.cuda.device(dev):
        stream = _stream(dev)
        for lvl in range(len(off) - 2, -1, -1):
            n = off[lvl + 1] - off[lvl]
            if n == 0:
                continue
            _lib.check(lib.nfst_beta_hat_level_f32(packed.c_struct(), order.data_ptr() + 4 * off[lvl], n, H,
                                                   proj.data_ptr(), wh_t.data_ptr(), w.data_ptr(), log_beta.data_ptr(),
                                                   beta_hat.data_ptr(), h_proj.data_ptr(), stream))
            launch_count += 1
    return log_beta, beta_hat


def beta_dense(packed: PackedLattices, beta: torch.Tensor, k: int = 1, dense_states: Optional[int] = None) -> torch.Tensor:
    """Real-space beta in the reference's layout ``[B*k, S]`` (``scorers.py:854``): row
    b*k+j is lattice b, column = original state id; trimmed states hold 0."""
    global launch_count
    lib = _lib.load()
    dev = packed.device
    if dense_states is None:
        if packed.dense_shape is None:
            raise ValueError("dense_states is required for lattices that were not packed from dense tables")
        dense_states = packed.dense_shape[1]
    out = torch.zeros(packed.n_lattices * k, dense_states, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nfst_beta_to_dense(packed.c_struct(), beta.data_ptr(), int(beta.dtype == torch.float64),
                                          packed.orig_state.data_ptr(), k, dense_states, out.data_ptr(), _stream(dev)))
        launch_count += 1
    return out


def compute_beta(emission: torch.Tensor, transition: torch.Tensor, theta: torch.Tensor, k: int = 1,
                 packed: Optional[PackedLattices] = None) -> torch.Tensor:
    """Drop-in for ``FSAGRUScorer.compute_beta()`` in the Wh = 0 regime
    (``scorers.py:753-875``): real-space ``beta[B*k, S]`` fp32 for the tables that
    ``set_masks`` / ``set_k`` would hold, with arc weight ``exp(theta[label])``."""
    if packed is None:
        packed = pack_dense(emission, transition, weighted=False)
    r = lattice_backward(packed, None, theta, want_beta=True)
    return beta_dense(packed, r["beta"], k)
