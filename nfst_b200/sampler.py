"""Lattice side of the reference's sampling loop, one fused kernel per time step.

``Sampler.stateful_sample`` (``src/modules/samplers.py:182-335``) walks the lattice one symbol per step:
score (network) -> beta look-ahead + masks (``scorers.py:577-593``, ``:340-357``, ``:1037-1054``) ->
``Categorical`` sample / ``log_prob`` (``samplers.py:251-283``) -> ``update_fsa_state`` (``scorers.py:683-690``).
Everything but the network is a function of the lattice and runs here in ONE launch over the CSR arcs of each
row's current state (``nfst_walk_step_f32``); the k-fold expansion of the dense tables that ``set_k``
materialises (``scorers.py:887-918``) never happens -- row r simply belongs to lattice r // k.

The network (embedding, GRU cell, ``beta_scorer``) stays with the caller: it hands ``prefix[N, V]`` in.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from . import ops
from .pack import PackedLattices


class LatticeWalker:
    """State of a batch of walks: ``k`` rows per lattice, all starting at the start state.

    ``beta_real``: REAL-space beta per packed state, float32 ``[S]`` (e.g. ``exp`` of ``lattice_backward``'s
    beta, or ``lattice_beta_hat``'s) -- the reference adds beta itself, not its log, to the logits
    (``scorers.py:590-592``).  ``faithful=True`` reproduces the reference's one-step-stale look-ahead
    (quirk Q9: ``scorers.py:584`` reads the state that ``:679`` has not advanced yet); ``False`` looks ahead
    from the state the symbol is actually drawn at.
    """

    def __init__(self, packed: PackedLattices, k: int, beta_real: torch.Tensor, pad_id: int, *,
                 temperature: float = 1.0, faithful: bool = True):
        if packed.device.type != "cuda":
            raise RuntimeError("nfst_b200 kernels need the packed lattices on a CUDA device (no CPU fallback)")
        if beta_real.numel() != packed.n_states:
            raise ValueError("beta_real must hold one value per packed state")
        self.packed, self.k, self.pad_id = packed, int(k), int(pad_id)
        self.temperature, self.faithful = float(temperature), bool(faithful)
        self.beta_real = beta_real.detach().to(device=packed.device, dtype=torch.float32).contiguous()
        self.n_rows = packed.n_lattices * self.k
        start = packed.start_state.to(torch.int32).repeat_interleave(self.k).contiguous()
        self.state = start  # where the next symbol is drawn
        self.look_state = start.clone()  # one step behind (the reference's look-ahead state)
        self._first = True

    def dense_state(self, state: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Row states as the reference numbers them (dense-table row ids)."""
        s = self.state if state is None else state
        return self.packed.orig_state[s.long()].to(torch.int64)

    def consume(self, symbols: torch.Tensor) -> None:
        """Advance every row over a symbol WITHOUT scoring it -- the reference feeds ``bos`` as the first input
        (``scorers.py:230-231``), so its walk starts one arc in."""
        z = torch.zeros(self.n_rows, self.packed.vocab, dtype=torch.float32, device=self.packed.device)
        _, _, nxt, _ = walk_step(self.packed, self.k, self.state, z, self.beta_real, self.pad_id,
                                 symbols=symbols, look_state=None)
        self.look_state, self.state = self.state, nxt

    def step(self, prefix: torch.Tensor, *, base_mask: Optional[torch.Tensor] = None,
             symbols: Optional[torch.Tensor] = None, uniform: Optional[torch.Tensor] = None
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One time step: returns (symbol[N], log_prob[N], logsumexp[N]) and advances the rows.  ``symbols``
        scores given symbols (``evaluate_only``); otherwise one is sampled per row (``uniform[N]`` in [0, 1),
        drawn here when not given)."""
        if symbols is None and uniform is None:
            uniform = torch.rand(self.n_rows, device=self.packed.device)
        sym, logp, nxt, logz = walk_step(
            self.packed, self.k, self.state, prefix, self.beta_real, self.pad_id, base_mask=base_mask,
            symbols=symbols, uniform=uniform, temperature=self.temperature,
            look_state=self.look_state if self.faithful else None)
        self.look_state, self.state = self.state, nxt
        return sym, logp, logz


def walk_step(packed: PackedLattices, k: int, state: torch.Tensor, prefix: torch.Tensor, beta_real: torch.Tensor,
              pad_id: int, *, base_mask: Optional[torch.Tensor] = None, symbols: Optional[torch.Tensor] = None,
              uniform: Optional[torch.Tensor] = None, temperature: float = 1.0,
              look_state: Optional[torch.Tensor] = None):
    """Functional form of one step (see ``nfst_walk_step_f32`` in include/nfst_b200.h).
    Returns (symbol int32 [N], log_prob float32 [N], next_state int32 [N] packed, logsumexp float32 [N])."""
    lib = _lib.load()
    dev = packed.device
    if dev.type != "cuda":
        raise RuntimeError("nfst_b200 kernels need the packed lattices on a CUDA device (no CPU fallback)")
    N = packed.n_lattices * k
    V = packed.vocab
    if state.numel() != N or tuple(prefix.shape) != (N, V):
        raise ValueError(f"state must be [{N}] and prefix [{N}, {V}]")
    if (symbols is None) == (uniform is None):
        raise ValueError("give exactly one of symbols (score) and uniform (sample)")
    i32 = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.int32).contiguous()  # noqa: E731
    f32 = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
    st, lk, pre, bm, be, sy, un = i32(state), i32(look_state), f32(prefix), f32(base_mask), f32(beta_real), i32(symbols), f32(uniform)
    if bm is not None and tuple(bm.shape) != (N, V):
        raise ValueError(f"base_mask must be [{N}, {V}]")
    sym = torch.empty(N, dtype=torch.int32, device=dev)
    nxt = torch.empty(N, dtype=torch.int32, device=dev)
    logp = torch.empty(N, dtype=torch.float32, device=dev)
    logz = torch.empty(N, dtype=torch.float32, device=dev)
    static = packed.static_scores
    p = ops._ptr
    with torch.cuda.device(dev):
        _lib.check(lib.nfst_walk_step_f32(packed.c_struct(), N, k, p(st), p(lk), p(pre), p(bm), p(be), p(static),
                                          float(temperature), int(pad_id), p(sy), p(un), p(sym), p(logp), p(nxt),
                                          p(logz), ops._stream(dev)))
        ops.launch_count += 1
    return sym, logp, nxt, logz


def sample_paths(packed: PackedLattices, k: int, arc_scores=None, theta=None, *, pad_id: int = 3,
                 uniform: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None):
    """k exact posterior samples per lattice for an arc-factored model (``WFSTScorer``, ``scorers.py:1663-1687``):
    the whole loop of ``Sampler.stateful_sample`` in one launch (``nfst_sample_paths_f32``).

    Returns ``(labels int32 [B*k, T] padded with pad_id, lengths int32 [B*k], log_q float32 [B*k],
    arcs int32 [B*k, T] canonical arc ids, logz [B])``.  ``log_q = score(path) - logZ``: with these samples
    every importance weight of ``Estimators.iwae`` (``estimatros.py:10-44``) is ``logZ`` itself."""
    lib = _lib.load()
    dev = packed.device
    if dev.type != "cuda":
        raise RuntimeError("nfst_b200 kernels need the packed lattices on a CUDA device (no CPU fallback)")
    sc, keep = ops._scores(packed, arc_scores, theta)
    if beta is None:
        r = ops.lattice_backward(packed, arc_scores, theta, want_beta=True)
        beta, logz = r["beta"], r["logz_bwd"]
    else:
        logz = beta[packed.start_state.long()]
    N, T = packed.n_lattices * k, max(packed.max_levels - 1, 1)
    if uniform is None:
        uniform = torch.rand(N, T, device=dev)
    uniform = uniform.detach().to(device=dev, dtype=torch.float32).contiguous()
    if tuple(uniform.shape) != (N, T):
        raise ValueError(f"uniform must be [{N}, {T}]")
    labels = torch.empty(N, T, dtype=torch.int32, device=dev)
    arcs = torch.empty(N, T, dtype=torch.int32, device=dev)
    length = torch.empty(N, dtype=torch.int32, device=dev)
    log_q = torch.empty(N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nfst_sample_paths_f32(packed.c_struct(), N, k, T, sc, beta.data_ptr(), int(beta.dtype == torch.float64),
                                             uniform.data_ptr(), int(pad_id), labels.data_ptr(), arcs.data_ptr(),
                                             length.data_ptr(), log_q.data_ptr(), ops._stream(dev)))
        ops.launch_count += 1
    del keep
    return labels, length, log_q, arcs, logz


def stripping_pad(sequences: torch.Tensor, pad_id: int) -> torch.Tensor:
    """``Sampler.stripping_pad`` (``src/modules/samplers.py:162-180``): every row of ``sequences[N, T]`` (int64)
    left-compacted, symbol id 0 dropped, padded with ``pad_id``; the result is as wide as the reference's (it stops
    at the first column that is ``pad`` in every row).  Two launches instead of T dependent scatter steps."""
    if sequences.dim() != 2:
        raise AssertionError("sequences must be [N, T]")  # samplers.py:163
    dev = sequences.device
    if dev.type != "cuda":
        raise RuntimeError("nfst_b200 kernels need CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    seq = sequences.detach().to(torch.int64).contiguous()
    N, T = seq.shape
    if N == 0 or T == 0:
        return seq.clone()
    out = torch.empty_like(seq)
    flags = torch.empty(T, dtype=torch.int32, device=dev)
    width = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nfst_strip_pad(seq.data_ptr(), N, T, int(pad_id), out.data_ptr(), flags.data_ptr(), width.data_ptr(),
                                      ops._stream(dev)))
        ops.launch_count += 2
    return out[:, : int(width.item())].contiguous()
