"""Build + ctypes binding of oracle/lattice_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__ (build + smoke check) and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lattice_oracle.c")
LIB = os.path.join(HERE, "_build", "liblattice_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + res.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        p = C.c_void_p
        lib.oracle_batch_forward_backward.restype = C.c_int
        lib.oracle_batch_forward_backward.argtypes = [C.c_int, p, p, p, p, p, C.c_int, C.c_int, p, p, p, p, C.c_int]
        lib.oracle_batch_viterbi_f32.restype = C.c_int
        lib.oracle_batch_viterbi_f32.argtypes = [C.c_int, p, p, p, p, p, C.c_int, p, p, p, C.c_int]
        lib.oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(load().oracle_max_threads())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _scan_order(arc_lattice, src, label):
    """Sort arcs by (lattice, source state, label) -- the reference's dense-table scan
    order (scorers.py:705-706) -- stably."""
    return np.lexsort((label, src, arc_lattice))


class Batch:
    """Arc list of a batch in scan order, with the permutation back to the input order."""

    def __init__(self, arc_lattice, src, dst, label, scores, n_states):
        arc_lattice = np.asarray(arc_lattice, dtype=np.int64)
        self.order = _scan_order(arc_lattice, np.asarray(src), np.asarray(label))
        self.src = np.ascontiguousarray(np.asarray(src)[self.order], dtype=np.int32)
        self.dst = np.ascontiguousarray(np.asarray(dst)[self.order], dtype=np.int32)
        self.label = np.ascontiguousarray(np.asarray(label)[self.order], dtype=np.int32)
        self.w = np.ascontiguousarray(np.asarray(scores)[self.order], dtype=np.float32)
        n_states = np.asarray(n_states, dtype=np.int64)
        self.B = len(n_states)
        self.state_off = np.concatenate([[0], np.cumsum(n_states)]).astype(np.int64)
        self.arc_off = np.concatenate([[0], np.cumsum(np.bincount(arc_lattice, minlength=self.B))]).astype(np.int64)
        self.n_arcs = len(self.src)

    def unsort(self, x):
        out = np.empty_like(x)
        out[self.order] = x
        return out


def forward_backward(batch: Batch, start: int = 0, want_post: bool = True, n_threads: int = 0, want_states: bool = True):
    """(logZ[B], alpha[S0], beta[S0], post[A] in the caller's arc order), float64; original
    state numbering (global = state_off[b] + local id)."""
    lib = load()
    S0 = int(batch.state_off[-1])
    logz = np.zeros(batch.B)
    alpha = np.zeros(S0) if (want_post and want_states) else None
    beta = np.zeros(S0) if want_states else None
    post = np.zeros(batch.n_arcs) if want_post else None
    bad = lib.oracle_batch_forward_backward(batch.B, _ptr(batch.state_off), _ptr(batch.arc_off), _ptr(batch.src),
                                            _ptr(batch.dst), _ptr(batch.w), start, int(want_post), _ptr(alpha),
                                            _ptr(beta), _ptr(post), _ptr(logz), n_threads)
    if bad:
        raise ValueError(f"{bad} cyclic lattice(s)")
    return logz, alpha, beta, (batch.unsort(post) if want_post else None)


def viterbi(batch: Batch, start: int = 0, n_threads: int = 0):
    """(score[B] float32, list of per-lattice arrays of arc indices in the caller's arc
    order, list of per-lattice label arrays)."""
    lib = load()
    S0 = int(batch.state_off[-1])
    score = np.zeros(batch.B, dtype=np.float32)
    path = np.zeros(max(S0, 1), dtype=np.int32)
    plen = np.zeros(batch.B, dtype=np.int32)
    lib.oracle_batch_viterbi_f32(batch.B, _ptr(batch.state_off), _ptr(batch.arc_off), _ptr(batch.src), _ptr(batch.dst),
                                 _ptr(batch.w), start, _ptr(score), _ptr(path), _ptr(plen), n_threads)
    paths, labels = [], []
    for b in range(batch.B):
        loc = path[batch.state_off[b] : batch.state_off[b] + plen[b]].astype(np.int64) + batch.arc_off[b]
        paths.append(batch.order[loc])
        labels.append(batch.label[loc].astype(np.int64))
    return score, paths, labels
