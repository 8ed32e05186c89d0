"""Lattice construction on the device: the transliteration edit lattices, straight from id strings.

The reference builds every training lattice offline with OpenFst (``src/preprocess/tr.py:142-190``): ``x o T o y``
for the one-state edit machine ``T`` of ``src/fsm/tr.py:321-390``, each arc replaced by the chain of its marks
(``OwnAST.mfst_weight_projection``, ``src/modules/path_semiring.py:120-180``), ``bos`` in front and ``eos`` behind
(``Preprocess.composed_to_matrices``, ``src/preprocess/preprocess.py:51-174``), and stores it as dense ``[S, V]``
tables.  For that machine the composition is the ``(|x|+1) x (|y|+1)`` edit grid, so the mark lattice can be written
down directly: ``edit_lattices`` emits its arc list with one kernel (``nfst_edit_lattice_arcs``) and packs it with
the device packer (``nfst_pack_small``) -- four launches from strings to a batch the DP kernels run on, no dense
table, no OpenFst.  (The reference also runs pynini's ``optimize()``; that renumbers and may merge states but keeps
the set of mark strings, which is all the dynamic programme sees.)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .pack import PackedLattices, _excl_cumsum, pack_arcs, pack_small_device


def _padded_ids(seqs, device):
    if isinstance(seqs, torch.Tensor):
        raise TypeError("pass (ids[B, L], lengths[B]) tensors or a list of id sequences")
    n = torch.tensor([len(s) for s in seqs], dtype=torch.int32)
    ids = torch.zeros((len(seqs), max(int(n.max()), 1)), dtype=torch.int32)
    for i, s in enumerate(seqs):
        ids[i, : len(s)] = torch.as_tensor(list(s), dtype=torch.int32)
    return ids.to(device), n.to(device)


def edit_lattices(x, y, *, vocab: int, bos: int, eos: int, input_mark: int, output_mark: int, sub_mark: Optional[int] = None,
                  x_len: Optional[torch.Tensor] = None, y_len: Optional[torch.Tensor] = None, device="cuda") -> PackedLattices:
    """Packed mark lattices of ``x[b] o T o y[b]`` for the edit machine of ``src/fsm/tr.py:321-390``.

    ``x`` / ``y``: lists of id sequences, or int tensors ``[B, L]`` with ``x_len`` / ``y_len``.  Marks: a deletion of
    ``x_i`` is the chain ``[input_mark, x_i]``, an insertion of ``y_j`` ``[output_mark, y_j]``, a substitution
    (``sub_mark`` given: the reference's ``add_sub`` machine) ``[sub_mark, input_mark, x_i, output_mark, y_j]``.
    ``arc_origin`` of the result indexes the constructed arc list (start state 0, grid row-major, chain states,
    sink; see ``include/nfst_b200.h``)."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("edit_lattices builds on a CUDA device (nfst_b200 has no CPU fallback)")
    if not isinstance(x, torch.Tensor):
        x, x_len = _padded_ids(x, dev)
    if not isinstance(y, torch.Tensor):
        y, y_len = _padded_ids(y, dev)
    x, y = x.to(device=dev, dtype=torch.int32).contiguous(), y.to(device=dev, dtype=torch.int32).contiguous()
    x_len, y_len = x_len.to(device=dev, dtype=torch.int32).contiguous(), y_len.to(device=dev, dtype=torch.int32).contiguous()
    B = int(x.shape[0])
    if y.shape[0] != B or x_len.numel() != B or y_len.numel() != B:
        raise ValueError("x, y, x_len, y_len must describe the same number of pairs")
    add_sub = sub_mark is not None
    n, m = x_len.to(torch.int64), y_len.to(torch.int64)
    c = 3 if add_sub else 2
    sub = 4 * n * m if add_sub else torch.zeros_like(n)
    n_states = 1 + (n + 1) * (m + 1) + n * (m + 1) + (n + 1) * m + sub + 1
    n_arcs = 1 + n * (c * m + 1) + m + 1 + n * (m + 1) + (n + 1) * m + sub
    state_off, arc_off = _excl_cumsum(n_states), _excl_cumsum(n_arcs)
    head = torch.stack([arc_off[-1], state_off[-1], n_states.max(), n_arcs.max(), (x_len > x.shape[1]).any().to(torch.int64),
                        (y_len > y.shape[1]).any().to(torch.int64)]).cpu().tolist()  # the sizes: one host read
    if head[4] or head[5]:
        raise ValueError("a length exceeds its id tensor")
    A0, S0, smax, amax = head[:4]
    if A0 >= 2**31 or S0 >= 2**31:
        raise ValueError("batch too large for int32 indices; shard it")
    src = torch.empty(A0, dtype=torch.int32, device=dev)
    dst = torch.empty(A0, dtype=torch.int32, device=dev)
    lab = torch.empty(A0, dtype=torch.int32, device=dev)
    arc_off32, state_off32 = arc_off.to(torch.int32), state_off.to(torch.int32)
    with torch.cuda.device(dev):
        _lib.check(lib.nfst_edit_lattice_arcs(B, x.data_ptr(), x_len.data_ptr(), int(x.shape[1]), y.data_ptr(), y_len.data_ptr(),
                                              int(y.shape[1]), int(bos), int(eos), int(input_mark), int(output_mark),
                                              int(sub_mark) if add_sub else -1, int(add_sub), arc_off32.data_ptr(), src.data_ptr(),
                                              dst.data_ptr(), lab.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    packed = pack_small_device(state_off32, arc_off32, src, dst, lab, vocab, src_is_global=False, max_states=int(smax),
                               max_arcs=int(amax), n_states_raw=int(S0))
    if packed is None:  # pairs too long for one SM's shared memory: the general packer
        lat = torch.repeat_interleave(torch.arange(B, device=dev), n_arcs)
        packed = pack_arcs(lat, src.to(torch.int64), dst.to(torch.int64), lab.to(torch.int64), n_states, vocab)
    return packed


def edit_lattice_size(n: int, m: int, add_sub: bool = True):
    """(states, arcs) of the lattice of one pair with ``|x| = n``, ``|y| = m``."""
    s, a = C.c_int64(), C.c_int64()
    _lib.load().nfst_edit_lattice_size(int(n), int(m), int(add_sub), C.byref(s), C.byref(a))
    return s.value, a.value
