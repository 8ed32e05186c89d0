"""ctypes binding of the C ABI declared in include/nfst_b200.h.

There is no CPU fallback: if the shared library is missing this raises, and every entry
point that does compute requires CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)


class PackedLatticesC(C.Structure):
    """Mirror of ``nfst_packed_lattices_t``."""

    _fields_ = [
        ("n_lattices", C.c_int32),
        ("n_states", C.c_int32),
        ("n_arcs", C.c_int32),
        ("vocab", C.c_int32),
        ("state_off", C.c_void_p),
        ("level_off", C.c_void_p),
        ("level_ptr", C.c_void_p),
        ("start_state", C.c_void_p),
        ("sink_off", C.c_void_p),
        ("sinks", C.c_void_p),
        ("in_ptr", C.c_void_p),
        ("src_in", C.c_void_p),
        ("label_in", C.c_void_p),
        ("in2out", C.c_void_p),
        ("out_ptr", C.c_void_p),
        ("dst_out", C.c_void_p),
        ("label_out", C.c_void_p),
        ("lanes_in_log2", C.c_void_p),
        ("lanes_out_log2", C.c_void_p),
        ("fwd_chunk_off", C.c_void_p),
        ("fwd_chunks", C.c_void_p),
        ("bwd_chunk_off", C.c_void_p),
        ("bwd_chunks", C.c_void_p),
        ("fwd_gather", C.c_void_p),
        ("bwd_order", C.c_void_p),
        ("out_deg8", C.c_void_p),
        ("sell_desc", C.c_void_p),
        ("sell_lvl_slice", C.c_void_p),
        ("tile_stream", C.c_void_p),
        ("tile_tab", C.c_void_p),
        ("tile_lw_off", C.c_void_p),
        ("tile_lat_info", C.c_void_p),
        ("out_arc", C.c_void_p),
    ]


class LaunchC(C.Structure):
    """Mirror of ``nfst_launch_t``."""

    _fields_ = [
        ("lattice_ids", C.c_void_p),
        ("n_ids", C.c_int32),
        ("block_threads", C.c_int32),
        ("window_states", C.c_int32),
        ("state_f64", C.c_int32),
        ("chunk_cap", C.c_int32),
        ("n_levels", C.c_int32),
        ("fwd_level_chunks", C.c_void_p),
        ("fwd_level_off", C.c_void_p),
        ("bwd_level_chunks", C.c_void_p),
        ("bwd_level_off", C.c_void_p),
        ("bwd_level_lat", C.c_void_p),
        ("small_max_states", C.c_int32),
        ("small_max_arcs", C.c_int32),
        ("small_max_levels", C.c_int32),
        ("sell", C.c_int32),
        ("sell_far", C.c_int32),
        ("tiles", C.c_int32),
        ("tile_ring", C.c_int32),
        ("tile_far", C.c_int32),
        ("tile_cap_arcs", C.c_int32),
        ("tile_cap_bytes", C.c_int32),
        ("tile_stages", C.c_int32),
        ("tile_flow_bits", C.c_int32),
    ]


class PackOutC(C.Structure):
    """Mirror of ``nfst_pack_out_t``."""

    _fields_ = [(n, C.c_void_p) for n in (
        "state_off", "arc_off", "level_off", "sink_off", "n_levels", "start_state", "level_ptr", "sinks", "orig_state",
        "in_ptr", "out_ptr", "src_in", "label_in", "in2out", "dst_out", "label_out", "src_out", "arc_origin", "out_deg8",
        "lattice_stats", "totals")]


class ScoresC(C.Structure):
    """Mirror of ``nfst_scores_t``."""

    _fields_ = [("arc_scores", C.c_void_p), ("theta", C.c_void_p)]


# every symbol include/nfst_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "nfst_abi_version": (C.c_int, []),
    "nfst_last_error_string": (C.c_char_p, []),
    "nfst_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "nfst_launch_smem_bytes": (C.c_size_t, [C.POINTER(LaunchC), C.c_int32] + [C.c_int] * 7),
    "nfst_fwd_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.POINTER(LaunchC), C.POINTER(ScoresC), _P, _P, _P]),
    "nfst_bwd_fused_f32": (
        C.c_int,
        [C.POINTER(PackedLatticesC), C.POINTER(LaunchC), C.POINTER(ScoresC), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    ),
    "nfst_fwd_bwd_small_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.POINTER(LaunchC), C.POINTER(ScoresC)] + [_P] * 8),
    "nfst_viterbi_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.POINTER(LaunchC), C.POINTER(ScoresC), _P, _P, _P, _P]),
    "nfst_backtrace": (C.c_int, [C.POINTER(PackedLatticesC), _P, _P, _P, _P, _P]),
    "nfst_viterbi_paths_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.POINTER(LaunchC), C.POINTER(ScoresC), _P, _P, _P, _P, _P, _P, _P]),
    "nfst_compact_paths": (C.c_int, [C.POINTER(PackedLatticesC), _P, _P, _P, _P, _P, _P, _P]),
    "nfst_pad_paths": (C.c_int, [C.POINTER(PackedLatticesC), _P, _P, _P, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P]),
    "nfst_sell_smem_bytes": (C.c_size_t, [C.POINTER(LaunchC), C.c_int32, C.c_int, C.c_int, C.c_int]),
    "nfst_sell_pull_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.POINTER(LaunchC), C.POINTER(ScoresC)] + [_P] * 7),
    "nfst_sell_flow_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.POINTER(LaunchC)] + [_P] * 9),
    "nfst_tile_smem_bytes": (C.c_size_t, [C.POINTER(LaunchC), C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int]),
    "nfst_tile_pull_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.POINTER(LaunchC), C.POINTER(ScoresC)] + [_P] * 7),
    "nfst_tile_flow_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.POINTER(LaunchC)] + [_P] * 5),
    "nfst_tile_debug_read": (C.c_int, [_P]),
    "nfst_walk_step_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_float, C.c_int32,
                                     _P, _P, _P, _P, _P, _P, _P]),
    "nfst_strip_pad": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, _P]),
    "nfst_sample_paths_f32": (C.c_int, [C.POINTER(PackedLatticesC), C.c_int32, C.c_int32, C.c_int32, C.POINTER(ScoresC), _P,
                                        C.c_int, _P, C.c_int32, _P, _P, _P, _P, _P]),
    "nfst_beta_hat_level_f32": (C.c_int, [C.POINTER(PackedLatticesC), _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "nfst_beta_to_dense": (C.c_int, [C.POINTER(PackedLatticesC), _P, C.c_int, _P, C.c_int32, C.c_int32, _P, _P]),
    "nfst_pack_small_smem_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "nfst_pack_small_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "nfst_pack_small": (C.c_int, [C.c_int32, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(PackOutC), _P,
                                  C.c_size_t, C.c_int64, C.c_int64, C.c_int32, _P]),
    "nfst_edit_lattice_size": (None, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "nfst_edit_lattice_arcs": (C.c_int, [C.c_int32, _P, _P, C.c_int32, _P, _P, C.c_int32] + [C.c_int32] * 6 + [_P, _P, _P, _P, _P]),
    "nfst_level_sweeps": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int32, _P]),
    "nfst_dense_count_arcs": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    "nfst_dense_extract_arcs": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int64, _P, _P, _P, _P]),
    "nfst_pack_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64]),
    "nfst_pack_dense": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(PackOutC), _P, C.c_size_t,
                                  C.c_int32, _P]),
}

_lib = None


def library_path() -> str:
    return os.environ.get("NFST_LIB", _build.LIB)  # NFST_LIB: load a variant build (timing hooks)


def load() -> C.CDLL:
    """Load libnfst_b200.so (built in-tree by ``nfst_b200.build``); raise if absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m nfst_b200.build` (or __graft_entry__.build()). "
            "nfst_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.nfst_abi_version() != 18:
        raise RuntimeError("libnfst_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().nfst_last_error_string()
        raise RuntimeError(f"nfst_b200 error {rc}: {msg.decode() if msg else ''}")
