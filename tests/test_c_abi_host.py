"""The C boundary without a GPU: ``include/nfst_b200.h`` against the shared library and the ctypes mirror.

No compute entry is called here (they need a device).  What is checked: the header is plain C (gcc compiles it,
pedantic, as C99 and as C++), every function it declares is exported by ``libnfst_b200.so`` and bound in
``nfst_b200._lib.SYMBOLS`` with the same number of arguments, the ctypes structures have the sizes and field offsets
the C compiler gives the header's structs, and the host-only entries (ABI version, size queries, the edit-lattice
size formula) answer without touching CUDA.
"""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from nfst_b200 import _lib
from nfst_b200 import build as nb_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nfst_b200.h")


def _header_text() -> str:
    with open(HEADER) as f:
        src = f.read()
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)  # comments name entry points too: drop them


def _declarations():
    """{function name: number of parameters} of every prototype in the header."""
    out = {}
    for m in re.finditer(r"\b(nfst_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", _header_text(), flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def _struct_fields(struct_tag: str):
    """field names of ``typedef struct <tag> { ... }`` in declaration order (arrays and multi-declarators included)"""
    m = re.search(r"typedef\s+struct\s+" + struct_tag + r"\s*\{(.*?)\}\s*\w+\s*;", _header_text(), flags=re.S)
    assert m, struct_tag
    names = []
    for decl in m.group(1).split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for piece in decl.split(","):
            names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", piece)[-1])
    return names


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(nb_build.LIB):
        nb_build.build()
    return _lib.load()


def test_header_functions_are_exported_and_bound(lib):
    decl = _declarations()
    assert len(decl) >= 30, decl
    assert set(decl) == set(_lib.SYMBOLS), (set(decl) ^ set(_lib.SYMBOLS))
    for name, n_params in decl.items():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
        _, argtypes = _lib.SYMBOLS[name]
        assert len(argtypes) == n_params, f"{name}: header has {n_params} parameters, the ctypes binding {len(argtypes)}"
    # nothing but the header's names leaves the library
    nm = subprocess.run(["nm", "-D", "--defined-only", nb_build.LIB], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in nm.splitlines() if " T " in ln and ln.split()[-1].startswith("nfst_")}
    assert exported == set(decl), exported ^ set(decl)


def test_abi_version_and_error_string(lib):
    m = re.search(r"#define\s+NFST_ABI_VERSION\s+(\d+)", _header_text())
    assert lib.nfst_abi_version() == int(m.group(1))
    s = lib.nfst_last_error_string()
    assert s is not None and isinstance(s, bytes)


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    mirrors = {
        "nfst_packed_lattices": ("nfst_packed_lattices_t", _lib.PackedLatticesC),
        "nfst_launch": ("nfst_launch_t", _lib.LaunchC),
        "nfst_scores": ("nfst_scores_t", _lib.ScoresC),
        "nfst_pack_out": ("nfst_pack_out_t", _lib.PackOutC),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nfst_b200.h"', "int main(void) {"]
    for tag, (tname, mirror) in mirrors.items():
        fields = _struct_fields(tag)
        assert fields == [f[0] for f in mirror._fields_], (tag, fields, [f[0] for f in mirror._fields_])
        lines.append(f'  printf("{tname} %zu\\n", sizeof({tname}));')
        for f in fields:
            lines.append(f'  printf("{tname}.{f} %zu\\n", offsetof({tname}, {f}));')
    lines.append('  printf("nfst_chunk_t %zu\\n", sizeof(nfst_chunk_t));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for tag, (tname, mirror) in mirrors.items():
        assert int(got[tname]) == C.sizeof(mirror), tname
        for f, _ in mirror._fields_:
            assert int(got[f"{tname}.{f}"]) == getattr(mirror, f).offset, (tname, f)
    assert int(got["nfst_chunk_t"]) == 16
    # the same header under a C++ compiler (the reference-side binding may be either)
    cpp = tmp_path / "inc.cpp"
    cpp.write_text('#include "nfst_b200.h"\nint main() { return NFST_ABI_VERSION > 0 ? 0 : 1; }\n')
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-fsyntax-only", str(cpp)], check=True)


def test_edit_lattice_size_formula_matches_the_host_construction(lib):
    """nfst_edit_lattice_size is host arithmetic: it must count what the numpy construction the tests use as the
    construction oracle emits (tests/test_edit_lattice_oracle.py pins that one to the alignments)."""
    from oracle import edit_lattice_oracle as elo

    rng = np.random.default_rng(5)
    for add_sub in (0, 1):
        for n, m in [(1, 1), (1, 4), (5, 1), (3, 3), (7, 5), (2, 9)]:
            ns, na = C.c_int64(), C.c_int64()
            lib.nfst_edit_lattice_size(n, m, add_sub, C.byref(ns), C.byref(na))
            x = [int(v) for v in rng.integers(5, 9, size=n)]
            y = [int(v) for v in rng.integers(9, 14, size=m)]
            arcs, n_states = elo.edit_lattice(x, y, bos=1, eos=2, input_mark=3, output_mark=4,
                                              sub_mark=14 if add_sub else None)
            assert ns.value == n_states, (n, m, add_sub, ns.value, n_states)
            assert na.value == len(arcs), (n, m, add_sub, na.value, len(arcs))


def test_size_queries_answer_on_the_host(lib):
    # workspace sizes are multiples of 16 bytes (carved into aligned arrays by the caller) and grow with the problem
    a = lib.nfst_pack_workspace_bytes(4, 16, 200)
    b = lib.nfst_pack_workspace_bytes(8, 16, 400)
    assert 0 < a < b
    s1 = lib.nfst_pack_small_workspace_bytes(100, 300)
    s2 = lib.nfst_pack_small_workspace_bytes(1000, 3000)
    assert 0 < s1 < s2
    m1 = lib.nfst_pack_small_smem_bytes(64, 256)
    m2 = lib.nfst_pack_small_smem_bytes(512, 2048)
    assert 0 < m1 < m2 <= 227 * 1024


def test_device_entry_fails_with_a_status_not_a_crash_when_there_is_no_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the no-device answer cannot be provoked")
    sm = C.c_int()
    rc = lib.nfst_device_info(0, C.byref(sm), None, None, None)
    assert rc in (-2, -3), rc  # NFST_ERR_CUDA / NFST_ERR_UNSUPPORTED_DEVICE
    assert lib.nfst_last_error_string()
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under nfst_b200/ (or tools/) may import, link or execute it."""
    for dirpath, _, files in list(os.walk(os.path.join(ROOT, "nfst_b200"))) + list(os.walk(os.path.join(ROOT, "tools"))):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "lattice_oracle" not in text and "oracle/_" not in text, fn
    code = "import sys; import nfst_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
