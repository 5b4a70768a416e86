"""The C-ABI library loads and exports exactly the symbols include/mdt_b200.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

import pytest

from medicaldetectiontoolkit_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mdt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_python_binding_covers_header():
    assert sorted(L.SIGNATURES) == _declared()


def test_load_and_host_only_calls():
    lib = L.load()
    assert lib.mdt_version() >= 100
    assert lib.mdt_error_string(0) == b"success"
    assert b"workspace" in lib.mdt_error_string(-2)
    # 1.25 GB mask at cfg4, + the grid scan's suppression bitmap (1563 words), 32-byte control block, 1563 tile extents
    assert lib.mdt_nms_workspace_bytes(100000) == 100000 * 1563 * 8 + 1563 * 8 + 32 + 1563 * 8
    assert lib.mdt_nms_workspace_bytes(0) == 0
    assert lib.mdt_anchor_match_workspace_bytes(8) >= 8 * 12


def test_no_oracle_in_product_path():
    """the product package must never import/load anything under oracle/"""
    pkg = os.path.join(ROOT, "medicaldetectiontoolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "libmdt_oracle" not in src and "import _oracle" not in src and "oracle/_ref" not in src, f


def test_binding_arity_matches_the_header():
    """ctypes never checks a prototype: count the parameters of every declaration in include/mdt_b200.h and compare with the bound argtypes
    (a missing or extra argument would silently shift every later one)"""
    text = open(os.path.join(ROOT, "include", "mdt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(mdt_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 40
    seen = set()
    for name, params in protos:
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert name in L.SIGNATURES, name
        assert len(L.SIGNATURES[name][1]) == n, "%s: header declares %d parameters, the ctypes binding passes %d" % (name, n, len(L.SIGNATURES[name][1]))
        seen.add(name)
    assert seen == set(L.SIGNATURES)


def test_binding_scalar_kinds_match_the_header():
    """pointer / int / float / double / size_t / 64-bit of every parameter: a float bound where the header says double (or the reverse) passes
    garbage without any error"""
    text = open(os.path.join(ROOT, "include", "mdt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    bad = []
    for name, params in re.findall(r"\b(mdt_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = " ".join(params.split())
        if params in ("", "void"):
            continue
        for i, (decl, ct) in enumerate(zip(params.split(","), L.SIGNATURES[name][1])):
            decl = decl.strip()
            if "*" in decl:
                ok = ct in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(ct, "_type_") and not issubclass(ct, ctypes._SimpleCData)
            elif re.search(r"\bdouble\b", decl):
                ok = ct is ctypes.c_double
            elif re.search(r"\bfloat\b", decl):
                ok = ct is ctypes.c_float
            elif re.search(r"\bsize_t\b", decl):
                ok = ct is ctypes.c_size_t
            elif re.search(r"\blong long\b|\bint64_t\b", decl):
                ok = ct in (ctypes.c_int64, ctypes.c_longlong)
            else:
                ok = ct in (ctypes.c_int, ctypes.c_uint)
            if not ok:
                bad.append((name, i, decl, ct))
    assert not bad, bad


def test_header_is_plain_c_and_links(tmp_path):
    """the boundary is a C ABI: include/mdt_b200.h must compile as C99 (no C++-only constructs outside the extern "C" guards) and a C program
    must link against the shared library and call a host-only entry point"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "mdt_b200.h"\n#include <stdio.h>\nint main(void) { printf("%d %s\\n", mdt_version(), mdt_error_string(-2)); return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir, "-lmdt_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)]).decode()
    assert out.split()[0] == str(L.load().mdt_version()) and "workspace" in out
