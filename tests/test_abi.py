"""The C-ABI library loads and exports exactly the symbols include/mdt_b200.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

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
