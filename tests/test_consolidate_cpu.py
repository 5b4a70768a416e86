"""CPU side of the consolidation row: the committed golden vectors are self-consistent and — where the reference tree is present
(build container) — regenerate bit-identically from the reference's unmodified predictor.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "consolidate.npz")


def test_golden_file_shapes():
    g = np.load(GOLD)
    for name in ("wbc3d_a", "wbc3d_b", "wbc2d_a", "wbc3d_big"):
        dim = 2 if g[name + "__dets"].shape[1] == 7 else 3
        assert g[name + "__keep_coords"].shape == (g[name + "__keep_scores"].shape[0], 2 * dim)
        assert np.all(g[name + "__keep_scores"] > 0.01)                             # predictor.py:700 filter
        assert np.unique(g[name + "__dets"][:, -3]).size == g[name + "__dets"].shape[0]   # unique scores: order is well defined
    for name in ("merge_a", "merge_b"):
        assert g[name + "__keep_z"].shape == (g[name + "__keep"].shape[0], 2)
        assert np.all(g[name + "__keep_z"][:, 1] - g[name + "__keep_z"][:, 0] >= 2)     # z1 = min - 1, z2 = max + 1


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present")
def test_golden_regenerates_from_the_reference(tmp_path):
    out = str(tmp_path / "consolidate.npz")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "golden", "make_consolidate_golden.py"), out], stdout=subprocess.DEVNULL)
    a, b = np.load(GOLD), np.load(out)
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import consolidate_oracle
    return consolidate_oracle


@pytest.mark.parametrize("name", ["wbc3d_a", "wbc3d_b", "wbc2d_a", "wbc3d_single", "wbc3d_big"])
def test_wbc_formulation_of_the_kernel_vs_reference_goldens(name):
    """oracle/consolidate_oracle.py restates predictor.py:597-706 in the device kernel's formulation (alive flags + head pointer, stamps for the
    distinct-patch count); same clusters in the same order as the reference's shrinking-`order` loop, fp64 averages to 1e-12"""
    g = np.load(GOLD)
    thresh, n_ens = g[name + "__args"]
    ks, kc = _oracle().weighted_box_clustering(g[name + "__dets"], g[name + "__pids"], float(thresh), float(n_ens))
    assert len(ks) == g[name + "__keep_scores"].shape[0]
    if len(ks):
        np.testing.assert_allclose(ks, g[name + "__keep_scores"], rtol=1e-12)
        np.testing.assert_allclose(np.array(kc), g[name + "__keep_coords"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["merge_a", "merge_b", "merge_single"])
def test_nms_2to3d_formulation_of_the_kernel_vs_reference_goldens(name):
    g = np.load(GOLD)
    keep, keep_z = _oracle().nms_2to3D(g[name + "__dets"], float(g[name + "__args"][0]))
    assert keep == g[name + "__keep"].tolist()
    assert np.array_equal(np.array(keep_z).reshape(-1, 2), g[name + "__keep_z"])
