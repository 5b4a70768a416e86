"""Inference-side consolidation (csrc/consolidate.cu) against golden vectors produced by the REFERENCE's own numpy functions
(predictor.py:597-706 weighted_box_clustering, :710-773 nms_2to3D; tests/golden/make_consolidate_golden.py, run in the build container).
Cluster membership, order and count must be identical; fp64 averages agree to 1e-9 relative (np.sum is pairwise, the kernel sums in lane order)."""
import os

import numpy as np
import pytest

from medicaldetectiontoolkit_b200 import predictor as P

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "consolidate.npz"), allow_pickle=False)
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["wbc3d_a", "wbc3d_b", "wbc2d_a", "wbc3d_single", "wbc3d_big"])
def test_weighted_box_clustering_vs_reference(name):
    dets, pids = GOLD[name + "__dets"], GOLD[name + "__pids"]
    thresh, n_ens = GOLD[name + "__args"]
    ks, kc = P.weighted_box_clustering(dets, pids, float(thresh), float(n_ens))
    want_s, want_c = GOLD[name + "__keep_scores"], GOLD[name + "__keep_coords"]
    assert len(ks) == want_s.shape[0]
    if len(ks):
        np.testing.assert_allclose(np.array(ks), want_s, rtol=1e-9, atol=0)
        np.testing.assert_allclose(np.array(kc), want_c, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name", ["merge_a", "merge_b", "merge_single"])
def test_nms_2to3d_vs_reference(name):
    dets = GOLD[name + "__dets"]
    keep, keep_z = P.nms_2to3D(dets, float(GOLD[name + "__args"][0]))
    assert keep == GOLD[name + "__keep"].tolist()                                 # index-exact
    assert np.array_equal(np.array(keep_z).reshape(-1, 2), GOLD[name + "__keep_z"])


def test_patient_wrappers_keep_the_reference_dict_contract():
    """apply_wbc_to_patient / merge_2D_to_3D_preds_per_patient (predictor.py:513-593): same dict keys, gt boxes passed through"""
    dets, pids = GOLD["wbc3d_a__dets"], GOLD["wbc3d_a__pids"]
    boxes = [{'box_type': 'det', 'box_pred_class_id': 1, 'box_coords': d[:6], 'box_score': d[6], 'box_patch_center_factor': d[7], 'box_n_overlaps': d[8],
              'patch_id': str(p)} for d, p in zip(dets, pids)]
    boxes.append({'box_type': 'gt', 'box_coords': np.arange(6), 'box_label': 1})
    out, pid = P.apply_wbc_to_patient([[boxes], "pat0", {1: 'benign', 2: 'malignant'}, 0.1, 4])
    assert pid == "pat0" and len(out) == 1
    det = [b for b in out[0] if b['box_type'] == 'det']
    assert len(det) == GOLD["wbc3d_a__keep_scores"].shape[0] and out[0][-1]['box_type'] == 'gt'
    np.testing.assert_allclose([b['box_score'] for b in det], GOLD["wbc3d_a__keep_scores"], rtol=1e-9)
    d2 = GOLD["merge_a__dets"]
    n_slices = int(d2[:, 5].max()) + 1
    per_slice = [[{'box_type': 'det', 'box_pred_class_id': 1, 'box_coords': r[:4], 'box_score': r[4]} for r in d2 if int(r[5]) == s] for s in range(n_slices)]
    out3, _ = P.merge_2D_to_3D_preds_per_patient([per_slice, "pat0", {1: 'x'}, 0.1])
    assert len(out3) == 1 and len(out3[0]) == GOLD["merge_a__keep"].shape[0]
    assert all(len(b['box_coords']) == 6 for b in out3[0])
