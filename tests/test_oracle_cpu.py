"""CPU suite: the oracle against the golden fixtures generated from the reference's Python code, the numpy anchor generator mirror
against the reference's arrays, and internal consistency of the C oracle.  No GPU, no /root/reference at run time."""
import hashlib
import os
import sys
import types

import numpy as np
import torch

import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import matching_oracle as MO  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from golden_cfg import cf2d, cf3d  # noqa: E402

from medicaldetectiontoolkit_b200 import model_utils as MU  # noqa: E402


def test_anchor_generator_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "anchors.npz"))
    a3 = MU.generate_pyramid_anchors(None, cf3d())
    a2 = MU.generate_pyramid_anchors(None, cf2d())
    assert a3.dtype == np.float64 and a3.shape == g["a3"].shape
    assert np.array_equal(a3, g["a3"])  # bit-identical fp64
    assert np.array_equal(a2, g["a2"])
    full = MU.generate_pyramid_anchors(None, cf3d((128, 128, 128)))
    assert tuple(full.shape) == tuple(g["full_shape"]) == (1347840, 6)
    assert np.array_equal(full[:64], g["full_head"]) and np.array_equal(full[-64:], g["full_tail"])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(full).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, g["full_sha256"])


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "matching.npz"))
    anc = np.load(os.path.join(golden_dir, "anchors.npz"))
    for name in ["m3_sub", "m3_nosub", "m3_g1", "m2_nosub", "m2_sub"]:
        dim, tpi, seed, iou = g[name + "_cfg"]
        yield name, int(dim), int(tpi), int(seed), float(iou), anc["a3" if int(dim) == 3 else "a2"], g[name + "_gt"], g[name + "_cls"], \
            g[name + "_matches"], g[name + "_targets"]


def test_matching_oracle_vs_golden(golden_dir):
    """the numpy restatement reproduces the reference's gt_anchor_matching outputs bit-for-bit (labels) incl. the seeded sub-sampling"""
    for name, dim, tpi, seed, iou, anchors, gt, cls, want_m, want_t in _cases(golden_dir):
        labels, row_arg = MO.match_labels(anchors, gt, cls, iou, dim)
        ids = np.where(labels > 0)[0]
        extra = len(ids) - tpi // 2
        if extra > 0:
            np.random.seed(seed)
            labels[np.random.choice(ids, extra, replace=False)] = 0
        assert np.array_equal(labels, want_m), name
        ids = np.where(labels > 0)[0]
        std = [0.1] * dim + [0.2] * dim
        t = MO.delta_targets(anchors, gt, row_arg, ids, min(tpi, 64) if tpi > 1000 else tpi, std, dim)
        assert np.array_equal(t[: want_t.shape[0]], want_t), name  # same fp64 ops -> identical
    g = np.load(os.path.join(golden_dir, "matching.npz"))
    anc = np.load(os.path.join(golden_dir, "anchors.npz"))
    labels, _ = MO.match_labels(anc["a3"], g["m3_rpn_gt"], None, 0.5, 3)
    assert np.array_equal(labels, g["m3_rpn_matches"])


def test_box_coding_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "boxcoding.npz"))
    b3, d3, g3 = (torch.from_numpy(g[k]) for k in ("b3", "d3", "g3"))
    assert torch.equal(MU.apply_box_deltas_3D(b3, d3), torch.from_numpy(g["apply3"]))
    assert torch.equal(MU.apply_box_deltas_2D(b3[:, :4], d3[:, :4]), torch.from_numpy(g["apply2"]))
    assert torch.equal(MU.clip_boxes_3D(b3, [0, 0, 40, 40, 0, 20]), torch.from_numpy(g["clip3"]))
    assert torch.equal(MU.clip_to_window([0, 0, 40, 40, 0, 20], b3.clone()), torch.from_numpy(g["clip3"]))
    assert torch.equal(MU.box_refinement(b3, g3), torch.from_numpy(g["refine3"]))
    assert torch.equal(MU.box_refinement(b3[:, :4], g3[:, :4]), torch.from_numpy(g["refine2"]))


def test_nms_oracle_self_consistency():
    """greedy form == mask + serial scan form (nms_cuda.c:47-58), 2D and 3D, incl. empty / single / all-overlapping inputs"""
    for dim in (2, 3):
        for n, thr, rounded in [(0, 0.5, True), (1, 0.5, True), (65, 0.3, True), (300, 0.1, False), (513, 0.7, False)]:
            boxes = O.synth_boxes(n, dim, seed=n + dim, rounded=rounded, extent=64.0)
            keep = O.nms(boxes, thr, dim)
            mask = O.nms_mask(boxes, thr, dim)
            cb = (n + 63) // 64
            remv = np.zeros(max(cb, 1), dtype=np.uint64)
            want = []
            for i in range(n):
                if not (int(remv[i // 64]) >> (i % 64)) & 1:
                    want.append(i)
                    remv[i // 64:cb] |= mask[i, i // 64:]
            assert keep.tolist() == want
        same = np.tile(np.array([[2, 2, 10, 10] + ([1, 5] if dim == 3 else []) + [0.5]], dtype=np.float32), (70, 1))
        same[:, -1] = np.linspace(1, 0, 70)
        assert O.nms(same, 0.5, dim).tolist() == [0]


def test_roi_align_oracle_properties():
    """constant image -> constant crops; identity box at crop == map size reproduces the map; backward is the adjoint of forward"""
    rs = np.random.RandomState(0)
    for dim, shape, crop in [(3, (2, 3, 6, 5, 4), (3, 2, 2)), (2, (2, 3, 6, 5), (4, 3))]:
        img = rs.rand(*shape).astype(np.float32)
        boxes, ind = O.synth_rois(7, dim, shape[0], seed=1)
        ind[3] = 5  # out of range -> zeros
        out = O.crop_and_resize_forward(np.full(shape, 2.5, np.float32), boxes, ind, crop)
        assert np.allclose(np.delete(out, 3, axis=0), 2.5) and np.all(out[3] == 0)
        full = np.array([[0, 0, 1, 1] + ([0, 1] if dim == 3 else [])], dtype=np.float32)
        ident = O.crop_and_resize_forward(img, full, np.zeros(1, np.int32), shape[2:])
        assert np.allclose(ident[0], img[0], atol=1e-6)
        y = O.crop_and_resize_forward(img, boxes, ind, crop)
        gy = rs.rand(*y.shape).astype(np.float32)
        gx = O.crop_and_resize_backward(gy, boxes, ind, shape)
        assert abs(float((y.astype(np.float64) * gy).sum()) - float((img.astype(np.float64) * gx).sum())) < 1e-3
