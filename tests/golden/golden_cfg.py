"""Config objects / random GT boxes shared by make_golden.py and the tests (no reference import here)."""
import types

import numpy as np


def cf3d(patch=(64, 64, 32)):
    cf = types.SimpleNamespace()
    cf.dim = 3
    cf.pyramid_levels = [0, 1, 2, 3]
    cf.rpn_anchor_ratios = [0.5, 1, 2]
    cf.rpn_anchor_stride = 1
    cf.backbone_strides = {'xy': [4, 8, 16, 32], 'z': [1, 2, 4, 8]}
    cf.rpn_anchor_scales = {'xy': [[8], [16], [32], [64]], 'z': [[2], [4], [8], [16]]}
    # retina variant: 3 scales per level (lidc_exp/configs.py:316-320)
    cf.rpn_anchor_scales['xy'] = [[ii[0], ii[0] * (2 ** (1 / 3)), ii[0] * (2 ** (2 / 3))] for ii in cf.rpn_anchor_scales['xy']]
    cf.rpn_anchor_scales['z'] = [[ii[0], ii[0] * (2 ** (1 / 3)), ii[0] * (2 ** (2 / 3))] for ii in cf.rpn_anchor_scales['z']]
    cf.backbone_shapes = np.array([[int(np.ceil(patch[0] / s)), int(np.ceil(patch[1] / s)), int(np.ceil(patch[2] / sz))]
                                   for s, sz in zip(cf.backbone_strides['xy'], cf.backbone_strides['z'])])
    cf.anchor_matching_iou = 0.5
    cf.rpn_train_anchors_per_image = 6
    cf.rpn_bbox_std_dev = np.array([0.1, 0.1, 0.1, 0.2, 0.2, 0.2])
    return cf


def cf2d(patch=(128, 128)):
    cf = types.SimpleNamespace()
    cf.dim = 2
    cf.pyramid_levels = [0, 1, 2, 3]
    cf.rpn_anchor_ratios = [0.5, 1, 2]
    cf.rpn_anchor_stride = 1
    cf.backbone_strides = {'xy': [4, 8, 16, 32]}
    cf.rpn_anchor_scales = {'xy': [[8], [16], [32], [64]]}
    cf.backbone_shapes = np.array([[int(np.ceil(patch[0] / s)), int(np.ceil(patch[1] / s))] for s in cf.backbone_strides['xy']])
    cf.anchor_matching_iou = 0.7
    cf.rpn_train_anchors_per_image = 64
    cf.rpn_bbox_std_dev = np.array([0.1, 0.1, 0.2, 0.2])
    return cf


def rand_gt(rs, n, extent, dim, lo=6, hi=40):
    out = []
    for _ in range(n):
        size = rs.randint(lo, hi, size=dim)
        c0 = [rs.randint(0, max(1, extent[k] - size[k])) for k in range(dim)]
        if dim == 3:
            out.append([c0[0], c0[1], c0[0] + size[0], c0[1] + size[1], c0[2], c0[2] + min(size[2], extent[2] - c0[2])])
        else:
            out.append([c0[0], c0[1], c0[0] + size[0], c0[1] + size[1]])
    return np.array(out)
