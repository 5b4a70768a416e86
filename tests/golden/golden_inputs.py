"""Seeded inputs of the model-surface goldens, shared by make_model_golden.py (fed to the REFERENCE's functions) and the tests (fed to
ours): identical numpy bits on both sides, so only the reference's OUTPUTS are stored in the fixtures.  No reference import here."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from medicaldetectiontoolkit_b200.configs import make_cf, synthetic_batch  # noqa: E402  (plain attribute bags, no CUDA)


def softmax64(logits):
    z = logits.astype(np.float64)
    z = z - z.max(-1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(-1, keepdims=True)).astype(np.float32)


def small_cf(model):
    cf = make_cf(model, 3, (32, 32, 16))
    if model == 'mrcnn':
        cf.post_nms_rois_training = 64
    else:
        cf.pre_nms_limit = 3000
    return cf


def n_anchors(cf):
    return int(cf.n_anchors_per_pos * sum(int(np.prod(s)) for s in cf.backbone_shapes))


# ------------------------------------------------------------------------------------------------------------ retina_unet functions
def retina_refine_inputs(seed=101):
    cf = small_cf('retina_unet')
    A = n_anchors(cf)
    rs = np.random.RandomState(seed)
    probs = softmax64(rs.randn(2 * A, 3) * 2.0)
    deltas = (rs.randn(2 * A, 6) * 0.5).astype(np.float32)
    batch_ixs = np.repeat(np.arange(2), A)
    return cf, probs, deltas, batch_ixs


def class_loss_cases(seed=202):
    """name -> (anchor_matches int32 [A], logits f32 [A, n_cls], shem_poolsize)"""
    rs = np.random.RandomState(seed)
    out = {}
    for name, A, n_pos, n_neutral, n_cls, pool in [("few_pos", 5000, 3, 200, 3, 20), ("many_pos", 5000, 100, 300, 3, 20),
                                                   ("no_pos", 3000, 0, 100, 3, 20), ("no_neg", 64, 5, 59, 3, 20),
                                                   ("rpn_2cls", 4000, 4, 150, 2, 10), ("pool_exceeds", 300, 20, 250, 3, 20)]:
        m = -np.ones(A, dtype=np.int32)
        perm = rs.permutation(A)
        m[perm[:n_pos]] = rs.randint(1, n_cls, size=n_pos)
        m[perm[n_pos:n_pos + n_neutral]] = 0
        out[name] = (m, (rs.randn(A, n_cls) * 1.5).astype(np.float32), pool)
    return out


def bbox_loss_cases(seed=203):
    """name -> (target_deltas f32 [T, 6], pred_deltas f32 [A, 6], anchor_matches int32 [A])"""
    rs = np.random.RandomState(seed)
    out = {}
    for name, A, n_pos, T in [("few_pos", 5000, 3, 6), ("many_pos", 5000, 100, 200), ("no_pos", 1000, 0, 6)]:
        m = -np.ones(A, dtype=np.int32)
        m[rs.permutation(A)[:n_pos]] = 1
        t = np.zeros((T, 6), dtype=np.float32)
        t[:n_pos] = rs.randn(n_pos, 6)
        out[name] = (t, rs.randn(A, 6).astype(np.float32), m)
    return out


# ------------------------------------------------------------------------------------------------------------ mrcnn functions
def proposal_inputs(seed=301):
    cf = small_cf('mrcnn')
    A = n_anchors(cf)
    rs = np.random.RandomState(seed)
    probs = softmax64(rs.randn(2, A, 2) * 2.0)
    deltas = (rs.randn(2, A, 6) * 0.5).astype(np.float32)
    return cf, probs, deltas, 64


def _rand_rois(rs, n, n_batch, lo=0.04, hi=0.9):
    """normalised (y1, x1, y2, x2, z1, z2, batch_ix); side lengths log-uniform so that round(4 + log2(sqrt(h*w))) covers all four levels"""
    side = np.exp(rs.uniform(np.log(lo), np.log(hi), size=(n, 1))) * np.exp(rs.uniform(-0.3, 0.3, size=(n, 3)))
    side = np.minimum(side, 0.98)
    c0 = rs.uniform(0, 1, size=(n, 3)) * (1 - side)
    b = np.stack([c0[:, 0], c0[:, 1], c0[:, 0] + side[:, 0], c0[:, 1] + side[:, 1], c0[:, 2], c0[:, 2] + side[:, 2]], 1)
    return np.concatenate([b, rs.randint(0, n_batch, size=(n, 1))], 1).astype(np.float32)


def fpn_shapes(patch, channels=36, batch=2):
    return [(batch, channels, patch[0] // s, patch[1] // s, patch[2] // sz) for s, sz in zip((4, 8, 16, 32), (1, 2, 4, 8))]


def pyramid_inputs(size, seed=401):
    """size 'small': 32x32x16 patch, 96 rois; 'cfg3': 128^3 patch, 1024 rois (BASELINE config 3: 512 proposals x 2)"""
    rs = np.random.RandomState(seed)
    patch, n = ((32, 32, 16), 96) if size == 'small' else ((128, 128, 128), 1024)
    fmaps = [rs.randn(*s).astype(np.float32) for s in fpn_shapes(patch)]
    rois = _rand_rois(rs, n, 2)
    return fmaps, rois


def detection_target_inputs(seed=501):
    cf = small_cf('mrcnn')
    rs = np.random.RandomState(seed)
    p = cf.patch_size
    gt_boxes, gt_cls, gt_masks = [], [], []
    for b in range(2):
        boxes, masks = [], []
        for _ in range(2):
            size = [int(rs.randint(6, 14)), int(rs.randint(6, 14)), int(rs.randint(4, 8))]
            lo = [int(rs.randint(0, p[k] - size[k] + 1)) for k in range(3)]
            boxes.append([lo[0], lo[1], lo[0] + size[0], lo[1] + size[1], lo[2], lo[2] + size[2]])
            m = np.zeros(tuple(p) + (1,), dtype=np.uint8)
            m[lo[0]:lo[0] + size[0], lo[1]:lo[1] + size[1], lo[2]:lo[2] + size[2], 0] = 1
            masks.append(m)
        gt_boxes.append(np.array(boxes))
        gt_cls.append(rs.randint(1, 3, size=2))
        gt_masks.append(np.array(masks))
    # third element without GT (all-negative path): class ids all 0
    P = 48
    props = []
    scale = np.array([p[0], p[1], p[0], p[1], p[2], p[2]], dtype=np.float64)
    for b in range(2):
        jit = np.repeat(gt_boxes[b], 6, axis=0) + rs.uniform(-1.5, 1.5, size=(12, 6))
        rnd = _rand_rois(rs, P - 12, 1)[:, :6] * scale
        bx = np.concatenate([jit, rnd], 0) / scale
        bx = bx[rs.permutation(P)]
        props.append(np.concatenate([bx, np.full((P, 1), b)], 1))
    batch_proposals = np.concatenate(props, 0).astype(np.float32)
    scores = softmax64(rs.randn(2 * P, 3))
    return cf, batch_proposals, scores, gt_cls, gt_boxes, gt_masks


def mrcnn_refine_inputs(seed=601):
    cf = small_cf('mrcnn')
    rs = np.random.RandomState(seed)
    P = 64
    rois = _rand_rois(rs, 2 * P, 1)[:, :6]
    probs = softmax64(rs.randn(2 * P, 3) * 1.5)
    deltas = (rs.randn(2 * P, 3, 6) * 0.3).astype(np.float32)
    batch_ixs = np.repeat(np.arange(2), P).astype(np.float32)
    return cf, rois, probs, deltas, batch_ixs


def mrcnn_loss_inputs(seed=701):
    rs = np.random.RandomState(seed)
    n = 12
    t_cls = np.array([1, 2, 1, 2, 1, 0, 0, 0, 0, 0, 0, 0], dtype=np.int32)
    logits = rs.randn(n, 3).astype(np.float32)
    t_deltas = (rs.randn(n, 6) * (t_cls[:, None] > 0)).astype(np.float32)
    p_deltas = rs.randn(n, 3, 6).astype(np.float32)
    t_masks = (rs.rand(n, 28, 28, 10) > 0.5).astype(np.float32) * (t_cls[:, None, None, None] > 0)
    p_masks = (1 / (1 + np.exp(-rs.randn(n, 3, 28, 28, 10)))).astype(np.float32)
    return t_cls, logits, t_deltas, p_deltas, t_masks.astype(np.float32), p_masks


# ------------------------------------------------------------------------------------------------------------ utils
def utils_inputs(seed=801):
    rs = np.random.RandomState(seed)

    def boxes(n, dim):
        lo = rs.uniform(0, 40, size=(n, dim))
        ext = rs.uniform(2, 30, size=(n, dim))
        if dim == 3:
            return np.stack([lo[:, 0], lo[:, 1], lo[:, 0] + ext[:, 0], lo[:, 1] + ext[:, 1], lo[:, 2], lo[:, 2] + ext[:, 2]], 1).astype(np.float32)
        return np.stack([lo[:, 0], lo[:, 1], lo[:, 0] + ext[:, 0], lo[:, 1] + ext[:, 1]], 1).astype(np.float32)

    return dict(b3a=boxes(50, 3), b3b=boxes(7, 3), b2a=boxes(40, 2), b2b=boxes(5, 2),
                uniq=rs.randint(0, 20, size=200).astype(np.int64),
                dice_pred=softmax64(rs.randn(2, 16, 16, 8, 3)).transpose(0, 4, 1, 2, 3).copy(), dice_seg=rs.randint(0, 3, size=(2, 1, 16, 16, 8)).astype(np.uint8),
                shem_probs=softmax64(rs.randn(500, 3)), log2_x=rs.uniform(0.01, 4, size=64).astype(np.float32))


# ------------------------------------------------------------------------------------------------------------ whole models
MODEL_CASES = {
    # name: (model, dim, patch, batch, overrides)
    "retina_unet_small": ("retina_unet", 3, (64, 64, 32), 2, {}),
    "retina_unet_cfg2": ("retina_unet", 3, (128, 128, 128), 2, {}),
    "mrcnn_small": ("mrcnn", 3, (64, 64, 32), 2, {"post_nms_rois_training": 64, "roi_chunk_size": 1024, "_seed": 7}),
    "mrcnn_cfg3": ("mrcnn", 3, (128, 128, 128), 2, {"post_nms_rois_training": 512, "post_nms_rois_inference": 512, "roi_chunk_size": 1024}),
    "retina_net_cfg1": ("retina_net", 2, (128, 128), 1, {}),
}

# final layers scaled down after detweights.fill_ so that logits / deltas are O(1): saturated softmaxes would tie the SHEM / top-k orders
TAME = {
    "retina_unet": {"Classifier.conv_final": 0.05, "BBRegressor.conv_final": 0.05, "final_conv": 0.2},
    "retina_net": {"Classifier.conv_final": 0.05, "BBRegressor.conv_final": 0.05},
    "mrcnn": {"rpn.conv_class": 0.05, "rpn.conv_bbox": 0.05, "classifier.linear_class": 0.01, "classifier.linear_bbox": 0.1, "mask.conv5": 0.2},
}

GRAD_KEYS = {
    "retina_unet": ["Fpn.C1.0.weight", "Fpn.C0.1.0.bias", "Fpn.P0_conv2.weight", "Fpn.C3.0.conv2.0.weight", "Fpn.P2_conv1.bias",
                    "Classifier.conv_1.0.weight", "Classifier.conv_final.weight", "BBRegressor.conv_4.0.bias", "BBRegressor.conv_final.weight",
                    "final_conv.weight"],
    "retina_net": ["Fpn.C1.0.weight", "Classifier.conv_final.weight", "BBRegressor.conv_final.weight"],
    "mrcnn": ["fpn.C1.0.weight", "fpn.C3.0.conv2.0.weight", "fpn.P2_conv2.weight", "rpn.conv_shared.0.weight", "rpn.conv_class.weight",
              "rpn.conv_bbox.bias", "classifier.conv1.0.weight", "classifier.linear_class.weight", "classifier.linear_bbox.bias",
              "mask.conv1.0.weight", "mask.deconv.weight", "mask.conv5.weight"],
}


def model_case(name):
    model, dim, patch, batch, over = MODEL_CASES[name]
    cf = make_cf(model, dim, patch, exp='toy_exp' if dim == 2 else 'lidc_exp', batch_size=batch)
    for k, v in over.items():
        if not k.startswith("_"):
            setattr(cf, k, v)
    return cf, model, batch


def case_seed(name):
    """seed of the synthetic image of a model case (chosen so that the discrete selections of the reference run have clear margins)"""
    return MODEL_CASES[name][4].get("_seed", 5)


def tame_(net, model):
    import torch
    with torch.no_grad():
        for name, p in net.named_parameters():
            for prefix, f in TAME[model].items():
                if name.startswith(prefix + "."):
                    p.mul_(f)
    return net
