"""Host-side mirror of the hot-path entry points of the reference's utils/model_utils.py, backed by libmdt_b200.so.

Same names, argument meaning and return types as the reference so its model files can bind to them unchanged:
  generate_pyramid_anchors   utils/model_utils.py:275-314 (+ generate_anchors :190-226, generate_anchors_3D :230-272)
  gt_anchor_matching         utils/model_utils.py:505-619  -> device kernels (csrc/anchor_match.cu)
  apply_box_deltas_{2D,3D}   :319-370, clip_boxes_{2D,3D} :376-398, clip_to_window :623-637, box_refinement :114-143
  NDConvGenerator            :732-781  -> re-exported from .conv (tcgen05 conv3d modules)
Anchor generation is init-time numpy (fp64, same operation order as the reference so the arrays are bit-identical); everything per
step runs on the GPU.
"""
import ctypes

import numpy as np
import torch

from . import _lib as L


# ----------------------------------------------------------------------------------------------------------------- anchors (init time)
def _anchor_shapes(scales_xy, scales_z, ratios):
    """per-position anchor extents, ratio-major / scale-minor like np.meshgrid(scales, ratios).flatten() in the reference"""
    s = np.asarray(scales_xy, dtype=np.float64)
    r = np.asarray(ratios, dtype=np.float64)
    sq = np.sqrt(np.repeat(r, len(s)))
    s_t = np.tile(s, len(r))
    heights = s_t / sq
    widths = s_t * sq
    depths = None
    if scales_z is not None:
        depths = np.tile(np.asarray(scales_z, dtype=np.float64), len(s_t) // len(scales_z))
    return heights, widths, depths


def generate_anchors(scales, ratios, shape, feature_stride, anchor_stride):
    """2D anchors [n_pos * n_shapes, (y1, x1, y2, x2)] f64; positions y-major/x-minor, shapes fastest (model_utils.py:190-226)"""
    h, w, _ = _anchor_shapes(scales, None, ratios)
    cy = (np.arange(0, shape[0], anchor_stride) * feature_stride).astype(np.float64)
    cx = (np.arange(0, shape[1], anchor_stride) * feature_stride).astype(np.float64)
    centers = np.stack(np.meshgrid(cy, cx, indexing="ij"), axis=-1).reshape(-1, 1, 2)
    sizes = np.stack([h, w], axis=-1).reshape(1, -1, 2)
    lo = (centers - 0.5 * sizes).reshape(-1, 2)
    hi = (centers + 0.5 * sizes).reshape(-1, 2)
    return np.concatenate([lo, hi], axis=1)


def generate_anchors_3D(scales_xy, scales_z, ratios, shape, feature_stride_xy, feature_stride_z, anchor_stride):
    """3D anchors [n_pos * n_shapes, (y1, x1, y2, x2, z1, z2)] f64; positions (y, x, z) with z fastest (model_utils.py:230-272)"""
    h, w, d = _anchor_shapes(scales_xy, scales_z, ratios)
    cy = (np.arange(0, shape[0], anchor_stride) * feature_stride_xy).astype(np.float64)
    cx = (np.arange(0, shape[1], anchor_stride) * feature_stride_xy).astype(np.float64)
    cz = (np.arange(0, shape[2], anchor_stride) * feature_stride_z).astype(np.float64)
    centers = np.stack(np.meshgrid(cy, cx, cz, indexing="ij"), axis=-1).reshape(-1, 1, 3)
    sizes = np.stack([h, w, d], axis=-1).reshape(1, -1, 3)
    lo = (centers - 0.5 * sizes).reshape(-1, 3)
    hi = (centers + 0.5 * sizes).reshape(-1, 3)
    return np.stack([lo[:, 0], lo[:, 1], hi[:, 0], hi[:, 1], lo[:, 2], hi[:, 2]], axis=1)


def generate_pyramid_anchors(logger, cf):
    """all pyramid levels concatenated, level order = cf.pyramid_levels (model_utils.py:275-314)"""
    out = []
    for level in cf.pyramid_levels:
        fshape = cf.backbone_shapes[level]
        if len(fshape) == 2:
            out.append(generate_anchors(cf.rpn_anchor_scales['xy'][level], cf.rpn_anchor_ratios, fshape,
                                        cf.backbone_strides['xy'][level], cf.rpn_anchor_stride))
        else:
            out.append(generate_anchors_3D(cf.rpn_anchor_scales['xy'][level], cf.rpn_anchor_scales['z'][level], cf.rpn_anchor_ratios, fshape,
                                           cf.backbone_strides['xy'][level], cf.backbone_strides['z'][level], cf.rpn_anchor_stride))
        if logger is not None:
            logger.info("level {}: built anchors {}".format(level, out[-1].shape))
    return np.concatenate(out, axis=0)


# ----------------------------------------------------------------------------------------------------------------- matching (per step, device)
_anchor_cache = {}


def _device_anchors(anchors, device):
    """fp64 device copy of the (constant) anchor array; cached so the 65 MB upload at A = 1.35 M happens once, not per step"""
    if torch.is_tensor(anchors):
        if anchors.dtype != torch.float64 or not anchors.is_cuda:
            raise L.MdtError("device anchors must be a float64 CUDA tensor")
        return anchors.contiguous()
    key = (anchors.__array_interface__['data'][0], anchors.shape, str(device))
    hit = _anchor_cache.get(key)
    if hit is None or hit[1] is not anchors:
        t = torch.from_numpy(np.ascontiguousarray(anchors, dtype=np.float64)).to(device)
        _anchor_cache.clear()
        _anchor_cache[key] = (t, anchors)
        return t
    return hit[0]


def anchor_match_device(anchors_dev, gt_boxes_dev, gt_class_ids_dev, dim, neg_iou_thresh, pos_iou_thresh):
    """Raw device call: labels BEFORE sub-sampling.  Returns (matches[A] i32, iou_argmax[A] i32, n_pos[1] i32), all CUDA, no sync."""
    lib = L.load()
    A = anchors_dev.shape[0]
    G = 0 if gt_boxes_dev is None else gt_boxes_dev.shape[0]
    dev = anchors_dev.device
    matches = torch.empty(A, dtype=torch.int32, device=dev)
    argmax = torch.empty(A, dtype=torch.int32, device=dev)
    n_pos = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.mdt_anchor_match_workspace_bytes(G)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.mdt_anchor_match(dim, L.ptr(anchors_dev), A, L.ptr(gt_boxes_dev), L.ptr(gt_class_ids_dev), G, float(neg_iou_thresh),
                                     float(pos_iou_thresh), L.ptr(ws), ws_bytes, L.ptr(matches), L.ptr(argmax), L.ptr(n_pos), L.stream_ptr()))
    return matches, argmax, n_pos


def anchor_delta_targets_device(anchors_dev, gt_boxes_dev, argmax, pos_ids, n_pos, max_targets, std_dev, dim):
    lib = L.load()
    out = torch.empty((max_targets, 2 * dim), dtype=torch.float64, device=anchors_dev.device)
    sd = (ctypes.c_double * (2 * dim))(*[float(v) for v in std_dev])
    with torch.cuda.device(anchors_dev.device):
        L.check(lib.mdt_anchor_delta_targets(dim, L.ptr(anchors_dev), L.ptr(gt_boxes_dev), L.ptr(argmax), L.ptr(pos_ids), int(n_pos),
                                             int(max_targets), sd, L.ptr(out), L.stream_ptr()))
    return out


def gt_anchor_matching_device(cf, anchors, gt_boxes, gt_class_ids=None, device=None, rng=None, return_pos=False):
    """Device-resident twin of gt_anchor_matching: same semantics, returns CUDA tensors (matches int32 [A], targets float64 [T, 2*dim]).

    rng: object with .choice(ids, extra, replace=False) used for the positive sub-sampling of model_utils.py:566-571; defaults to the
    numpy global RNG so that a seeded run reproduces the reference's stream exactly.
    return_pos: also return the ascending indices of the positive anchors (int64 CUDA) — the loss functions then need no search over the
    anchor array for them.
    """
    dim = cf.dim
    if device is None:
        device = anchors.device if torch.is_tensor(anchors) else torch.device("cuda", torch.cuda.current_device())
    a_dev = _device_anchors(anchors, device)
    A = a_dev.shape[0]
    T = cf.rpn_train_anchors_per_image
    if gt_boxes is None or len(gt_boxes) == 0:
        # model_utils.py:525-527 (gt_boxes is None): all negative, zero targets
        out = (torch.full((A,), -1, dtype=torch.int32, device=device), torch.zeros((T, 2 * dim), dtype=torch.float64, device=device))
        return out + (torch.zeros(0, dtype=torch.long, device=device),) if return_pos else out
    g_dev = torch.as_tensor(np.asarray(gt_boxes, dtype=np.float64)).reshape(-1, 2 * dim).to(device) if not torch.is_tensor(gt_boxes) \
        else gt_boxes.to(device=device, dtype=torch.float64).contiguous()
    c_dev = None
    if gt_class_ids is not None:
        c_dev = torch.as_tensor(np.asarray(gt_class_ids).astype(np.int32)).to(device) if not torch.is_tensor(gt_class_ids) \
            else gt_class_ids.to(device=device, dtype=torch.int32).contiguous()
    neg_t = 0.1 if dim == 2 else 0.01  # model_utils.py:549-552
    matches, argmax, n_pos = anchor_match_device(a_dev, g_dev, c_dev, dim, neg_t, cf.anchor_matching_iou)
    pos_ids = torch.nonzero(matches > 0).squeeze(1)  # ascending, like np.where; sizes the result -> one host sync
    extra = pos_ids.numel() - (T // 2)
    if extra > 0:
        chooser = rng if rng is not None else np.random
        drop = chooser.choice(pos_ids.cpu().numpy(), extra, replace=False)
        matches[torch.as_tensor(drop, device=device, dtype=torch.long)] = 0
        pos_ids = torch.nonzero(matches > 0).squeeze(1)
    targets = anchor_delta_targets_device(a_dev, g_dev, argmax, pos_ids.int().contiguous(), pos_ids.numel(), T, cf.rpn_bbox_std_dev, dim)
    return (matches, targets, pos_ids) if return_pos else (matches, targets)


def gt_anchor_matching(cf, anchors, gt_boxes, gt_class_ids=None):
    """Drop-in for utils/model_utils.py:505-619: numpy in, numpy out (anchor_class_matches int32 [A], anchor_delta_targets f64 [T, 2*dim]);
    the IoU matrix, arg-maxes and labelling run on the GPU in fp64 with numpy's exact operation order."""
    if gt_boxes is None:
        # the reference returns an int64 array here (np.full without dtype, :526)
        return np.full((anchors.shape[0],), -1), np.zeros((cf.rpn_train_anchors_per_image, 2 * cf.dim))
    m, t = gt_anchor_matching_device(cf, anchors, gt_boxes, gt_class_ids)
    return m.cpu().numpy(), t.cpu().numpy()


# ----------------------------------------------------------------------------------------------------------------- box coding (torch, device)
def _split(boxes):
    return [boxes[:, i] for i in range(boxes.shape[1])]


def apply_box_deltas_2D(boxes, deltas):
    """boxes [N,(y1,x1,y2,x2)], deltas [N,(dy,dx,log dh,log dw)] -> refined boxes (model_utils.py:319-340)"""
    h = boxes[:, 2] - boxes[:, 0]
    w = boxes[:, 3] - boxes[:, 1]
    cy = boxes[:, 0] + 0.5 * h + deltas[:, 0] * h
    cx = boxes[:, 1] + 0.5 * w + deltas[:, 1] * w
    h = h * torch.exp(deltas[:, 2])
    w = w * torch.exp(deltas[:, 3])
    y1 = cy - 0.5 * h
    x1 = cx - 0.5 * w
    return torch.stack([y1, x1, y1 + h, x1 + w], dim=1)


def apply_box_deltas_3D(boxes, deltas):
    """boxes [N,(y1,x1,y2,x2,z1,z2)], deltas [N,(dy,dx,dz,log dh,log dw,log dd)] (model_utils.py:343-370)"""
    h = boxes[:, 2] - boxes[:, 0]
    w = boxes[:, 3] - boxes[:, 1]
    d = boxes[:, 5] - boxes[:, 4]
    cy = boxes[:, 0] + 0.5 * h + deltas[:, 0] * h
    cx = boxes[:, 1] + 0.5 * w + deltas[:, 1] * w
    cz = boxes[:, 4] + 0.5 * d + deltas[:, 2] * d
    h = h * torch.exp(deltas[:, 3])
    w = w * torch.exp(deltas[:, 4])
    d = d * torch.exp(deltas[:, 5])
    y1 = cy - 0.5 * h
    x1 = cx - 0.5 * w
    z1 = cz - 0.5 * d
    return torch.stack([y1, x1, y1 + h, x1 + w, z1, z1 + d], dim=1)


def clip_boxes_2D(boxes, window):
    lo = boxes.new_tensor([window[0], window[1], window[0], window[1]], dtype=boxes.dtype)
    hi = boxes.new_tensor([window[2], window[3], window[2], window[3]], dtype=boxes.dtype)
    return torch.min(torch.max(boxes, lo), hi)


def clip_boxes_3D(boxes, window):
    lo = boxes.new_tensor([window[0], window[1], window[0], window[1], window[4], window[4]], dtype=boxes.dtype)
    hi = boxes.new_tensor([window[2], window[3], window[2], window[3], window[5], window[5]], dtype=boxes.dtype)
    return torch.min(torch.max(boxes, lo), hi)


def clip_to_window(window, boxes):
    """(model_utils.py:623-637) — note the argument order (window first).  Like the reference, only the coordinate columns are clamped;
    trailing columns (e.g. the batch index of mrcnn's sample_proposals) pass through."""
    n = 6 if boxes.shape[1] > 5 else 4
    clipped = clip_boxes_3D(boxes[:, :6], window) if n == 6 else clip_boxes_2D(boxes[:, :4], window)
    return clipped if boxes.shape[1] == n else torch.cat((clipped, boxes[:, n:]), dim=1)


def box_refinement(box, gt_box):
    """deltas that move `box` onto `gt_box` (model_utils.py:114-143)"""
    dim = box.shape[1] // 2
    cols = [(0, 2), (1, 3)] + ([(4, 5)] if dim == 3 else [])
    shift, scale = [], []
    for lo, hi in cols:
        e = box[:, hi] - box[:, lo]
        ge = gt_box[:, hi] - gt_box[:, lo]
        shift.append(((gt_box[:, lo] + 0.5 * ge) - (box[:, lo] + 0.5 * e)) / e)
        scale.append(torch.log(ge / e))
    return torch.stack(shift + scale, dim=1)


def unique1d(tensor):
    """sorted unique values (model_utils.py:645-654)"""
    if tensor.numel() < 2:
        return tensor
    return torch.unique(tensor, sorted=True)


def bbox_overlaps(boxes1, boxes2):
    """IoU matrix [n1, n2] of two box sets on the device, no +1 extents (utils/model_utils.py:430-501 bbox_overlaps_{2D,3D}),
    by broadcasting instead of the reference's repeat/tile copies; same operation order per pair: (dx*dy)[*dz], v1 + v2 - inter."""
    dim = boxes1.shape[1] // 2
    a, b = boxes1.unsqueeze(1), boxes2.unsqueeze(0)
    dy = (torch.min(a[..., 2], b[..., 2]) - torch.max(a[..., 0], b[..., 0])).clamp_min(0)
    dx = (torch.min(a[..., 3], b[..., 3]) - torch.max(a[..., 1], b[..., 1])).clamp_min(0)
    inter = dx * dy
    v1 = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    v2 = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    if dim == 3:
        inter = inter * (torch.min(a[..., 5], b[..., 5]) - torch.max(a[..., 4], b[..., 4])).clamp_min(0)
        v1 = v1 * (a[..., 5] - a[..., 4])
        v2 = v2 * (b[..., 5] - b[..., 4])
    return inter / (v1 + v2 - inter)


bbox_overlaps_2D = bbox_overlaps
bbox_overlaps_3D = bbox_overlaps


# ---- the remaining small helpers of utils/model_utils.py (host-side numpy / eager torch; none of them is on the training hot path, they complete
# the `mutils` surface that the reference's models, predictor and evaluator import)
def compute_iou_2D(box, boxes, box_area, boxes_area):
    """IoU of one box [y1, x1, y2, x2] with an array of boxes, areas passed in (utils/model_utils.py:35-54; no +1 extents)"""
    return _np_iou(box, boxes, box_area, boxes_area, 2)


def compute_iou_3D(box, boxes, box_volume, boxes_volume):
    """3D twin, boxes [y1, x1, y2, x2, z1, z2] (utils/model_utils.py:58-79)"""
    return _np_iou(box, boxes, box_volume, boxes_volume, 3)


def _np_iou(box, boxes, vol, vols, dim):
    box, boxes = np.asarray(box), np.asarray(boxes)
    inter = np.maximum(np.minimum(box[3], boxes[:, 3]) - np.maximum(box[1], boxes[:, 1]), 0) * \
        np.maximum(np.minimum(box[2], boxes[:, 2]) - np.maximum(box[0], boxes[:, 0]), 0)
    if dim == 3:
        inter = inter * np.maximum(np.minimum(box[5], boxes[:, 5]) - np.maximum(box[4], boxes[:, 4]), 0)
    return inter / (vol + vols - inter)


def compute_overlaps(boxes1, boxes2):
    """IoU matrix [n1, n2] of two numpy box sets (utils/model_utils.py:83-110), one broadcast instead of the loop over boxes2; the per-element
    expression and its operation order are the reference's, so the values are bit-identical"""
    b1, b2 = np.asarray(boxes1), np.asarray(boxes2)
    dim = b1.shape[1] // 2
    v1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    v2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    a, b = b1[:, None, :], b2[None, :, :]
    inter = np.maximum(np.minimum(b[..., 3], a[..., 3]) - np.maximum(b[..., 1], a[..., 1]), 0) * \
        np.maximum(np.minimum(b[..., 2], a[..., 2]) - np.maximum(b[..., 0], a[..., 0]), 0)
    if dim == 3:
        v1 = v1 * (b1[:, 5] - b1[:, 4])
        v2 = v2 * (b2[:, 5] - b2[:, 4])
        inter = inter * np.maximum(np.minimum(b[..., 5], a[..., 5]) - np.maximum(b[..., 4], a[..., 4]), 0)
    out = np.zeros((b1.shape[0], b2.shape[0]))
    out[...] = inter / (v2[None, :] + v1[:, None] - inter)
    return out


def clip_boxes_numpy(boxes, window):
    """clip [N, 4] / [N, 6] numpy boxes to the image (utils/model_utils.py:402-426).  As in the reference, y1 AND x1 are clipped to window[0]
    and y2 AND x2 to window[1] (square in-plane patches make this the same as a per-axis clip); z to window[2]."""
    boxes = np.asarray(boxes)
    hi = [window[0], window[0], window[1], window[1]] + ([window[2], window[2]] if boxes.shape[1] == 6 else [])
    return np.stack([np.clip(boxes[:, i], 0, hi[i]) for i in range(boxes.shape[1])], axis=1)


def intersect1d(tensor1, tensor2):
    """values present in both 1-D tensors, descending (utils/model_utils.py:667-670; both inputs hold unique values)"""
    aux = torch.cat((tensor1, tensor2), dim=0).sort(descending=True)[0]
    return aux[:-1][aux[1:] == aux[:-1]]


def sum_tensor(input, axes, keepdim=False):
    """sum over several axes (utils/model_utils.py:821-830)"""
    axes = sorted(set(int(a) for a in np.unique(axes)))
    return input.sum(dim=axes, keepdim=keepdim) if axes else input


def get_dice_per_batch_and_class(pred, y, n_classes):
    """hard dice per batch element and class of two label maps (b, 1, y, x, (z)) -> (b, c)   (utils/model_utils.py:803-818)"""
    p, t = get_one_hot_encoding(pred, n_classes), get_one_hot_encoding(y, n_classes)
    axes = tuple(range(2, p.ndim))
    return 2.0 * np.sum(p * t, axis=axes) / (np.sum(p, axis=axes) + np.sum(t, axis=axes) + 1e-8)


def batch_dice_mask(pred, y, mask, false_positive_weight=1.0, smooth=1e-6):
    """batch_dice restricted to a pixel mask in 2D (utils/model_utils.py:863-890; the 3D branch of the reference ignores the mask and so does this)"""
    if pred.dim() == 4:
        m = mask.unsqueeze(1).expand(-1, pred.shape[1], -1, -1)
        axes = (0, 2, 3)
        intersect = (pred * y * m).sum(axes)
        denom = (false_positive_weight * pred * m + y * m).sum(axes)
    elif pred.dim() == 5:
        axes = (0, 2, 3, 4)
        intersect = (pred * y).sum(axes)
        denom = (false_positive_weight * pred + y).sum(axes)
    else:
        raise ValueError('wrong input dimension in dice loss')
    return torch.mean(((2 * intersect + smooth) / (denom + smooth))[1:])


def unmold_mask_2D(mask, bbox, image_shape):
    """small mask -> full-size image at its box, scipy linear zoom (utils/model_utils.py:147-163)"""
    return _unmold(mask, [bbox[0], bbox[1]], [bbox[2], bbox[3]], tuple(image_shape[:2]))


def unmold_mask_3D(mask, bbox, image_shape):
    """3D twin (utils/model_utils.py:167-183)"""
    return _unmold(mask, [bbox[0], bbox[1], bbox[4]], [bbox[2], bbox[3], bbox[5]], tuple(image_shape[:3]))


def _unmold(mask, lo, hi, shape):
    import scipy.ndimage
    lo, hi = [int(v) for v in lo], [int(v) for v in hi]
    full = np.zeros(shape)
    size = [h - l for l, h in zip(lo, hi)]
    z = scipy.ndimage.zoom(mask, [s / m for s, m in zip(size, mask.shape)], order=1).astype(np.float32)
    full[tuple(slice(l, h) for l, h in zip(lo, hi))] = z
    return full


# Sampling mode of every stochastic draw on the path (SHEM pools, positive-roi sub-sampling): "random" = device RNG (the reference draws
# from the CPU RNG: same distribution, different stream); "identity" = the permutation is the identity, i.e. the reference with
# torch.randperm neutralised — the mode the reference-pinned parity tests run both sides in (tests/golden/make_model_golden.py).
SAMPLING = "random"


def randperm(n, device, generator=None):
    """torch.randperm(n) of utils/model_utils.py:690 / models/mrcnn.py:532 on the device"""
    if SAMPLING == "identity":
        return torch.arange(n, device=device)
    return torch.randperm(n, device=device, generator=generator)


def rand_keys(k, device, generator=None):
    """k sort keys in [0, 1): taking the j smallest keys of a masked subset == the first j entries of a random permutation of it"""
    if SAMPLING == "identity":
        return torch.arange(k, device=device, dtype=torch.float32) / max(k, 1)
    return torch.rand(k, device=device, generator=generator)


def shem(roi_probs_neg, negative_count, ohem_poolsize):
    """stochastic hard example mining (utils/model_utils.py:674-691): sample `negative_count` indices out of the
    `negative_count * ohem_poolsize` highest-scoring (max foreground probability) candidates; device randperm."""
    probs, order = roi_probs_neg[:, 1:].max(1)[0].sort(descending=True)
    select = min(ohem_poolsize * int(negative_count), order.shape[0])
    pool = order[:select]
    return pool[randperm(pool.shape[0], pool.device)[:negative_count]]


def log2(x):
    """utils/model_utils.py:658-663"""
    return torch.log2(x)


def get_one_hot_encoding(y, n_classes):
    """numpy drop-in of utils/model_utils.py:785-799: y (b, 1, y, x, (z)) integer labels -> (b, n_classes, y, x, (z)) int32"""
    y = np.asarray(y)
    out = np.zeros((y.shape[0], n_classes) + tuple(y.shape[2:]), dtype='int32')
    for c in range(n_classes):
        out[:, c][y[:, 0] == c] = 1
    return out


def batch_dice(pred, y, false_positive_weight=1.0, smooth=1e-6):
    """soft dice over the batch pseudo-volume, foreground classes only (utils/model_utils.py:833-858)"""
    axes = (0,) + tuple(range(2, pred.dim()))
    intersect = (pred * y).sum(axes)
    denom = (false_positive_weight * pred + y).sum(axes)
    return torch.mean(((2 * intersect + smooth) / (denom + smooth))[1:])


def initialize_weights(net):
    """cf.weight_init in {'xavier_uniform', 'xavier_normal', 'kaiming_uniform', 'kaiming_normal'} applied to every conv / transposed-conv /
    linear layer exactly as utils/model_utils.py:695-728 does for nn.Conv{2,3}d, nn.ConvTranspose{2,3}d and nn.Linear (the modules here keep
    the reference's parameter layouts, so the fan computations and the RNG consumption order are the same)."""
    import numpy as _np
    import torch.nn as nn
    init_type = net.cf.weight_init
    if init_type is None:
        return
    if init_type not in ('xavier_uniform', 'xavier_normal', 'kaiming_uniform', 'kaiming_normal'):
        raise NotImplementedError("cf.weight_init = %r (the reference silently leaves the default initialisation in that case)" % (init_type,))
    from .conv import Conv2d, Conv3d
    kinds = (Conv2d, Conv3d, nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d, nn.Linear)
    for m in net.modules():
        if not (isinstance(m, kinds) or type(m).__name__ == '_Deconv2x'):
            continue
        if init_type == 'xavier_uniform':
            nn.init.xavier_uniform_(m.weight.data)
            if m.bias is not None:
                m.bias.data.zero_()
        elif init_type == 'xavier_normal':
            nn.init.xavier_normal_(m.weight.data)
            if m.bias is not None:
                m.bias.data.zero_()
        elif init_type == 'kaiming_uniform':
            nn.init.kaiming_uniform_(m.weight.data, mode='fan_out', nonlinearity=net.cf.relu, a=0)
            if m.bias is not None:
                _, fan_out = nn.init._calculate_fan_in_and_fan_out(m.weight.data)
                bound = 1 / _np.sqrt(fan_out)
                nn.init.uniform_(m.bias, -bound, bound)
        else:
            nn.init.kaiming_normal_(m.weight.data, mode='fan_out', nonlinearity=net.cf.relu, a=0)
            if m.bias is not None:
                _, fan_out = nn.init._calculate_fan_in_and_fan_out(m.weight.data)
                bound = 1 / _np.sqrt(fan_out)
                nn.init.normal_(m.bias, -bound, bound)
