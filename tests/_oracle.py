"""ctypes access to oracle/libmdt_oracle.so and oracle/_ref/*.so — the CHECKERS.  Imported by tests, smoke() and bench.py's CPU legs only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "libmdt_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        _lib = ctypes.CDLL(path)
        _lib.oracle_nms.restype = ctypes.c_int64
        _lib.oracle_cpu_nms_baseline.restype = ctypes.c_int64
    return _lib


def nms(boxes_sorted, thresh, dim):
    b = _f32(boxes_sorted)
    n = b.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = lib().oracle_nms(b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n), ctypes.c_int(dim), ctypes.c_float(thresh),
                         keep.ctypes.data_as(ctypes.c_void_p))
    return keep[:k].copy()


def cpu_nms_baseline(boxes_sorted, thresh, dim):
    b = _f32(boxes_sorted)
    n = b.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    k = lib().oracle_cpu_nms_baseline(b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n), ctypes.c_int(dim), ctypes.c_float(thresh),
                                      keep.ctypes.data_as(ctypes.c_void_p))
    return keep[:k].copy()


def nms_mask(boxes_sorted, thresh, dim):
    b = _f32(boxes_sorted)
    n = b.shape[0]
    mask = np.zeros((n, (n + 63) // 64), dtype=np.uint64)
    lib().oracle_nms_mask(b.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n), ctypes.c_int(dim), ctypes.c_float(thresh),
                          mask.ctypes.data_as(ctypes.c_void_p))
    return mask


def _geom(image, crop):
    dim = image.ndim - 2
    B, C = image.shape[:2]
    H, W = image.shape[2:4]
    Z = image.shape[4] if dim == 3 else 1
    ch, cw = crop[:2]
    cz = crop[2] if dim == 3 else 1
    return dim, B, C, H, W, Z, ch, cw, cz


def crop_and_resize_forward(image, boxes, box_ind, crop):
    """image NC(D)HW contiguous numpy f32; returns [n, C, *crop]"""
    image = _f32(image)
    boxes = _f32(boxes)
    box_ind = np.ascontiguousarray(box_ind, dtype=np.int32)
    dim, B, C, H, W, Z, ch, cw, cz = _geom(image, crop)
    n = boxes.shape[0]
    out = np.zeros((n, C) + tuple(crop), dtype=np.float32)
    lib().oracle_crop_and_resize_forward(image.ctypes.data_as(ctypes.c_void_p), boxes.ctypes.data_as(ctypes.c_void_p),
                                         box_ind.ctypes.data_as(ctypes.c_void_p), n, B, C, H, W, Z, ch, cw, cz, dim,
                                         out.ctypes.data_as(ctypes.c_void_p))
    return out


def crop_and_resize_backward(grads, boxes, box_ind, image_shape):
    grads = _f32(grads)
    boxes = _f32(boxes)
    box_ind = np.ascontiguousarray(box_ind, dtype=np.int32)
    crop = grads.shape[2:]
    dim, B, C, H, W, Z, ch, cw, cz = _geom(np.empty(image_shape, dtype=np.float32), crop)
    out = np.zeros(image_shape, dtype=np.float32)
    lib().oracle_crop_and_resize_backward(grads.ctypes.data_as(ctypes.c_void_p), boxes.ctypes.data_as(ctypes.c_void_p),
                                          box_ind.ctypes.data_as(ctypes.c_void_p), boxes.shape[0], B, C, H, W, Z, ch, cw, cz, dim,
                                          out.ctypes.data_as(ctypes.c_void_p))
    return out


# ------------------------------------------------------------------ the reference's own kernels (oracle/_ref, GPU only)
_ref = {}


def ref_lib(name):
    """name in {nms2d, nms3d, roi2d, roi3d}; None if oracle/_ref was not built (the build needs /root/reference)"""
    if name not in _ref:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_%s.so" % name)
        _ref[name] = ctypes.CDLL(path) if os.path.exists(path) else None
        if _ref[name] is not None and name.startswith("nms"):
            _ref[name].ref_nms.restype = ctypes.c_longlong
    return _ref[name]


def ref_nms(boxes_sorted, thresh, dim, times=None):
    L = ref_lib("nms%dd" % dim)
    b = _f32(boxes_sorted)
    n = b.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    t = (ctypes.c_double * 3)()
    k = L.ref_nms(b.ctypes.data_as(ctypes.c_void_p), n, 2 * dim + 1, ctypes.c_float(thresh), keep.ctypes.data_as(ctypes.c_void_p), t)
    if k < 0:
        raise RuntimeError("ref_nms failed: %d" % k)
    if times is not None:
        times[:] = list(t)
    return keep[:k].copy()


def ref_nms_mask(boxes_sorted, thresh, dim):
    L = ref_lib("nms%dd" % dim)
    b = _f32(boxes_sorted)
    n = b.shape[0]
    mask = np.zeros((n, (n + 63) // 64), dtype=np.uint64)
    rc = L.ref_nms_mask(b.ctypes.data_as(ctypes.c_void_p), n, 2 * dim + 1, ctypes.c_float(thresh), mask.ctypes.data_as(ctypes.c_void_p))
    if rc:
        raise RuntimeError("ref_nms_mask failed: %d" % rc)
    return mask


def ref_crop_and_resize_forward(image, boxes, box_ind, crop, iters=1, times=None):
    image = _f32(image)
    boxes = _f32(boxes)
    box_ind = np.ascontiguousarray(box_ind, dtype=np.int32)
    dim, B, C, H, W, Z, ch, cw, cz = _geom(image, crop)
    L = ref_lib("roi%dd" % dim)
    out = np.zeros((boxes.shape[0], C) + tuple(crop), dtype=np.float32)
    t = (ctypes.c_double * 1)()
    rc = L.ref_crop_and_resize_forward(image.ctypes.data_as(ctypes.c_void_p), boxes.ctypes.data_as(ctypes.c_void_p),
                                       box_ind.ctypes.data_as(ctypes.c_void_p), boxes.shape[0], B, C, H, W, Z, ch, cw, cz,
                                       out.ctypes.data_as(ctypes.c_void_p), iters, t)
    if rc:
        raise RuntimeError("ref roi fwd failed: %d" % rc)
    if times is not None:
        times[:] = [t[0]]
    return out


def ref_crop_and_resize_backward(grads, boxes, box_ind, image_shape, iters=1, times=None):
    grads = _f32(grads)
    boxes = _f32(boxes)
    box_ind = np.ascontiguousarray(box_ind, dtype=np.int32)
    crop = grads.shape[2:]
    dim, B, C, H, W, Z, ch, cw, cz = _geom(np.empty(image_shape, dtype=np.float32), crop)
    L = ref_lib("roi%dd" % dim)
    out = np.zeros(image_shape, dtype=np.float32)
    t = (ctypes.c_double * 1)()
    rc = L.ref_crop_and_resize_backward(grads.ctypes.data_as(ctypes.c_void_p), boxes.ctypes.data_as(ctypes.c_void_p),
                                        box_ind.ctypes.data_as(ctypes.c_void_p), boxes.shape[0], B, C, H, W, Z, ch, cw, cz,
                                        out.ctypes.data_as(ctypes.c_void_p), iters, t)
    if rc:
        raise RuntimeError("ref roi bwd failed: %d" % rc)
    if times is not None:
        times[:] = [t[0]]
    return out


# ------------------------------------------------------------------ synthetic inputs
def synth_boxes(n, dim, seed, rounded=True, extent=128.0, smin=4.0, smax=48.0):
    """[n, 2*dim+1] f32 sorted by descending UNIQUE score (torch.sort tie order is unspecified, SURVEY §7 hard part 5)"""
    rs = np.random.RandomState(seed)
    c = rs.uniform(0, extent, size=(n, dim))
    s = rs.uniform(smin, smax, size=(n, dim))
    lo = np.clip(c - s / 2, 0, extent)
    hi = np.clip(c + s / 2, 0, extent)
    if rounded:
        lo, hi = np.round(lo), np.round(hi)
    cols = [lo[:, 0], lo[:, 1], hi[:, 0], hi[:, 1]] + ([lo[:, 2], hi[:, 2]] if dim == 3 else [])
    scores = rs.permutation(np.linspace(0, 1, n))
    order = np.argsort(-scores, kind="stable")
    out = np.stack(cols + [scores], axis=1).astype(np.float32)
    return out[order]


def synth_rois(n, dim, batch, seed, spread_levels=True):
    """normalised boxes (y1,x1,y2,x2[,z1,z2]) in [0,1], box_ind int32"""
    rs = np.random.RandomState(seed)
    size = np.exp(rs.uniform(np.log(0.03), np.log(0.6), size=(n, dim))) if spread_levels else rs.uniform(0.05, 0.5, size=(n, dim))
    lo = rs.uniform(0, 1, size=(n, dim)) * (1 - size)
    hi = lo + size
    cols = [lo[:, 0], lo[:, 1], hi[:, 0], hi[:, 1]] + ([lo[:, 2], hi[:, 2]] if dim == 3 else [])
    return np.stack(cols, axis=1).astype(np.float32), rs.randint(0, batch, size=n).astype(np.int32)
