"""drop-in for cuda_functions/nms_2D/pth_nms.py"""
from ...native_ops import nms_gpu as _nms


def nms_gpu(dets, thresh):
    """dets [N, 5] (y1, x1, y2, x2, score) cuda -> indices of kept boxes, descending score"""
    return _nms(dets, thresh, dim=2)
