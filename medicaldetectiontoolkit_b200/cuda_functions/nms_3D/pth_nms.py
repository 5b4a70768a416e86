"""drop-in for cuda_functions/nms_3D/pth_nms.py:5-17"""
from ...native_ops import nms_gpu as _nms


def nms_gpu(dets, thresh):
    """dets [N, 7] (y1, x1, y2, x2, z1, z2, score) cuda -> indices of kept boxes, descending score"""
    return _nms(dets, thresh, dim=3)
