"""CPU restatement (test infrastructure only) of the inference-side consolidation of the reference's predictor.py in the FORMULATION the device
kernels use (csrc/consolidate.cu): a fixed descending-score order with alive flags and a head pointer instead of a shrinking `order` array,
per-iteration stamps for the number of distinct patches, a slice-presence map for the z connectivity.  Pinned against golden vectors produced by
the reference's own functions (tests/golden/make_consolidate_golden.py -> tests/golden/consolidate.npz, tests/test_consolidate_cpu.py).

  weighted_box_clustering   predictor.py:597-706
  nms_2to3D                 predictor.py:710-773
"""
import numpy as np


def _iou_to_all(dets, i, dim):
    """IoU of box i with every box, `+ 1` pixel convention, the reference's expression order (predictor.py:640-657)"""
    y1, x1, y2, x2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    areas = (y2 - y1 + 1) * (x2 - x1 + 1)
    w = np.maximum(0.0, np.minimum(x2[i], x2) - np.maximum(x1[i], x1) + 1)
    h = np.maximum(0.0, np.minimum(y2[i], y2) - np.maximum(y1[i], y1) + 1)
    inter = w * h
    if dim == 3:
        z1, z2 = dets[:, 4], dets[:, 5]
        areas = areas * (z2 - z1 + 1)
        inter = inter * np.maximum(0.0, np.minimum(z2[i], z2) - np.maximum(z1[i], z1) + 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        return inter / (areas[i] + areas - inter), areas


def weighted_box_clustering(dets, box_patch_id, thresh, n_ens):
    dets = np.asarray(dets, dtype=np.float64)
    n = dets.shape[0]
    dim = 2 if dets.shape[1] == 7 else 3
    nc = 2 * dim
    scores, pc, novs = dets[:, nc], dets[:, nc + 1], dets[:, nc + 2]
    _, pid = np.unique(np.asarray(box_patch_id), return_inverse=True)
    order = np.argsort(-scores, kind="stable")
    alive = np.ones(n, dtype=bool)
    stamp = np.full(int(pid.max()) + 1 if n else 0, -1)
    keep_scores, keep_coords = [], []
    head, it = 0, 0
    while True:
        while head < n and not alive[order[head]]:
            head += 1
        if head >= n:
            break
        i = order[head]
        ovr, areas = _iou_to_all(dets, i, dim)
        match = alive & (ovr > thresh)
        m = np.nonzero(match)[0]
        wgt = ovr[m] * areas[m] * pc[m]                      # match_ov_facts * match_areas * match_pc_facts
        sc = scores[m] * wgt
        uniq = 0
        for j in m:                                          # np.unique(match_patch_id).shape[0] via per-iteration stamps
            if stamp[pid[j]] != it:
                stamp[pid[j]] = it
                uniq += 1
        c = len(m)
        if c:
            n_missing = max(0.0, n_ens * (novs[m].sum() / c) - uniq)
            denom = wgt.sum() + n_missing * (wgt.sum() / c)
            avg = sc.sum() / denom
            if avg > 0.01:
                keep_scores.append(float(avg))
                keep_coords.append([float((dets[m, k] * sc).sum() / sc.sum()) for k in range(nc)])
        alive[m] = False
        head += 1
        it += 1
    return keep_scores, keep_coords


def nms_2to3D(dets, thresh):
    dets = np.asarray(dets, dtype=np.float64)
    n = dets.shape[0]
    scores, slices = dets[:, 4], dets[:, 5].astype(np.int64)
    order = np.argsort(-scores, kind="stable")
    alive = np.ones(n, dtype=bool)
    n_slices = int(slices.max()) + 1 if n else 0
    keep, keep_z = [], []
    head = 0
    while True:
        while head < n and not alive[order[head]]:
            head += 1
        if head >= n:
            break
        i = order[head]
        ovr, _ = _iou_to_all(dets[:, :4], i, 2)
        match = alive & (ovr > thresh)
        present = np.zeros(n_slices, dtype=bool)
        present[slices[match]] = True
        core, smin, smax = slices[i], slices[match].min(), slices[match].max()
        hi = smax
        for s in range(core, smax):                           # first slice without a prediction above the core slice (:746-749)
            if not present[s]:
                hi = s
                break
        lo = smin
        for s in range(core - 1, smin - 1, -1):               # ... and below it
            if not present[s]:
                lo = s
                break
        zm = match & (slices <= hi) & (slices >= lo)
        keep.append(int(i))
        keep_z.append([float(slices[zm].min() - 1), float(slices[zm].max() + 1)])
        alive[zm] = False
        head += 1
    return keep, keep_z
