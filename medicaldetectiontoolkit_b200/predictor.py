"""Inference-side box consolidation on libmdt_b200 (csrc/consolidate.cu) behind the call surface of the reference's predictor.py:

  weighted_box_clustering(dets, box_patch_id, thresh, n_ens)   predictor.py:597-706
  nms_2to3D(dets, thresh)                                      predictor.py:710-773
  apply_wbc_to_patient(inputs)                                 predictor.py:513-549
  merge_2D_to_3D_preds_per_patient(inputs)                     predictor.py:553-593

Same arguments (numpy arrays / lists of box dicts) and return values as the reference, so `Predictor` can call them unchanged; the greedy cluster
loops run in one kernel each on the current CUDA device (fp64, the reference's expression order).  No CPU fallback.
"""
import numpy as np
import torch

from . import _lib as L


def _device():
    if not torch.cuda.is_available():
        raise L.MdtError("libmdt_b200 consolidation ops need a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def weighted_box_clustering(dets, box_patch_id, thresh, n_ens):
    """dets (n, (y1, x1, y2, x2, (z1), (z2), score, box_pc_fact, box_n_ov)), box_patch_id (n,) any hashable dtype
    -> (keep_scores: list of float, keep_coords: list of [y1, x1, y2, x2, (z1, z2)])"""
    dets = np.ascontiguousarray(dets, dtype=np.float64)
    n = dets.shape[0]
    if n == 0:
        return [], []
    dim = 2 if dets.shape[1] == 7 else 3
    lib = L.load()
    dev = _device()
    _, dense = np.unique(np.asarray(box_patch_id), return_inverse=True)
    n_patches = int(dense.max()) + 1
    d = torch.from_numpy(dets).to(dev)
    pid = torch.from_numpy(dense.astype(np.int32)).to(dev)
    order = torch.sort(d[:, -3], descending=True, stable=True)[1].to(torch.int32)
    scores = torch.empty(n, dtype=torch.float64, device=dev)
    coords = torch.empty((n, 2 * dim), dtype=torch.float64, device=dev)
    n_keep = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.mdt_wbc_workspace_bytes(n, n_patches)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.mdt_wbc(L.ptr(d), L.ptr(pid), L.ptr(order), n, dim, n_patches, float(thresh), float(n_ens), L.ptr(scores), L.ptr(coords), L.ptr(n_keep),
                            L.ptr(ws), ws_bytes, L.stream_ptr()))
    k = int(n_keep.item())
    return scores[:k].cpu().tolist(), coords[:k].cpu().tolist()


def nms_2to3D(dets, thresh):
    """dets (n, (y1, x1, y2, x2, score, slice_id)) -> (keep: list of indices, keep_z: list of [z1, z2])"""
    dets = np.ascontiguousarray(dets, dtype=np.float64)
    n = dets.shape[0]
    if n == 0:
        return [], []
    lib = L.load()
    dev = _device()
    if dets[:, -1].min() < 0:
        raise L.MdtError("nms_2to3D: slice ids must be non-negative")
    n_slices = int(dets[:, -1].max()) + 1
    d = torch.from_numpy(dets).to(dev)
    order = torch.sort(d[:, -2], descending=True, stable=True)[1].to(torch.int32)
    keep = torch.empty(n, dtype=torch.int64, device=dev)
    keep_z = torch.empty((n, 2), dtype=torch.float64, device=dev)
    n_keep = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.mdt_nms_2to3d_workspace_bytes(n, n_slices)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.mdt_nms_2to3d(L.ptr(d), L.ptr(order), n, float(thresh), n_slices, L.ptr(keep), L.ptr(keep_z), L.ptr(n_keep), L.ptr(ws), ws_bytes,
                                  L.stream_ptr()))
    k = int(n_keep.item())
    return keep[:k].cpu().tolist(), keep_z[:k].cpu().tolist()


def apply_wbc_to_patient(inputs):
    """(in_patient_results_list, pid, class_dict, wcs_iou, n_ens) -> [out_patient_results_list, pid]   (predictor.py:513-549)"""
    in_patient_results_list, pid, class_dict, wcs_iou, n_ens = inputs
    out_patient_results_list = [[] for _ in range(len(in_patient_results_list))]
    for bix, b in enumerate(in_patient_results_list):
        for cl in list(class_dict.keys()):
            boxes = [box for box in b if (box['box_type'] == 'det' and box['box_pred_class_id'] == cl)]
            if len(boxes) == 0:
                continue
            dets = np.concatenate((np.array([bx['box_coords'] for bx in boxes]),
                                   np.array([bx['box_score'] for bx in boxes])[:, None],
                                   np.array([bx['box_patch_center_factor'] for bx in boxes])[:, None],
                                   np.array([bx['box_n_overlaps'] for bx in boxes])[:, None]), axis=1)
            keep_scores, keep_coords = weighted_box_clustering(dets, np.array([bx['patch_id'] for bx in boxes]), wcs_iou, n_ens)
            for sc, co in zip(keep_scores, keep_coords):
                out_patient_results_list[bix].append({'box_type': 'det', 'box_coords': co, 'box_score': sc, 'box_pred_class_id': cl})
        out_patient_results_list[bix].extend([box for box in b if box['box_type'] == 'gt'])
    return [out_patient_results_list, pid]


def merge_2D_to_3D_preds_per_patient(inputs):
    """(in_patient_results_list, pid, class_dict, merge_3D_iou) -> [[boxes], pid]: slices in the batch dimension -> cubes (predictor.py:553-593)"""
    in_patient_results_list, pid, class_dict, merge_3D_iou = inputs
    out_patient_results_list = []
    for cl in list(class_dict.keys()):
        boxes, slice_ids = [], []
        for bix, b in enumerate(in_patient_results_list):
            det_boxes = [box for box in b if (box['box_type'] == 'det' and box['box_pred_class_id'] == cl)]
            boxes += det_boxes
            slice_ids += [bix] * len(det_boxes)
        if len(boxes) == 0:
            continue
        box_coords = np.array([bx['box_coords'] for bx in boxes])
        box_scores = np.array([bx['box_score'] for bx in boxes])
        keep_ix, keep_z = nms_2to3D(np.concatenate((box_coords, box_scores[:, None], np.array(slice_ids)[:, None]), axis=1), merge_3D_iou)
        for kix, kz in zip(keep_ix, keep_z):
            out_patient_results_list.append({'box_type': 'det', 'box_coords': list(box_coords[kix]) + kz, 'box_score': box_scores[kix],
                                             'box_pred_class_id': cl})
    out_patient_results_list += [box for b in in_patient_results_list for box in b if box['box_type'] == 'gt']
    return [[out_patient_results_list], pid]
