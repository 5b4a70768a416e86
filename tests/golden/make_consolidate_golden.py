#!/usr/bin/env python
"""Golden vectors for the inference-side consolidation, produced by the REFERENCE's own numpy functions
(/root/reference/predictor.py:597-706 weighted_box_clustering, :710-773 nms_2to3D) imported unmodified under the import shims.
Run in the build container:  python tests/golden/make_consolidate_golden.py [out.npz]   ->  tests/golden/consolidate.npz
Inputs are seeded; scores are unique (argsort()[::-1] leaves the order of ties to numpy's unstable quicksort)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims as RS  # noqa: E402


def wbc_case(rs, n_centres, per_centre, dim, extent=256.0):
    """overlapping patch predictions: `per_centre` jittered boxes around each of `n_centres` objects, from several patches"""
    rows, pids = [], []
    for c in range(n_centres):
        centre = rs.uniform(20, extent - 20, size=dim)
        size = rs.uniform(6, 40, size=dim)
        for k in range(rs.randint(1, per_centre + 1)):
            ctr = centre + rs.normal(0, 1.5, size=dim)
            sz = size * rs.uniform(0.8, 1.25, size=dim)
            lo, hi = np.round(ctr - sz / 2), np.round(ctr + sz / 2)
            box = [lo[0], lo[1], hi[0], hi[1]] + ([lo[2], hi[2]] if dim == 3 else [])
            rows.append(box + [0.0, rs.uniform(0.05, 1.0), float(rs.choice([1, 2, 4, 8]))])
            pids.append("p%d_%d" % (rs.randint(0, 12), rs.randint(0, 3)))
    dets = np.array(rows, dtype=np.float64)
    dets[:, -3] = rs.permutation(np.linspace(0.02, 0.99, dets.shape[0]))       # unique scores
    return dets, np.array(pids)


def merge_case(rs, n_objects, n_slices, extent=256.0):
    rows = []
    for o in range(n_objects):
        centre = rs.uniform(20, extent - 20, size=2)
        size = rs.uniform(8, 40, size=2)
        z0 = rs.randint(0, n_slices - 3)
        z1 = min(n_slices, z0 + rs.randint(2, 14))
        for z in range(z0, z1):
            if rs.rand() < 0.18:                                            # holes interrupt the cube (predictor.py:746-751)
                continue
            for _ in range(rs.randint(1, 3)):
                ctr = centre + rs.normal(0, 1.0, size=2)
                sz = size * rs.uniform(0.85, 1.15, size=2)
                lo, hi = np.round(ctr - sz / 2), np.round(ctr + sz / 2)
                rows.append([lo[0], lo[1], hi[0], hi[1], 0.0, float(z)])
    dets = np.array(rows, dtype=np.float64)
    dets[:, 4] = rs.permutation(np.linspace(0.1, 0.99, dets.shape[0]))
    return dets


def main():
    RS.install_import_shims()
    import predictor as ref                                                   # the reference module, unmodified
    out = {}
    cases = [("wbc3d_a", 3, 12, 6, 0.1, 4), ("wbc3d_b", 3, 60, 9, 1e-5, 1), ("wbc2d_a", 2, 25, 7, 0.1, 5), ("wbc3d_single", 3, 1, 1, 0.1, 3),
             ("wbc3d_big", 3, 300, 8, 0.1, 4)]
    for name, dim, nc, per, thresh, n_ens in cases:
        rs = np.random.RandomState(sum(map(ord, name)))
        dets, pids = wbc_case(rs, nc, per, dim)
        ks, kc = ref.weighted_box_clustering(dets.copy(), pids, thresh, n_ens)
        out[name + "__dets"] = dets
        out[name + "__pids"] = pids
        out[name + "__args"] = np.array([thresh, n_ens], dtype=np.float64)
        out[name + "__keep_scores"] = np.array(ks, dtype=np.float64)
        out[name + "__keep_coords"] = np.array(kc, dtype=np.float64).reshape(len(ks), 2 * dim)
        print(name, dets.shape, "->", len(ks), "clusters")
    for name, nobj, nsl, thresh in [("merge_a", 6, 40, 0.1), ("merge_b", 40, 96, 0.3), ("merge_single", 1, 8, 0.1)]:
        rs = np.random.RandomState(sum(map(ord, name)))
        dets = merge_case(rs, nobj, nsl)
        keep, keep_z = ref.nms_2to3D(dets.copy(), thresh)
        out[name + "__dets"] = dets
        out[name + "__args"] = np.array([thresh], dtype=np.float64)
        out[name + "__keep"] = np.array(keep, dtype=np.int64)
        out[name + "__keep_z"] = np.array(keep_z, dtype=np.float64).reshape(len(keep), 2)
        print(name, dets.shape, "->", len(keep), "cubes")
    np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "consolidate.npz"), **out)


if __name__ == "__main__":
    main()
