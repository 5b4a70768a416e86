#!/usr/bin/env python
"""Op-level microbenchmarks of BASELINE.json: 3D RoIAlign fwd/bwd (cfg3), 3D NMS (cfg4 + RPN/Retina shapes), anchor matching (cfg4 + cfg2),
each timed with CUDA events after warm-up with an L2 flush between iterations, next to the reference's own kernels (oracle/_ref) on the
same GPU and the CPU oracle on the host.  Prints one JSON object; `python tools/microbench.py > gpurun_out/microbench.json`."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import _oracle as O  # noqa: E402
import matching_oracle as MO  # noqa: E402
from golden_cfg import cf3d, rand_gt  # noqa: E402
from medicaldetectiontoolkit_b200 import model_utils as MU  # noqa: E402
from medicaldetectiontoolkit_b200 import native_ops as NO  # noqa: E402

DEV = "cuda:0"
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
_flush = None


def time_us(fn, iters=20, warmup=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        _flush.fill_(1)  # evict L2 (126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def graph_us(call, reps=20):
    """GPU-side duration of one `call` (us): `reps` calls captured into ONE CUDA graph with an L2 flush in front of each, minus the same
    graph with the flushes only.  The event-timed numbers above include the host's launch path (~25-30 us per Python->ctypes call),
    which hides every kernel shorter than that; this one does not."""
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            call()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

        def build(with_op):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(reps):
                    _flush.fill_(1)
                    if with_op:
                        call()
            return g

        def run(g):
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            return ts[len(ts) // 2]
        ga, gb = build(True), build(False)
        run(ga), run(gb)
        return (run(ga) - run(gb)) / reps
    except Exception as e:   # e.g. a launch type that cannot be captured
        torch.cuda.synchronize()
        return "graph timing failed: %s" % (str(e)[:200],)


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": PEAKS["hbm_gbs"]}
    # ---------------------------------------------------------------- RoIAlign cfg3
    rs = np.random.RandomState(0)
    roi = {}
    for name, shape, crop in [("P2_7x7x3", (2, 36, 32, 32, 128), (7, 7, 3)), ("P2_14x14x5", (2, 36, 32, 32, 128), (14, 14, 5)),
                              ("P3_7x7x3", (2, 36, 16, 16, 64), (7, 7, 3))]:
        img = rs.randn(*shape).astype(np.float32)
        boxes, ind = O.synth_rois(1024, 3, 2, seed=8)
        n, C = 1024, shape[1]
        P = int(np.prod(crop))
        alg_fwd = 4 * n * C * P + 24 * n + 4 * min(8 * n * C * P, int(np.prod(shape)))
        alg_bwd = 4 * n * C * P + 24 * n + 4 * int(np.prod(shape)) * 2                         # read grads + zero & scatter the image grad
        rec = {"alg_bytes_fwd": alg_fwd, "alg_bytes_bwd": alg_bwd}
        tb, ti = torch.from_numpy(boxes).to(DEV), torch.from_numpy(ind).to(DEV)
        for layout in ("channels_last", "ncdhw"):
            x = torch.from_numpy(img).to(DEV)
            if layout == "channels_last":
                x = x.contiguous(memory_format=torch.channels_last_3d)
            fn = NO.CropAndResizeFunction(*crop, 0)
            us = time_us(lambda: fn(x, tb, ti))
            xg = x.clone().requires_grad_(True)
            y = fn(xg, tb, ti)
            g = torch.randn_like(y)
            us_b = time_us(lambda: torch.autograd.grad(y, xg, g, retain_graph=True))
            rec[layout] = {"fwd_us": us, "fwd_gbs": alg_fwd / us / 1e3, "fwd_frac_hbm": alg_fwd / us / 1e3 / PEAKS["hbm_gbs"],
                           "bwd_us": us_b, "bwd_gbs": alg_bwd / us_b / 1e3, "bwd_frac_hbm": alg_bwd / us_b / 1e3 / PEAKS["hbm_gbs"]}
        # kernel-only: straight C-ABI calls on preallocated tensors inside a CUDA graph
        from medicaldetectiontoolkit_b200 import _lib as L
        lib = L.load()
        xcl = torch.from_numpy(img).to(DEV).contiguous(memory_format=torch.channels_last_3d)
        ycl = torch.empty([n, C] + list(crop), dtype=torch.float32, device=DEV, memory_format=torch.channels_last_3d)
        gcl = torch.randn_like(ycl)
        dxcl = torch.empty_like(xcl)
        tbf, tii = tb.contiguous().float(), ti.contiguous().int()
        xs, ys = L.i64arr(xcl.stride()), L.i64arr(ycl.stride())

        def k_fwd():
            L.check(lib.mdt_crop_and_resize_3d_forward(L.ptr(xcl), xs, L.ptr(tbf), L.ptr(tii), n, shape[0], shape[2], shape[3], shape[4],
                                                       crop[0], crop[1], crop[2], C, 0.0, L.ptr(ycl), ys, L.stream_ptr()))

        def k_bwd():
            L.check(lib.mdt_crop_and_resize_3d_backward(L.ptr(gcl), ys, L.ptr(tbf), L.ptr(tii), n, shape[0], shape[2], shape[3], shape[4],
                                                        crop[0], crop[1], crop[2], C, L.ptr(dxcl), xs, 1, dxcl.numel(), L.stream_ptr()))
        kf, kb = graph_us(k_fwd), graph_us(k_bwd)
        rec["channels_last_kernel_only"] = {"fwd_us": kf, "bwd_us_incl_memset": kb}
        if not isinstance(kf, str):
            rec["channels_last_kernel_only"].update({"fwd_gbs": alg_fwd / kf / 1e3, "fwd_frac_hbm": alg_fwd / kf / 1e3 / PEAKS["hbm_gbs"]})
        if not isinstance(kb, str):
            rec["channels_last_kernel_only"].update({"bwd_gbs": alg_bwd / kb / 1e3, "bwd_frac_hbm": alg_bwd / kb / 1e3 / PEAKS["hbm_gbs"]})
        if O.ref_lib("roi3d") is not None:
            t = [0.0]
            O.ref_crop_and_resize_forward(img, boxes, ind, crop, iters=20, times=t)
            rec["reference_kernel_fwd_us"] = t[0] * 1e3
            gy = rs.randn(n, C, *crop).astype(np.float32)
            O.ref_crop_and_resize_backward(gy, boxes, ind, shape, iters=20, times=t)
            rec["reference_kernel_bwd_us_incl_memset"] = t[0] * 1e3
        roi[name] = rec
    out["roi_align_3d"] = roi
    # ---------------------------------------------------------------- NMS
    nms = {}
    for name, n, thr, rounded in [("cfg4_100k_iou0.5", 100000, 0.5, True), ("retina_50k_1e-5", 50000, 1e-5, True), ("rpn_6000_0.7", 6000, 0.7, False)]:
        b = O.synth_boxes(n, 3, seed=n, rounded=rounded)
        t = torch.from_numpy(b).to(DEV)
        us = time_us(lambda: NO.nms_sorted(t, thr, 3), iters=10)
        keep, num = NO.nms_sorted(t, thr, 3)
        rec = {"n": n, "thresh": thr, "us": us, "kept": int(num.item()), "pair_tests_per_s": n * (n - 1) / 2 / (us * 1e-6),
               "alg_bytes": 28 * n + 8 * n * ((n + 63) // 64) * 2}
        os.environ["MDT_NMS_SCAN"] = "1"      # A/B: the single-CTA greedy reduction
        rec["us_single_cta_scan"] = time_us(lambda: NO.nms_sorted(t, thr, 3), iters=10)
        k1, n1 = NO.nms_sorted(t, thr, 3)
        rec["scan_variants_agree"] = bool(int(n1.item()) == int(num.item()) and torch.equal(k1[: int(n1.item())], keep[: int(num.item())]))
        del os.environ["MDT_NMS_SCAN"]
        if O.ref_lib("nms3d") is not None:
            times = [0.0, 0.0, 0.0]
            ref_keep = O.ref_nms(b, thr, 3, times)
            rec["reference_kernel_us"] = {"mask_kernel": times[0] * 1e3, "d2h_mask": times[1] * 1e3, "host_scan": times[2] * 1e3,
                                          "total": sum(times) * 1e3}
            rec["bit_identical_to_reference_kernel"] = bool(ref_keep.tolist() == keep[: int(num.item())].cpu().numpy().tolist())
        if n <= 10000:
            rec["us_kernel_only"] = graph_us(lambda: NO.nms_sorted(t, thr, 3))
        if n <= 20000:
            t0 = time.perf_counter()
            O.cpu_nms_baseline(b, thr, 3)
            rec["cpu_nms_c_1core_us"] = (time.perf_counter() - t0) * 1e6
        nms[name] = rec
    out["nms_3d"] = nms
    # ---------------------------------------------------------------- matching
    full = MU.generate_pyramid_anchors(None, cf3d((128, 128, 128)))
    sub = full[np.random.RandomState(0).permutation(full.shape[0])[:50000]]
    match = {}
    for name, anchors, G in [("cfg4_50k_G8", sub, 8), ("cfg4_50k_G64", sub, 64), ("cfg2_1.35M_G8", full, 8)]:
        gt = rand_gt(np.random.RandomState(G), G, (128, 128, 128), 3, 4, 48).astype(np.float64)
        cls = np.random.RandomState(1).randint(1, 3, size=G).astype(np.int32)
        a, g_, c_ = torch.from_numpy(anchors).to(DEV), torch.from_numpy(gt).to(DEV), torch.from_numpy(cls).to(DEV)
        us = time_us(lambda: MU.anchor_match_device(a, g_, c_, 3, 0.01, 0.5))
        A = anchors.shape[0]
        alg = 48 * A + 48 * G + 4 * A
        t0 = time.perf_counter()
        want, _ = MO.match_labels(anchors, gt, cls, 0.5, 3)
        cpu_us = (time.perf_counter() - t0) * 1e6
        got = MU.anchor_match_device(a, g_, c_, 3, 0.01, 0.5)[0].cpu().numpy()
        ko = graph_us(lambda: MU.anchor_match_device(a, g_, c_, 3, 0.01, 0.5))
        match[name] = {"A": A, "G": G, "us": us, "us_kernel_only": ko, "frac_hbm_kernel_only": None if isinstance(ko, str) else alg / ko / 1e3 / PEAKS["hbm_gbs"], "alg_bytes": alg, "gbs": alg / us / 1e3, "frac_hbm": alg / us / 1e3 / PEAKS["hbm_gbs"],
                       "numpy_f64_host_us": cpu_us, "bit_identical_to_numpy": bool(np.array_equal(got, want))}
    out["anchor_matching"] = match
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
