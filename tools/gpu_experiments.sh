#!/usr/bin/env bash
# GPU trip (under gpurun, repo root): opt-in parity test + A/B timing of the NMS division filter (MDT_NMS_FILTER=1).
# Every step is wrapped in `timeout`; a trap inside a kernel fails that step only.
mkdir -p gpurun_out
MDT_TEST_NMS_FILTER=1 timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k division_filter 2>&1 | tail -8 | tee gpurun_out/pytest_nms_filter.txt
for f in 0 1; do
  echo "== MDT_NMS_FILTER=$f" | tee -a gpurun_out/nms_filter_ab.txt
  MDT_NMS_FILTER=$f timeout 300 python - <<'PY' 2>&1 | tee -a gpurun_out/nms_filter_ab.txt
import sys, torch
sys.path.insert(0, "tests")
import _oracle as O
from medicaldetectiontoolkit_b200 import native_ops as NO
for n, thr in [(100000, 0.5), (50000, 1e-5)]:
    t = torch.from_numpy(O.synth_boxes(n, 3, seed=n, rounded=True)).cuda()
    for _ in range(3):
        NO.nms_sorted(t, thr, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        NO.nms_sorted(t, thr, 3)
    e1.record(); torch.cuda.synchronize()
    print(n, thr, "%.3f ms" % (e0.elapsed_time(e1) / 5))
PY
done
