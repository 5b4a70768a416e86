#!/usr/bin/env bash
# First GPU trip of the next round (under gpurun, repo root): run what was written without hardware access.
#   1. tools/mma2_probe.cu       CTA-pair MMA semantics / cross-CTA TMA signalling / issue rate
#   2. opt-in parity tests       conv3d_tc_pair.cu (MDT_TC_PAIR=1), NMS division filter (MDT_NMS_FILTER=1)
#   3. A/B timings               conv layer bench and NMS microbench with and without the switches
# Every step is wrapped in `timeout`; a trap inside a kernel fails that step only.
mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/mma2_probe tools/mma2_probe.cu -lcuda 2>&1 | tail -3
timeout 120 /tmp/mma2_probe > gpurun_out/mma2_probe.txt 2>&1; cat gpurun_out/mma2_probe.txt
MDT_TEST_PAIR=1 timeout 300 python -m pytest tests/test_conv_gpu.py -q -x -k pair 2>&1 | tail -15 | tee gpurun_out/pytest_pair.txt
MDT_TEST_NMS_FILTER=1 timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k division_filter 2>&1 | tail -8 | tee gpurun_out/pytest_nms_filter.txt
for pair in 0 1; do
  echo "== MDT_TC_PAIR=$pair" | tee -a gpurun_out/layer_bench_pair.txt
  MDT_TC_PAIR=$pair PASSES=01 timeout 300 python tools/conv_layer_bench.py p0_36 c0_18 head64 2>&1 | tee -a gpurun_out/layer_bench_pair.txt
done
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
