#!/usr/bin/env bash
# Compiles the reference's own CUDA kernels, UNMODIFIED and from where they lie under $REF (default /root/reference), for sm_100a,
# together with oracle/ref_harness.cu, into oracle/_ref/libref_{nms2d,nms3d,roi2d,roi3d}.so.
# oracle/_ref/ is git-ignored (no reference code enters history) but travels to the GPU box with the snapshot.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/cuda_functions" ]; then echo "build_ref: $REF not present - keeping prebuilt $OUT" >&2; exit 0; fi
mkdir -p "$OUT"
NV="nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared"
for d in 2 3; do
  S="$REF/cuda_functions/nms_${d}D/src/cuda"
  $NV -DREF_NMS -I"$S" -o "$OUT/libref_nms${d}d.so" "$S/nms_kernel.cu" "$HERE/ref_harness.cu"
  S="$REF/cuda_functions/roi_align_${d}D/roi_align/src/cuda"
  $NV -DREF_ROI${d}D -I"$S" -o "$OUT/libref_roi${d}d.so" "$S/crop_and_resize_kernel.cu" "$HERE/ref_harness.cu"
done
ls -la "$OUT"
