// ref_harness.cu — drives the reference's OWN, unmodified CUDA kernels on the B200.  TEST INFRASTRUCTURE ONLY (see mdt_oracle.c header).
//
// build_ref.sh compiles this file together with one reference source taken where it lies under /root/reference
// (cuda_functions/nms_{2D,3D}/src/cuda/nms_kernel.cu, cuda_functions/roi_align_{2D,3D}/roi_align/src/cuda/crop_and_resize_kernel.cu)
// into oracle/_ref/libref_<op>.so.  No reference source is copied into this repository.
//
// The TH/THC/cffi host glue of the reference cannot be built any more (torch.utils.ffi is gone); what it does is restated here:
//   gpu_nms                       cuda_functions/nms_3D/src/nms_cuda.c:17-67      (mask alloc, _nms, D2H copy, serial host scan)
//   crop_and_resize_gpu_forward   cuda_functions/roi_align_3D/roi_align/src/crop_and_resize_gpu.c:7-39   (zero crops, launch)
//   crop_and_resize_gpu_backward  .../crop_and_resize_gpu.c:42-73                 (zero grads_image, launch)
// All pointers are HOST pointers; timings (ms) are returned through `times` when non-null.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return -(int)e_ - 1000; } while (0)

#if defined(REF_NMS)
extern "C" void _nms(int boxes_num, float *boxes_dev, unsigned long long *mask_dev, float nms_overlap_thresh);

// returns number kept (>= 0) or a negative error.  times[0] = mask kernel ms, times[1] = D2H ms, times[2] = host scan ms
extern "C" long long ref_nms(const float *boxes_sorted, int n, int floats_per_box, float thresh, long long *keep, double *times) {
    if (n == 0) return 0;
    const int cb = (n + 63) / 64;
    float *d_boxes = nullptr;
    unsigned long long *d_mask = nullptr;
    CK(cudaMalloc(&d_boxes, (size_t)n * floats_per_box * sizeof(float)));
    CK(cudaMalloc(&d_mask, (size_t)n * cb * sizeof(unsigned long long)));
    CK(cudaMemcpy(d_boxes, boxes_sorted, (size_t)n * floats_per_box * sizeof(float), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    cudaEventRecord(e0, 0);
    _nms(n, d_boxes, d_mask, thresh);  // legacy default stream, as in the reference
    cudaEventRecord(e1, 0);
    std::vector<unsigned long long> mask((size_t)n * cb);
    CK(cudaMemcpy(mask.data(), d_mask, mask.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaEventRecord(e2, 0);
    CK(cudaEventSynchronize(e2));
    auto t0 = std::chrono::steady_clock::now();
    std::vector<unsigned long long> remv(cb, 0ULL);
    long long kept = 0;
    for (int i = 0; i < n; ++i) {
        const int nblock = i / 64, inblock = i % 64;
        if (!(remv[nblock] & (1ULL << inblock))) {
            keep[kept++] = i;
            const unsigned long long *p = mask.data() + (size_t)i * cb;
            for (int j = nblock; j < cb; ++j) remv[j] |= p[j];
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (times) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1); times[0] = ms;
        cudaEventElapsedTime(&ms, e1, e2); times[1] = ms;
        times[2] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    cudaFree(d_boxes); cudaFree(d_mask);
    return kept;
}
// raw mask for bit-level comparison with mdt_nms_mask_*
extern "C" int ref_nms_mask(const float *boxes_sorted, int n, int floats_per_box, float thresh, unsigned long long *mask_out) {
    if (n == 0) return 0;
    const int cb = (n + 63) / 64;
    float *d_boxes = nullptr;
    unsigned long long *d_mask = nullptr;
    CK(cudaMalloc(&d_boxes, (size_t)n * floats_per_box * sizeof(float)));
    CK(cudaMalloc(&d_mask, (size_t)n * cb * sizeof(unsigned long long)));
    CK(cudaMemcpy(d_boxes, boxes_sorted, (size_t)n * floats_per_box * sizeof(float), cudaMemcpyHostToDevice));
    _nms(n, d_boxes, d_mask, thresh);
    CK(cudaMemcpy(mask_out, d_mask, (size_t)n * cb * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(d_boxes); cudaFree(d_mask);
    return 0;
}
#endif

#if defined(REF_ROI3D)
extern "C" void CropAndResizeLaucher(const float *image_ptr, const float *boxes_ptr, const int *box_ind_ptr, int num_boxes, int batch, int image_height,
                                     int image_width, int image_zdepth, int crop_height, int crop_width, int crop_zdepth, int depth,
                                     float extrapolation_value, float *crops_ptr, cudaStream_t stream);
extern "C" void CropAndResizeBackpropImageLaucher(const float *grads_ptr, const float *boxes_ptr, const int *box_ind_ptr, int num_boxes, int batch,
                                                  int image_height, int image_width, int image_zdepth, int crop_height, int crop_width,
                                                  int crop_zdepth, int depth, float *grads_image_ptr, cudaStream_t stream);
#define ZARGS(z) , z
#define ZDEF , int Z, int cz
#define ZMUL(a, z) ((a) * (size_t)(z))
#elif defined(REF_ROI2D)
extern "C" void CropAndResizeLaucher(const float *image_ptr, const float *boxes_ptr, const int *box_ind_ptr, int num_boxes, int batch, int image_height,
                                     int image_width, int crop_height, int crop_width, int depth, float extrapolation_value, float *crops_ptr,
                                     cudaStream_t stream);
extern "C" void CropAndResizeBackpropImageLaucher(const float *grads_ptr, const float *boxes_ptr, const int *box_ind_ptr, int num_boxes, int batch,
                                                  int image_height, int image_width, int crop_height, int crop_width, int depth,
                                                  float *grads_image_ptr, cudaStream_t stream);
#endif

#if defined(REF_ROI3D) || defined(REF_ROI2D)
#if defined(REF_ROI3D)
static const int kBoxF = 6;
#else
static const int kBoxF = 4;
#endif
// Z / cz are ignored (must be 1) by the 2D build.  iters > 1 repeats the launch for timing; times[0] = average kernel ms.
extern "C" int ref_crop_and_resize_forward(const float *image, const float *boxes, const int *box_ind, int num_boxes, int batch, int C, int H, int W,
                                           int Z, int ch, int cw, int cz, float *crops, int iters, double *times) {
    const size_t img_n = (size_t)batch * C * H * W * Z, crop_n = (size_t)num_boxes * C * ch * cw * cz;
    float *d_img, *d_boxes, *d_crops; int *d_ind;
    CK(cudaMalloc(&d_img, img_n * 4)); CK(cudaMalloc(&d_boxes, (size_t)num_boxes * kBoxF * 4 + 4)); CK(cudaMalloc(&d_ind, (size_t)num_boxes * 4 + 4));
    CK(cudaMalloc(&d_crops, crop_n * 4 + 4));
    CK(cudaMemcpy(d_img, image, img_n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_boxes, boxes, (size_t)num_boxes * kBoxF * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ind, box_ind, (size_t)num_boxes * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (iters < 1) iters = 1;
    CK(cudaMemset(d_crops, 0, crop_n * 4));  // crop_and_resize_gpu.c:27
    cudaEventRecord(e0, 0);
    for (int it = 0; it < iters; ++it) {
#if defined(REF_ROI3D)
        CropAndResizeLaucher(d_img, d_boxes, d_ind, num_boxes, batch, H, W, Z, ch, cw, cz, C, 0.f, d_crops, 0);
#else
        CropAndResizeLaucher(d_img, d_boxes, d_ind, num_boxes, batch, H, W, ch, cw, C, 0.f, d_crops, 0);
#endif
    }
    cudaEventRecord(e1, 0);
    CK(cudaEventSynchronize(e1));
    if (times) { float ms; cudaEventElapsedTime(&ms, e0, e1); times[0] = ms / iters; }
    CK(cudaMemcpy(crops, d_crops, crop_n * 4, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_img); cudaFree(d_boxes); cudaFree(d_ind); cudaFree(d_crops);
    return 0;
}

extern "C" int ref_crop_and_resize_backward(const float *grads, const float *boxes, const int *box_ind, int num_boxes, int batch, int C, int H, int W,
                                            int Z, int ch, int cw, int cz, float *grads_image, int iters, double *times) {
    const size_t img_n = (size_t)batch * C * H * W * Z, crop_n = (size_t)num_boxes * C * ch * cw * cz;
    float *d_img, *d_boxes, *d_grads; int *d_ind;
    CK(cudaMalloc(&d_img, img_n * 4)); CK(cudaMalloc(&d_boxes, (size_t)num_boxes * kBoxF * 4 + 4)); CK(cudaMalloc(&d_ind, (size_t)num_boxes * 4 + 4));
    CK(cudaMalloc(&d_grads, crop_n * 4 + 4));
    CK(cudaMemcpy(d_grads, grads, crop_n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_boxes, boxes, (size_t)num_boxes * kBoxF * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ind, box_ind, (size_t)num_boxes * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (iters < 1) iters = 1;
    cudaEventRecord(e0, 0);
    for (int it = 0; it < iters; ++it) {
        CK(cudaMemsetAsync(d_img, 0, img_n * 4, 0));  // crop_and_resize_gpu.c:61
#if defined(REF_ROI3D)
        CropAndResizeBackpropImageLaucher(d_grads, d_boxes, d_ind, num_boxes, batch, H, W, Z, ch, cw, cz, C, d_img, 0);
#else
        CropAndResizeBackpropImageLaucher(d_grads, d_boxes, d_ind, num_boxes, batch, H, W, ch, cw, C, d_img, 0);
#endif
    }
    cudaEventRecord(e1, 0);
    CK(cudaEventSynchronize(e1));
    if (times) { float ms; cudaEventElapsedTime(&ms, e0, e1); times[0] = ms / iters; }
    CK(cudaMemcpy(grads_image, d_img, img_n * 4, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_img); cudaFree(d_boxes); cudaFree(d_ind); cudaFree(d_grads);
    return 0;
}
#endif
