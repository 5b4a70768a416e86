/*
 * mdt_b200.h — C-ABI of libmdt_b200.so: the B200 (sm_100a) drop-in for the native ops of
 * MIC-DKFZ/medicaldetectiontoolkit's detection hot path.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no framework types; every pointer is a DEVICE pointer unless the
 *     parameter name ends in _host;
 *   - the caller owns all memory including scratch ("workspace"); query sizes with *_workspace_bytes;
 *   - work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); no host sync inside;
 *   - return value: 0 = success, <0 = MDT_E* argument error, >0 = cudaError_t of the failed launch.
 *     Nothing ever calls exit() (the reference's launchers do: roi_align_3D/.../crop_and_resize_kernel.cu:326-331).
 *
 * Each declaration cites the reference interface it replaces (paths relative to the reference root).
 */
#ifndef MDT_B200_H
#define MDT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDT_OK 0
#define MDT_EINVAL (-1)     /* bad argument (null pointer, negative size, unsupported shape) */
#define MDT_EWORKSPACE (-2) /* workspace too small */
#define MDT_EUNSUPPORTED (-3)
#define MDT_EDRIVER (-4)    /* driver entry point (tensor-map encode) unavailable or failed */

int mdt_version(void);
/* human-readable text for a return code of any entry point */
const char *mdt_error_string(int code);

/* ------------------------------------------------------------------ NMS ------------------------------------------------------------------
 * replaces: cuda_functions/nms_3D/src/cuda/nms_kernel.h:11-12  `void _nms(int boxes_num, float* boxes_dev, unsigned long long* mask_dev, float thresh)`
 *           cuda_functions/nms_3D/src/nms_cuda.h:1             `int gpu_nms(THLongTensor* keep, THLongTensor* num_out, THCudaTensor* boxes, float thresh)`
 *           (2D twins under cuda_functions/nms_2D/)
 * boxes are sorted by descending score: 3D rows = (y1,x1,y2,x2,z1,z2,score) [N,7] f32, 2D rows = (y1,x1,y2,x2,score) [N,5] f32, contiguous.
 * IoU uses +1 extents and the strict `>` of nms_kernel.cu:16-28,71.
 */

/* bitmask only — same contract as the reference `_nms` plus stream/status; mask is [N, ceil(N/64)] u64, fully written */
int mdt_nms_mask_3d(int boxes_num, const float *boxes_dev, unsigned long long *mask_dev, float nms_overlap_thresh, void *stream);
int mdt_nms_mask_2d(int boxes_num, const float *boxes_dev, unsigned long long *mask_dev, float nms_overlap_thresh, void *stream);

/* fused: bitmask + ON-DEVICE greedy reduction (replaces the D2H copy + host loop of nms_cuda.c:33-61).
 * keep_out [N] int64: first *num_out entries are the kept row indices (into the sorted input), ascending; num_out [1] int32 (device).
 * workspace: the upper-triangular mask [N][ceil(N/64)] u64 followed by the reduction's suppression bitmap and control block; the
 * reduction is a cooperative launch (<= one CTA per SM) that the call zero-initialises itself - the caller only provides the bytes. */
size_t mdt_nms_workspace_bytes(int boxes_num);
int mdt_nms_3d(const float *boxes_dev, int boxes_num, float nms_overlap_thresh, void *workspace, size_t workspace_bytes,
               int64_t *keep_out, int *num_out, void *stream);
int mdt_nms_2d(const float *boxes_dev, int boxes_num, float nms_overlap_thresh, void *workspace, size_t workspace_bytes,
               int64_t *keep_out, int *num_out, void *stream);

/* --------------------------------------------------------------- RoIAlign ----------------------------------------------------------------
 * replaces: cuda_functions/roi_align_3D/roi_align/src/cuda/crop_and_resize_kernel.h:8-18 (CropAndResizeLaucher / CropAndResizeBackpropImageLaucher)
 *           and the 2D twins. Same argument order; added: element strides of the image (so channels-last maps need no copy) and int status.
 * image  logical [batch, depth(C), H, W, Z] f32 with element strides img_strides[5] (2D: [batch, C, H, W], img_strides[4])
 * boxes  [num_boxes, 6] f32 normalised (y1,x1,y2,x2,z1,z2)   (2D: [num_boxes,4])
 * box_ind[num_boxes] int32 in [0,batch); out-of-range => that crop stays zero (crop_and_resize_kernel.cu:43-47 + crop_and_resize_gpu.c:27)
 * crops  logical [num_boxes, C, ch, cw, cz] f32 with element strides crop_strides[5]; fully written (zeros for bad box_ind).
 * extrapolation_value is accepted and ignored, as in the reference GPU kernel.
 */
int mdt_crop_and_resize_3d_forward(const float *image, const int64_t *img_strides_host, const float *boxes, const int *box_ind, int num_boxes,
                                   int batch, int image_height, int image_width, int image_zdepth, int crop_height, int crop_width,
                                   int crop_zdepth, int depth, float extrapolation_value, float *crops, const int64_t *crop_strides_host,
                                   void *stream);
/* grads_image must be zero-filled by the caller OR pass zero_init=1 to have the library memset it (contiguous-in-memory extent given by image_numel) */
int mdt_crop_and_resize_3d_backward(const float *grads, const int64_t *grad_strides_host, const float *boxes, const int *box_ind, int num_boxes,
                                    int batch, int image_height, int image_width, int image_zdepth, int crop_height, int crop_width,
                                    int crop_zdepth, int depth, float *grads_image, const int64_t *img_strides_host, int zero_init,
                                    int64_t image_numel, void *stream);
int mdt_crop_and_resize_2d_forward(const float *image, const int64_t *img_strides_host, const float *boxes, const int *box_ind, int num_boxes,
                                   int batch, int image_height, int image_width, int crop_height, int crop_width, int depth,
                                   float extrapolation_value, float *crops, const int64_t *crop_strides_host, void *stream);
int mdt_crop_and_resize_2d_backward(const float *grads, const int64_t *grad_strides_host, const float *boxes, const int *box_ind, int num_boxes,
                                    int batch, int image_height, int image_width, int crop_height, int crop_width, int depth,
                                    float *grads_image, const int64_t *img_strides_host, int zero_init, int64_t image_numel, void *stream);

/* pyramid RoIAlign: ONE launch over all FPN levels.  replaces: models/mrcnn.py:373-457 pyramid_roi_align's per-level loop (boolean-mask gather,
 * up to 4 CropAndResize calls, concat, un-permute through a sort).  roi_level[n] in [0, nlevels) selects the map of RoI n (the reference's
 * round(4 + log2(sqrt(h*w))) rule is evaluated by the caller); every crop row is written exactly once, bad box_ind rows are zeros.
 * images / grad_images: HOST arrays of nlevels device pointers (channels-last maps, C % 4 == 0); image_strides: nlevels x 5 int64 element strides
 * (b, c, y, x, z); image_dims: nlevels x 3 (H, W, Z).  Returns MDT_EUNSUPPORTED for layouts outside the vector kernel (caller then loops over
 * levels with mdt_crop_and_resize_*).  dim = 2 or 3 (2: crop_zdepth and the Z entries are ignored). */
int mdt_pyramid_roi_align_forward(int dim, const float *const *images, const int64_t *image_strides, const int *image_dims, int nlevels, const float *boxes,
                                  const int *box_ind, const int *roi_level, int num_boxes, int batch, int crop_height, int crop_width, int crop_zdepth,
                                  int depth, float *crops, const int64_t *crop_strides, void *stream);
int mdt_pyramid_roi_align_backward(int dim, const float *grads, const int64_t *grad_strides, const float *boxes, const int *box_ind, const int *roi_level,
                                   int num_boxes, int batch, int crop_height, int crop_width, int crop_zdepth, int depth, float *const *grad_images,
                                   const int64_t *image_strides, const int *image_dims, int nlevels, int zero_init, const int64_t *image_numel, void *stream);

/* ------------------------------------------------------------ anchor <-> GT matching --------------------------------------------------------
 * replaces: utils/model_utils.py:505-619 gt_anchor_matching (+ compute_overlaps :83-110, compute_iou_{2D,3D} :35-79), which runs in numpy f64 on the host.
 * anchors [A, 2*dim] f64, gt_boxes [G, 2*dim] f64, gt_class_ids [G] int32 (NULL => all 1, the RPN case of mrcnn.py:894).
 * Outputs (device):
 *   matches   [A] int32   : -1 negative / 0 neutral / class id positive — state BEFORE the random sub-sampling of model_utils.py:566-571
 *   iou_argmax[A] int32   : row argmax (first index on ties, numpy semantics) — needed by the caller for delta targets
 *   n_pos     [1] int32   : number of positive anchors
 * neg_iou_thresh is 0.01 (3D) / 0.1 (2D) in the reference (:549-552); pos_iou_thresh = cf.anchor_matching_iou.
 * workspace: mdt_anchor_match_workspace_bytes(G).
 */
size_t mdt_anchor_match_workspace_bytes(int num_gt);
int mdt_anchor_match(int dim, const double *anchors, int num_anchors, const double *gt_boxes, const int *gt_class_ids, int num_gt,
                     double neg_iou_thresh, double pos_iou_thresh, void *workspace, size_t workspace_bytes, int *matches, int *iou_argmax,
                     int *n_pos, void *stream);
/* delta targets for up to max_targets positive anchors in ascending anchor order (model_utils.py:573-617): out [max_targets, 2*dim] f64, zero padded.
 * pos_ids [n_pos] int32 ascending anchor indices (after any sub-sampling done by the caller). std_dev [2*dim] f64 (host). */
int mdt_anchor_delta_targets(int dim, const double *anchors, const double *gt_boxes, const int *iou_argmax, const int *pos_ids, int n_pos,
                             int max_targets, const double *std_dev_host, double *targets_out, void *stream);

/* ------------------------------------------------------------------ conv3d ----------------------------------------------------------------
 * replaces: the nn.Conv3d (+bias, +ReLU) built by utils/model_utils.py:739-781 NDConvGenerator and consumed by models/backbone.py:27-206.
 * Layout: activations NDHWC ("channels_last_3d"): x [N, D, H, W, C] f32 contiguous, where (D,H,W) are the reference's (y,x,z) axes.
 *         weights in the reference's parameter layout [Cout, Cin, kd, kh, kw] f32 contiguous (state-dict compatible).
 * Descriptor struct keeps the signature stable across kernels.
 */
typedef struct mdt_conv3d_desc {
    int n, d, h, w;          /* input batch and spatial extent */
    int cin, cout;
    int kd, kh, kw;
    int sd, sh, sw;          /* strides */
    int pd, ph, pw;          /* zero padding */
    int relu;                /* fuse ReLU into the fprop epilogue */
    int precision;           /* 0 = fp32-faithful (3-pass split-bf16 on tcgen05, or fp32 SIMT); 1 = single-pass bf16 (throughput mode) */
    int algo;                /* 0 = auto, 1 = force SIMT direct kernels, 2 = force tcgen05 implicit GEMM, 4 = force pointwise streaming (error if unsupported) */
} mdt_conv3d_desc;

size_t mdt_conv3d_workspace_bytes(const mdt_conv3d_desc *desc_host, int pass /*0 fprop,1 dgrad,2 wgrad*/);
/* y = relu?(conv(x, w) + bias); bias may be NULL; residual (same shape as y, may be NULL) is added before the ReLU */
int mdt_conv3d_fprop(const mdt_conv3d_desc *desc_host, const float *x, const float *w, const float *bias, const float *residual, float *y,
                     void *workspace, size_t workspace_bytes, void *stream);
/* dx = conv_transpose(dy, w) */
int mdt_conv3d_dgrad(const mdt_conv3d_desc *desc_host, const float *dy, const float *w, float *dx, void *workspace, size_t workspace_bytes,
                     void *stream);
/* dw = x (*) dy ; db = sum(dy) (db may be NULL) */
int mdt_conv3d_wgrad(const mdt_conv3d_desc *desc_host, const float *x, const float *dy, float *dw, float *db, void *workspace,
                     size_t workspace_bytes, void *stream);
/* Fused backward of one conv layer (tcgen05 path): ONE streaming pass over dy yields the split-bf16 operand shared by dgrad and wgrad, applies
 * the ReLU mask of a fused-ReLU conv (y_relu = its forward output, NULL = no mask; replaces aten::threshold_backward of the reference graph),
 * accumulates db, and optionally writes the masked fp32 gradient (dy_masked_out, needed as the gradient of a fused residual input).
 * dx may be NULL (first layer).  mdt_conv3d_backward_fused() tells whether this path applies; otherwise use dgrad + wgrad.
 * x_split (optional, else NULL and x is split internally): the canonical split form of x produced by mdt_conv3d_split. */
int mdt_conv3d_backward_fused(const mdt_conv3d_desc *desc_host, int need_dx);
size_t mdt_conv3d_backward_workspace_bytes(const mdt_conv3d_desc *desc_host, int need_dx);
int mdt_conv3d_backward(const mdt_conv3d_desc *desc_host, const float *x, const void *x_split, const float *dy, const float *y_relu,
                        const float *w, float *dx, float *dw, float *db, float *dy_masked_out, void *workspace, size_t workspace_bytes,
                        void *stream);
/* Canonical split form of an activation tensor (two bf16 planes, channels padded, W-line interleaved): it depends on the tensor alone, so ONE
 * split serves every conv that reads the tensor (sibling layers, fprop and the later weight gradient).  mdt_conv3d_split_bytes gives the buffer
 * size for the INPUT x of `desc`; mdt_conv3d_fprop_presplit is mdt_conv3d_fprop on such a buffer (tcgen05 path only, else MDT_EUNSUPPORTED). */
size_t mdt_conv3d_split_bytes(const mdt_conv3d_desc *desc_host);
int mdt_conv3d_split(const mdt_conv3d_desc *desc_host, const float *x, void *x_split, void *stream);
int mdt_conv3d_fprop_presplit(const mdt_conv3d_desc *desc_host, const void *x_split, const float *w, const float *bias, const float *residual,
                              float *y, void *workspace, size_t workspace_bytes, void *stream);
/* mdt_conv3d_fprop_presplit that ALSO emits the result y in canonical split form (mdt_conv3d_out_split_bytes bytes; layout of mdt_conv3d_split
 * applied to y, padding channels zero) from the conv epilogue: the consumer conv then needs no split pass over y (tcgen05 path only). */
size_t mdt_conv3d_out_split_bytes(const mdt_conv3d_desc *desc_host);
int mdt_conv3d_fprop_presplit_out(const mdt_conv3d_desc *desc_host, const void *x_split, const float *w, const float *bias, const float *residual,
                                  float *y, void *y_split, void *workspace, size_t workspace_bytes, void *stream);
/* fp32-input fprop that also emits the split planes of y (as mdt_conv3d_fprop_presplit_out does for a pre-split input): pointwise path only
 * (mdt_conv3d_algo == 4), else MDT_EUNSUPPORTED.  Used where a 1x1x1 conv feeds a tcgen05 conv (ResBlock conv1 -> conv2, P0_conv1 -> P0_conv2). */
int mdt_conv3d_fprop_out(const mdt_conv3d_desc *desc_host, const float *x, const float *w, const float *bias, const float *residual, float *y,
                         void *y_split, void *workspace, size_t workspace_bytes, void *stream);
/* which algorithm `auto` resolves to for this descriptor/pass: 1 SIMT, 2 tcgen05, 4 pointwise fp32 streaming (1x1x1, stride 1, no padding,
 * cin * cout <= 2592 under `auto` (where it measured faster than tcgen05), <= 6144 when forced with algo = 4: HBM-bound layers, csrc/conv3d_pw.cu) */
int mdt_conv3d_algo(const mdt_conv3d_desc *desc_host, int pass);
/* which kernel family runs this descriptor/pass: 1 fp32 SIMT / direct stem, 2 tcgen05 halo-window implicit GEMM (conv3d_tc.cu, wgrad:
 * conv3d_tc_wgrad.cu), 3 tcgen05 tap-stacked implicit GEMM (conv3d_tcw.cu: fprop/dgrad of lines of 65..128 voxels), 4 pointwise fp32 streaming (conv3d_pw.cu); 0 = unsupported */
int mdt_conv3d_variant(const mdt_conv3d_desc *desc_host, int pass);
/* diagnostics: cycle counters of CTA 0 of the tap-stacked kernel, accumulated while the environment has MDT_TCW_PROF=1 (16 values: producer
 * wait-A / wait-B / total, MMA wait-A / wait-B / wait-accumulator / total, epilogue wait / total, tiles); synchronises the device and clears them */
int mdt_debug_conv_tcw_prof(unsigned long long *out16_host);
/* ------------------------------------------------------------- decoder up-sampling -----------------------------------------------------------
 * replaces: F.interpolate(x, scale_factor=(2,2,1), mode='trilinear', align_corners=False) of models/backbone.py:209-218 (P2/P1_upsample), NDHWC.
 * x [n, d, h, w, c] -> y [n, 2d, 2h, w, c]; c % 4 == 0; backward is the exact adjoint, written as a gather (no atomics, no zero fill). */
int mdt_upsample221_forward(const float *x, float *y, int n, int d, int h, int w, int c, void *stream);
int mdt_upsample221_backward(const float *gy, float *gx, int n, int d, int h, int w, int c, void *stream);

/* ------------------------------------------------------------- pooling / nearest up-sampling -------------------------------------------------
 * replaces: nn.MaxPool3d(kernel_size=3, stride=(2,2,1), padding=1) / nn.MaxPool2d(3, 2, 1) of models/backbone.py:63-64 (applied at :129) and its
 * autograd backward.  NDHWC maps (2D: d = 1, kernel {1,3,3}).  x [n,d,h,w,c] -> y [n,od,oh,ow,c] with o = (i + 2p - k) / s + 1 (floor mode);
 * argmax [n,od,oh,ow,c] holds the winning window offset ((a*kh + b)*kw + e) per element, ATen's update rule (first maximum, NaN propagates).
 * Backward is a gather over the windows containing an input voxel: no atomics, gx fully written. */
int mdt_maxpool3d_forward(const float *x, float *y, unsigned char *argmax, int n, int d, int h, int w, int c, const int *kernel3, const int *stride3,
                          const int *pad3, void *stream);
int mdt_maxpool3d_backward(const float *gy, const unsigned char *argmax, float *gx, int n, int d, int h, int w, int c, const int *kernel3,
                           const int *stride3, const int *pad3, void *stream);
/* replaces: F.interpolate(x, scale_factor=2) (mode 'nearest') of the FPN top-down path, models/backbone.py:147-153.  x [n,d,h,w,c] ->
 * y [n,d*fd,h*fh,w*fw,c], factors 1 or 2 per axis; backward sums the children of each input voxel (gather, gx fully written). */
int mdt_upsample_nearest_forward(const float *x, float *y, int n, int d, int h, int w, int c, int fd, int fh, int fw, void *stream);
int mdt_upsample_nearest_backward(const float *gy, float *gx, int n, int d, int h, int w, int c, int fd, int fh, int fw, void *stream);

/* ------------------------------------------------------------- loss-side kernels (csrc/loss_ops.cu) ------------------------------------------
 * Segmentation loss.  replaces: F.softmax(seg_logits, 1) + get_one_hot_encoding + batch_dice (utils/model_utils.py:785-799, 833-858) and
 * F.cross_entropy(seg_logits, seg[:, 0]) of models/retina_unet.py:395,446-448 (same calls in models/mrcnn.py is n/a; ufrcnn.py uses them too).
 * logits[b, c, v] = logits + b*strides3[0] + c*strides3[1] + v*strides3[2] (elements; NCDHW: {C*V, V, 1}, channels-last: {V*C, 1, C});
 * target uint8 [n, voxels]; n_classes <= 8.  sums (device, 3*n_classes + 1 doubles) = per-class intersect, sum p, sum y, then sum(-log p_t) —
 * saved for the backward; out2 (device) = {dice score (mean over foreground classes, batch pseudo-volume), mean cross-entropy}.
 * Deterministic (fixed-order fp64 reduction).  backward: grad_out2 (device) = {dL/d dice_score, dL/d ce}; grad_logits has logits' strides. */
size_t mdt_seg_loss_workspace_bytes(int n_classes);
int mdt_seg_loss_forward(const float *logits, const long long *strides3, const unsigned char *target, int n, long long voxels, int n_classes,
                         float false_positive_weight, float smooth, double *sums, float *out2, void *ws, size_t ws_bytes, void *stream);
int mdt_seg_loss_backward(const float *logits, const long long *strides3, const unsigned char *target, int n, long long voxels, int n_classes,
                          float false_positive_weight, float smooth, const double *sums, const float *grad_out2, float *grad_logits, void *stream);
/* Class loss with stochastic hard-example mining.  replaces: compute_class_loss (models/retina_unet.py:126-164) / compute_rpn_class_loss
 * (models/mrcnn.py:176-213) incl. mutils.shem (utils/model_utils.py:674-691): softmax over all anchors, sort of the negatives' max foreground
 * probability, randperm sample, two cross-entropies.  logits [n_anchors, n_classes] contiguous fp32, matches int32 (-1 negative, 0 neutral,
 * > 0 positive class); pos_ids = the first n_pos_list positive anchor ids (int64, ascending; at most k_pos are used); k_pool = size of the
 * candidate pool (<= 1024, = shem_poolsize * k_pos), k_neg <= k_pool sampled negatives at most; rand_keys [k_pool] uniform [0,1) floats from the
 * caller's generator (the negative_count = max(1, #positives) pool members with the smallest keys are the sample).
 * Outputs (device): loss[1] = (CE_pos + CE_neg) / 2; neg_ix [k_neg] int64 = rank of each sampled negative inside the negative subset (-1 pad);
 * sel_rows / sel_labels / sel_w [k_pos + k_neg] = the rows entering the loss, their labels and weights (for the backward).
 * backward: grad_logits [n_anchors, n_classes] is zero-filled, then rows += grad_loss * w * (softmax - onehot). */
size_t mdt_shem_workspace_bytes(int n_anchors, int k_pool);
int mdt_shem_class_loss_forward(const float *logits, const int *matches, int n_anchors, int n_classes, const long long *pos_ids, int n_pos_list, int k_pos,
                                int k_pool, int k_neg, int shem_poolsize, const float *rand_keys, float *loss, long long *neg_ix, int *sel_rows,
                                int *sel_labels, float *sel_w, void *ws, size_t ws_bytes, void *stream);
int mdt_shem_class_loss_backward(const float *logits, int n_anchors, int n_classes, const int *sel_rows, const int *sel_labels, const float *sel_w,
                                 int n_sel, const float *grad_loss, float *grad_logits, void *stream);

/* ------------------------------------------------------------- inference-side consolidation (csrc/consolidate.cu) ----------------------------
 * replaces: weighted_box_clustering(dets, box_patch_id, thresh, n_ens) of predictor.py:597-706 (numpy loop per patient and class).
 * dets [n, 2*dim + 3] fp64 rows = (y1, x1, y2, x2, (z1, z2), score, patch-centre factor, number of overlapping patches); patch_id [n] int32 in
 * [0, n_patches) (the caller maps the reference's patch-id strings to dense ints); order [n] int32 = box indices by descending score.
 * Outputs (device): keep_scores [<= n], keep_coords [<= n, 2*dim] in cluster order, n_keep[1].  One resident CTA runs the whole greedy loop. */
size_t mdt_wbc_workspace_bytes(int n, int n_patches);
int mdt_wbc(const double *dets, const int *patch_id, const int *order, int n, int dim, int n_patches, double thresh, double n_ens, double *keep_scores,
            double *keep_coords, int *n_keep, void *ws, size_t ws_bytes, void *stream);
/* replaces: nms_2to3D(dets, thresh) of predictor.py:710-773.  dets [n, 6] fp64 rows = (y1, x1, y2, x2, score, slice id), slice ids in
 * [0, n_slices); order as above.  Outputs: keep [<= n] int64 indices of the cluster cores, keep_z [<= n, 2] = (z1, z2), n_keep[1]. */
size_t mdt_nms_2to3d_workspace_bytes(int n, int n_slices);
int mdt_nms_2to3d(const double *dets, const int *order, int n, double thresh, int n_slices, long long *keep, double *keep_z, int *n_keep, void *ws,
                  size_t ws_bytes, void *stream);

/* number of kernel launches issued by this library since load (all entry points) — feeds bench.py's gpu_launches */
unsigned long long mdt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MDT_B200_H */
