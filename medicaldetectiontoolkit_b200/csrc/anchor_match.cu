// Anchor <-> ground-truth IoU matching on the device (fp64), replacing the host-side numpy routine of the reference.
//
// Reference semantics followed (paths relative to the reference root):
//   IoU                utils/model_utils.py:35-79   compute_iou_2D / compute_iou_3D  (no +1 extents, fp64, op order preserved below)
//   overlaps matrix    utils/model_utils.py:83-110  compute_overlaps                 (volumes: (h*w)*d)
//   labelling          utils/model_utils.py:505-563 gt_anchor_matching steps 1-3:
//                        1. row max < neg_thresh            -> -1
//                        2. every GT's best anchor          -> its class (np.argmax over axis 0: first anchor on ties; later GT overwrites)
//                        3. row max >= anchor_matching_iou  -> class of the row argmax (np.argmax over axis 1: first GT on ties)
//   delta targets      utils/model_utils.py:573-617
// Every fp64 operation uses an explicit round-to-nearest intrinsic so that no FMA contraction can change a comparison result:
// labels are compared bit-for-bit with numpy.
//
// Kernels: (1) row pass — one thread per anchor loops over the GT boxes held in shared memory, keeps the running row max/argmax and
// contributes to the per-GT column maximum through a warp-shuffle max followed by ONE 64-bit atomicMax per warp per GT;
// (2) column-argmin pass — anchors whose IoU equals the column maximum race with atomicMin on the anchor index (first index wins);
// (3) finalise — labels, positive count. The A x G matrix is never materialised (the reference allocates it: 86 MB at A = 1.35 M, G = 8).
#include "mdt_common.cuh"

namespace mdt {

constexpr int kMaxGtSmem = 512;  // GT boxes staged per shared-memory chunk

__device__ __forceinline__ unsigned long long ordered_key(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}

template <int DIM>
__device__ __forceinline__ double box_volume(const double *b) {
    double v = __dmul_rn(__dsub_rn(b[2], b[0]), __dsub_rn(b[3], b[1]));
    if (DIM == 3) v = __dmul_rn(v, __dsub_rn(b[5], b[4]));
    return v;
}

// iou of GT box g (volume vg) with anchor a (volume va), op order of compute_iou_{2D,3D}
template <int DIM>
__device__ __forceinline__ double iou_f64(const double *g, double vg, const double *a, double va) {
    const double y1 = fmax(g[0], a[0]), y2 = fmin(g[2], a[2]);
    const double x1 = fmax(g[1], a[1]), x2 = fmin(g[3], a[3]);
    double inter = __dmul_rn(fmax(__dsub_rn(x2, x1), 0.0), fmax(__dsub_rn(y2, y1), 0.0));
    if (DIM == 3) {
        const double z1 = fmax(g[4], a[4]), z2 = fmin(g[5], a[5]);
        inter = __dmul_rn(inter, fmax(__dsub_rn(z2, z1), 0.0));
    }
    const double uni = __dsub_rn(__dadd_rn(vg, va), inter);
    // Disjoint boxes (almost every anchor/GT pair): 0 / uni is exactly +0.0 for uni > 0, so the ~30-instruction fp64 division is
    // skipped. Degenerate unions (<= 0 or NaN) still take the division so that they produce numpy's -0.0 / NaN / inf bit patterns.
    if (inter == 0.0 && uni > 0.0) return 0.0;
    return __ddiv_rn(inter, uni);
}

// A box is "regular" when every extent is > 0 and its volume is positive and finite.  For two regular boxes the intersection is
// exactly 0 iff they are separated along some axis (min(hi) <= max(lo), decided by comparisons alone: fp64 subtraction of finite
// numbers is 0 only for equal operands), and then numpy's IoU is 0 / (vg + va) = +0.0.  ~200 instructions of NaN-aware fp64 min/max,
// products and a division shrink to 2*DIM comparisons for the overwhelmingly common disjoint pair (ncu: the row pass was issue-bound).
template <int DIM>
__device__ __forceinline__ bool box_regular(const double *b, double vol) {
    bool ok = b[2] > b[0] && b[3] > b[1] && vol > 0.0 && vol < __longlong_as_double(0x7ff0000000000000LL);
    if (DIM == 3) ok = ok && b[5] > b[4];
    return ok;
}
template <int DIM>
__device__ __forceinline__ bool separated(const double *g, const double *a) {
    bool d = (g[2] <= a[0]) | (a[2] <= g[0]) | (g[3] <= a[1]) | (a[3] <= g[1]);
    if (DIM == 3) d = d | (g[5] <= a[4]) | (a[5] <= g[4]);
    return d;
}

// stage GT boxes [g0, g0+gn) with their volumes and regularity flags in shared memory
template <int DIM>
__device__ __forceinline__ void stage_gt(const double *__restrict__ gt, int g0, int gn, double *s_gt, unsigned char *s_ok) {
    constexpr int B = 2 * DIM;
    for (int i = threadIdx.x; i < gn; i += blockDim.x) {
        double gb[B];
#pragma unroll
        for (int k = 0; k < B; ++k) { gb[k] = gt[(size_t)(g0 + i) * B + k]; s_gt[i * (B + 1) + k] = gb[k]; }
        const double vg = box_volume<DIM>(gb);
        s_gt[i * (B + 1) + B] = vg;
        s_ok[i] = box_regular<DIM>(gb, vg) ? 1 : 0;
    }
}

template <int DIM>
__global__ void __launch_bounds__(256) match_rows_kernel(const double *__restrict__ anchors, int A, const double *__restrict__ gt, int G,
                                                        int *__restrict__ row_argmax, unsigned long long *__restrict__ col_max_key) {
    constexpr int B = 2 * DIM;
    __shared__ double s_gt[kMaxGtSmem * (B + 1)];
    __shared__ unsigned long long s_colmax[kMaxGtSmem];
    __shared__ unsigned char s_ok[kMaxGtSmem];
    const unsigned long long key_zero = ordered_key(0.0);   // column maxima start at IoU 0 (every IoU is >= 0): zero overlaps never touch an atomic
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = a < A;
    double box[B];
    double va = 0.0;
    if (live) {
#pragma unroll
        for (int k = 0; k < B; ++k) box[k] = anchors[(size_t)a * B + k];
        va = box_volume<DIM>(box);
    }
    const bool a_ok = live && box_regular<DIM>(box, va);
    double best = 0.0;
    int best_g = 0;
    for (int g0 = 0; g0 < G; g0 += kMaxGtSmem) {
        const int gn = min(G - g0, kMaxGtSmem);
        __syncthreads();
        stage_gt<DIM>(gt, g0, gn, s_gt, s_ok);
        for (int i = threadIdx.x; i < gn; i += blockDim.x) s_colmax[i] = key_zero;
        __syncthreads();
        for (int i = 0; i < gn; ++i) {
            const double *gb = s_gt + i * (B + 1);
            double v = live ? 0.0 : -1.0;
            if (live && !(a_ok && s_ok[i] && separated<DIM>(gb, box))) v = iou_f64<DIM>(gb, gb[B], box, va);
            if (live && (g0 + i == 0 || v > best)) { best = v; best_g = g0 + i; }  // first index on ties (np.argmax axis=1)
            // column maximum: warp max, then one atomic per warp. Columns start at IoU 0, so a warp in which no lane overlaps this GT
            // (the common case) has nothing to contribute and skips the shuffle tree.
            unsigned long long key = live ? ordered_key(v) : 0ULL;
            if (!__any_sync(0xffffffffu, key > key_zero)) continue;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            if ((threadIdx.x & 31) == 0 && key > key_zero) atomicMax(&s_colmax[i], key);   // warp max -> block max (shared memory)
        }
        __syncthreads();
        for (int i = threadIdx.x; i < gn; i += blockDim.x)
            if (s_colmax[i] > key_zero) atomicMax(col_max_key + g0 + i, s_colmax[i]);             // one global atomic per block and GT at most
    }
    if (live) row_argmax[a] = best_g;
}

template <int DIM>
__global__ void __launch_bounds__(256) match_cols_kernel(const double *__restrict__ anchors, int A, const double *__restrict__ gt, int G,
                                                        const unsigned long long *__restrict__ col_max_key, int *__restrict__ col_argmax) {
    constexpr int B = 2 * DIM;
    __shared__ double s_gt[kMaxGtSmem * (B + 1)];
    __shared__ unsigned long long s_cm[kMaxGtSmem];
    __shared__ unsigned char s_ok[kMaxGtSmem];
    const unsigned long long key_zero = ordered_key(0.0);
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = a < A;
    double box[B];
    double va = 0.0;
    if (live) {
#pragma unroll
        for (int k = 0; k < B; ++k) box[k] = anchors[(size_t)a * B + k];
        va = box_volume<DIM>(box);
    }
    const bool a_ok = live && box_regular<DIM>(box, va);
    for (int g0 = 0; g0 < G; g0 += kMaxGtSmem) {
        const int gn = min(G - g0, kMaxGtSmem);
        __syncthreads();
        stage_gt<DIM>(gt, g0, gn, s_gt, s_ok);
        for (int i = threadIdx.x; i < gn; i += blockDim.x) s_cm[i] = col_max_key[g0 + i];
        __syncthreads();
        if (!live) continue;
        for (int i = 0; i < gn; ++i) {
            // first anchor on ties (np.argmax axis=0).  A column whose maximum is 0 is won by anchor 0 by definition: resolved in finalize,
            // so a pair with IoU +0.0 (every separated pair) can never be a winner here.
            const unsigned long long cm = s_cm[i];
            if (cm == key_zero) continue;
            const double *gb = s_gt + i * (B + 1);
            if (a_ok && s_ok[i] && separated<DIM>(gb, box)) continue;
            const double v = iou_f64<DIM>(gb, gb[B], box, va);
            if (ordered_key(v) == cm) atomicMin(col_argmax + g0 + i, a);
        }
    }
}

template <int DIM>
__global__ void __launch_bounds__(256) match_finalize_kernel(const double *__restrict__ anchors, int A, const double *__restrict__ gt,
                                                            const int *__restrict__ gt_cls, int G, double neg_t, double pos_t,
                                                            const int *__restrict__ row_argmax, const int *__restrict__ col_argmax,
                                                            const unsigned long long *__restrict__ col_max_key, int *__restrict__ matches,
                                                            int *__restrict__ n_pos) {
    constexpr int B = 2 * DIM;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    int label = 0;
    if (a < A) {
        double box[B], gb[B];
#pragma unroll
        for (int k = 0; k < B; ++k) box[k] = anchors[(size_t)a * B + k];
        const int ga = row_argmax[a];
#pragma unroll
        for (int k = 0; k < B; ++k) gb[k] = __ldg(gt + (size_t)ga * B + k);
        const double mx = iou_f64<DIM>(gb, box_volume<DIM>(gb), box, box_volume<DIM>(box));
        if (mx < neg_t) label = -1;                                       // step 1
        for (int g = 0; g < G; ++g)                                        // step 2, ascending: later GT overwrites
            if ((__ldg(col_max_key + g) == ordered_key(0.0) ? 0 : __ldg(col_argmax + g)) == a) label = gt_cls ? __ldg(gt_cls + g) : 1;
        if (mx >= pos_t) label = gt_cls ? __ldg(gt_cls + ga) : 1;         // step 3
        matches[a] = label;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, label > 0);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_pos, __popc(ballot));
}

__global__ void match_init_kernel(unsigned long long *col_max_key, int *col_argmax, int G, int *n_pos) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) { col_max_key[g] = ordered_key(0.0); col_argmax[g] = 0x7fffffff; }
    if (g == 0) *n_pos = 0;
}

__global__ void fill_int_kernel(int *p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct StdDev6 { double v[6]; };
template <int DIM>
__global__ void delta_targets_byval_kernel(const double *__restrict__ anchors, const double *__restrict__ gt, const int *__restrict__ row_argmax,
                                           const int *__restrict__ pos_ids, int n_pos, int max_targets, StdDev6 sd, double *__restrict__ out) {
    constexpr int B = 2 * DIM;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= max_targets) return;
    double t[B];
#pragma unroll
    for (int j = 0; j < B; ++j) t[j] = 0.0;
    if (k < n_pos) {
        const int a_ix = pos_ids[k];
        const double *a = anchors + (size_t)a_ix * B;
        const double *g = gt + (size_t)row_argmax[a_ix] * B;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            const int lo = (ax < 2) ? ax : 4, hi = (ax < 2) ? ax + 2 : 5;
            const double g_e = __dsub_rn(g[hi], g[lo]), a_e = __dsub_rn(a[hi], a[lo]);
            const double g_c = __dadd_rn(g[lo], __dmul_rn(0.5, g_e)), a_c = __dadd_rn(a[lo], __dmul_rn(0.5, a_e));
            t[ax] = __ddiv_rn(__ddiv_rn(__dsub_rn(g_c, a_c), a_e), sd.v[ax]);
            t[DIM + ax] = __ddiv_rn(log(__ddiv_rn(g_e, a_e)), sd.v[DIM + ax]);
        }
    }
#pragma unroll
    for (int j = 0; j < B; ++j) out[(size_t)k * B + j] = t[j];
}

template <int DIM>
static int anchor_match_impl(const double *anchors, int A, const double *gt, const int *gt_cls, int G, double neg_t, double pos_t, void *ws,
                             int *matches, int *row_argmax, int *n_pos, cudaStream_t st) {
    auto *col_max_key = reinterpret_cast<unsigned long long *>(ws);
    int *col_argmax = reinterpret_cast<int *>(col_max_key + G);
    const int blocks = ceil_div(A, 256);
    match_init_kernel<<<ceil_div(G, 256), 256, 0, st>>>(col_max_key, col_argmax, G, n_pos);
    int rc = launch_status();
    if (rc) return rc;
    match_rows_kernel<DIM><<<blocks, 256, 0, st>>>(anchors, A, gt, G, row_argmax, col_max_key);
    if ((rc = launch_status())) return rc;
    match_cols_kernel<DIM><<<blocks, 256, 0, st>>>(anchors, A, gt, G, col_max_key, col_argmax);
    if ((rc = launch_status())) return rc;
    match_finalize_kernel<DIM><<<blocks, 256, 0, st>>>(anchors, A, gt, gt_cls, G, neg_t, pos_t, row_argmax, col_argmax, col_max_key, matches, n_pos);
    return launch_status();
}

}  // namespace mdt

extern "C" {

size_t mdt_anchor_match_workspace_bytes(int num_gt) {
    if (num_gt <= 0) return 16;
    return (size_t)num_gt * (sizeof(unsigned long long) + sizeof(int)) + 16;
}

int mdt_anchor_match(int dim, const double *anchors, int A, const double *gt, const int *gt_cls, int G, double neg_t, double pos_t, void *ws,
                     size_t ws_bytes, int *matches, int *row_argmax, int *n_pos, void *stream) {
    using namespace mdt;
    cudaStream_t st = as_stream(stream);
    if ((dim != 2 && dim != 3) || A < 0 || G < 0 || !n_pos) return MDT_EINVAL;
    if (A > 0 && (!anchors || !matches || !row_argmax)) return MDT_EINVAL;
    if (G == 0) {  // gt_boxes is None: every anchor negative (model_utils.py:525-527)
        cudaError_t e = cudaMemsetAsync(n_pos, 0, sizeof(int), st);
        if (e != cudaSuccess) return (int)e;
        if (A == 0) return MDT_OK;
        fill_int_kernel<<<ceil_div(A, 256), 256, 0, st>>>(matches, A, -1);
        int rc = launch_status();
        if (rc) return rc;
        e = cudaMemsetAsync(row_argmax, 0, (size_t)A * sizeof(int), st);
        return e == cudaSuccess ? MDT_OK : (int)e;
    }
    if (!gt || !ws) return MDT_EINVAL;
    if (ws_bytes < mdt_anchor_match_workspace_bytes(G)) return MDT_EWORKSPACE;
    if (A == 0) {
        cudaError_t e = cudaMemsetAsync(n_pos, 0, sizeof(int), st);
        return e == cudaSuccess ? MDT_OK : (int)e;
    }
    return dim == 3 ? anchor_match_impl<3>(anchors, A, gt, gt_cls, G, neg_t, pos_t, ws, matches, row_argmax, n_pos, st)
                    : anchor_match_impl<2>(anchors, A, gt, gt_cls, G, neg_t, pos_t, ws, matches, row_argmax, n_pos, st);
}

int mdt_anchor_delta_targets(int dim, const double *anchors, const double *gt, const int *row_argmax, const int *pos_ids, int n_pos,
                             int max_targets, const double *std_dev_host, double *out, void *stream) {
    using namespace mdt;
    if ((dim != 2 && dim != 3) || max_targets < 0 || n_pos < 0 || !std_dev_host) return MDT_EINVAL;
    if (max_targets == 0) return MDT_OK;
    if (!out || (n_pos > 0 && (!anchors || !gt || !row_argmax || !pos_ids))) return MDT_EINVAL;
    if (n_pos > max_targets) return MDT_EINVAL;  // the reference would index past anchor_delta_targets here
    StdDev6 sd;  // passed by value as a kernel argument: no device staging, no sync
    for (int k = 0; k < 6; ++k) sd.v[k] = k < 2 * dim ? std_dev_host[k] : 1.0;
    const int blocks = ceil_div(max_targets, 64);
    cudaStream_t st = as_stream(stream);
    if (dim == 3) delta_targets_byval_kernel<3><<<blocks, 64, 0, st>>>(anchors, gt, row_argmax, pos_ids, n_pos, max_targets, sd, out);
    else          delta_targets_byval_kernel<2><<<blocks, 64, 0, st>>>(anchors, gt, row_argmax, pos_ids, n_pos, max_targets, sd, out);
    return launch_status();
}

}  // extern "C"
