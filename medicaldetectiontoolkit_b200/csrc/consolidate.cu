// Inference-side consolidation on the device: weighted box clustering and the 2D -> 3D slice merge of the reference's
// predictor.py:597-706 / :710-773 (numpy loops over a shrinking `order` array, one patient and class at a time, in a process pool).
//
// Both are greedy cluster loops like NMS with a data-dependent number of iterations and tiny inputs (hundreds to a few thousand boxes), so
// the design is ONE resident CTA per problem that runs the whole loop: all per-box state stays in registers/shared memory/L2, every iteration
// is a strided pass over the still-alive boxes + a block reduction, nothing returns to the host between clusters.  All arithmetic is fp64 in
// numpy's expression order (`ovr = inter / (area_i + area_j - inter)`, `+ 1` pixel convention), so the cluster membership decisions
// (`ovr > thresh`) are the reference's; only the order of the SUMS differs from np.sum's pairwise order (results agree to ~1e-15 relative).
#include "mdt_common.cuh"

namespace mdt {

constexpr int kConsThreads = 1024;
constexpr int kMaxRed = 12;   // doubles reduced per iteration: sw, ss, sn, 6 coords (+ spare)

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// IoU of boxes i and j in the reference's arithmetic (predictor.py:640-657): [y1, x1, y2, x2, (z1, z2)] with inclusive pixel extents
template <int DIM>
__device__ __forceinline__ double wbc_iou(const double *__restrict__ bi, double area_i, const double *__restrict__ bj, double area_j) {
    const double xx1 = fmax(bi[1], bj[1]), yy1 = fmax(bi[0], bj[0]);
    const double xx2 = fmin(bi[3], bj[3]), yy2 = fmin(bi[2], bj[2]);
    const double w = fmax(0.0, __dadd_rn(__dsub_rn(xx2, xx1), 1.0)), h = fmax(0.0, __dadd_rn(__dsub_rn(yy2, yy1), 1.0));
    double inter = __dmul_rn(w, h);
    if (DIM == 3) {
        const double zz1 = fmax(bi[4], bj[4]), zz2 = fmin(bi[5], bj[5]);
        const double d = fmax(0.0, __dadd_rn(__dsub_rn(zz2, zz1), 1.0));
        inter = __dmul_rn(inter, d);
    }
    return __ddiv_rn(inter, __dsub_rn(__dadd_rn(area_i, area_j), inter));
}

template <int DIM>
__device__ __forceinline__ double box_area(const double *__restrict__ b) {
    double a = __dmul_rn(__dadd_rn(__dsub_rn(b[2], b[0]), 1.0), __dadd_rn(__dsub_rn(b[3], b[1]), 1.0));
    if (DIM == 3) a = __dmul_rn(a, __dadd_rn(__dsub_rn(b[5], b[4]), 1.0));
    return a;
}

// dets [n, 2*DIM + 3] = coords, score, patch-centre factor, number of overlapping patches; order = indices by descending score.
// ws: alive [n] bytes | stamp [n_patches] ints (last iteration that saw a patch, for np.unique(match_patch_id).shape[0])
template <int DIM>
__global__ void __launch_bounds__(kConsThreads) wbc_kernel(const double *__restrict__ dets, const int *__restrict__ patch_id, const int *__restrict__ order, int n,
                                                           int n_patches, double thresh, double n_ens, double *__restrict__ keep_scores,
                                                           double *__restrict__ keep_coords, int *__restrict__ n_keep, unsigned char *alive, int *stamp) {
    constexpr int W = 2 * DIM + 3, NC = 2 * DIM;
    __shared__ double s_red[kConsThreads / 32][kMaxRed];
    __shared__ int s_cnt[kConsThreads / 32][2];
    __shared__ double s_box[6];
    __shared__ double s_area;
    __shared__ int s_head, s_i, s_out;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int j = tid; j < n; j += kConsThreads) alive[j] = 1;
    for (int j = tid; j < n_patches; j += kConsThreads) stamp[j] = -1;
    if (tid == 0) { s_head = 0; s_out = 0; }
    __syncthreads();
    for (int it = 0;; ++it) {
        if (tid == 0) {
            int h = s_head;
            while (h < n && !alive[order[h]]) ++h;
            s_head = h;
            s_i = h < n ? order[h] : -1;
            if (h < n) {
                const double *b = dets + (size_t)order[h] * W;
                for (int k = 0; k < NC; ++k) s_box[k] = b[k];
                s_area = box_area<DIM>(b);
            }
        }
        __syncthreads();
        const int i = s_i, head = s_head;
        if (i < 0) break;
        double acc[3 + NC];
#pragma unroll
        for (int k = 0; k < 3 + NC; ++k) acc[k] = 0.0;
        int cnt = 0, uniq = 0;
        for (int p = head + tid; p < n; p += kConsThreads) {
            const int j = order[p];
            if (!alive[j]) continue;
            const double *b = dets + (size_t)j * W;
            const double area_j = box_area<DIM>(b);
            const double ovr = wbc_iou<DIM>(s_box, s_area, b, area_j);
            if (!(ovr > thresh)) continue;
            const double wgt = __dmul_rn(__dmul_rn(ovr, area_j), b[NC + 1]);     // match_ov_facts * match_areas * match_pc_facts
            const double sc = __dmul_rn(b[NC], wgt);                              // match_scores *= match_score_weights
            acc[0] += wgt; acc[1] += sc; acc[2] += b[NC + 2];
#pragma unroll
            for (int k = 0; k < NC; ++k) acc[3 + k] += __dmul_rn(b[k], sc);
            ++cnt;
            const int pid = patch_id[j];
            if (pid >= 0 && pid < n_patches && atomicExch(&stamp[pid], it) != it) ++uniq;
            alive[j] = 0;
        }
#pragma unroll
        for (int k = 0; k < 3 + NC; ++k) {
            const double v = warp_sum_f64(acc[k]);
            if (lane == 0) s_red[warp][k] = v;
        }
        cnt = warp_sum_i32(cnt); uniq = warp_sum_i32(uniq);
        if (lane == 0) { s_cnt[warp][0] = cnt; s_cnt[warp][1] = uniq; }
        __syncthreads();
        if (tid == 0) {
            double tot[3 + NC];
            for (int k = 0; k < 3 + NC; ++k) tot[k] = 0.0;
            int c = 0, u = 0;
            for (int w = 0; w < kConsThreads / 32; ++w) {
                for (int k = 0; k < 3 + NC; ++k) tot[k] += s_red[w][k];
                c += s_cnt[w][0]; u += s_cnt[w][1];
            }
            const double n_expected = n_ens * (tot[2] / (double)c);              // n_ens * np.mean(match_n_ovs)
            const double n_missing = fmax(0.0, n_expected - (double)u);         // np.max((0, n_expected - #unique patches))
            const double denom = tot[0] + n_missing * (tot[0] / (double)c);      // sum(w) + n_missing * mean(w)
            const double avg = tot[1] / denom;
            if (avg > 0.01) {
                const int o = s_out++;
                keep_scores[o] = avg;
                for (int k = 0; k < NC; ++k) keep_coords[(size_t)o * NC + k] = tot[3 + k] / tot[1];
            }
            s_head = head + 1;                                                    // box i matched itself (ovr = 1) and is gone
        }
        __syncthreads();
    }
    if (tid == 0) *n_keep = s_out;
}

// dets [n, 6] = y1, x1, y2, x2, score, slice id.  keep[k] = index of the cluster's core box, keep_z[k] = {z1, z2} (predictor.py:710-773)
__global__ void __launch_bounds__(kConsThreads) nms_2to3d_kernel(const double *__restrict__ dets, const int *__restrict__ order, int n, double thresh,
                                                                 int n_slices, long long *__restrict__ keep, double *__restrict__ keep_z,
                                                                 int *__restrict__ n_keep, unsigned char *alive, unsigned char *match, int *present) {
    __shared__ int s_head, s_i, s_out, s_lo, s_hi, s_zmin, s_zmax, s_smin, s_smax;
    __shared__ double s_box[4];
    __shared__ double s_area;
    const int tid = threadIdx.x;
    for (int j = tid; j < n; j += kConsThreads) alive[j] = 1;
    if (tid == 0) { s_head = 0; s_out = 0; }
    __syncthreads();
    for (;;) {
        for (int s = tid; s < n_slices; s += kConsThreads) present[s] = 0;
        if (tid == 0) {
            int h = s_head;
            while (h < n && !alive[order[h]]) ++h;
            s_head = h;
            s_i = h < n ? order[h] : -1;
            if (h < n) {
                const double *b = dets + (size_t)order[h] * 6;
                for (int k = 0; k < 4; ++k) s_box[k] = b[k];
                s_area = __dmul_rn(__dadd_rn(__dsub_rn(b[3], b[1]), 1.0), __dadd_rn(__dsub_rn(b[2], b[0]), 1.0));
            }
            s_smin = 0x7fffffff; s_smax = -0x7fffffff; s_zmin = 0x7fffffff; s_zmax = -0x7fffffff;
        }
        __syncthreads();
        const int i = s_i, head = s_head;
        if (i < 0) break;
        // pass 1: the xy matches among the alive boxes and the slices they occupy
        for (int p = head + tid; p < n; p += kConsThreads) {
            const int j = order[p];
            match[j] = 0;
            if (!alive[j]) continue;
            const double *b = dets + (size_t)j * 6;
            const double area_j = __dmul_rn(__dadd_rn(__dsub_rn(b[3], b[1]), 1.0), __dadd_rn(__dsub_rn(b[2], b[0]), 1.0));
            const double ovr = wbc_iou<2>(s_box, s_area, b, area_j);
            if (!(ovr > thresh)) continue;
            match[j] = 1;
            const int sl = (int)b[5];
            if (sl >= 0 && sl < n_slices) present[sl] = 1;
            atomicMin(&s_smin, sl);
            atomicMax(&s_smax, sl);
        }
        __syncthreads();
        if (tid == 0) {
            // "connected" slices around the core slice: up to the first slice without a prediction on either side (:746-751)
            const int core = (int)dets[(size_t)i * 6 + 5];
            int hi = s_smax, lo = s_smin;
            for (int s = core; s < s_smax; ++s)
                if (s < 0 || s >= n_slices || !present[s]) { hi = s; break; }
            for (int s = core - 1; s >= s_smin; --s)
                if (s < 0 || s >= n_slices || !present[s]) { lo = s; break; }
            s_hi = hi; s_lo = lo;
        }
        __syncthreads();
        // pass 2: the z-connected matches form the cube and leave; the other matches stay for later clusters
        for (int p = head + tid; p < n; p += kConsThreads) {
            const int j = order[p];
            if (!match[j]) continue;
            const int sl = (int)dets[(size_t)j * 6 + 5];
            if (sl > s_hi || sl < s_lo) continue;
            alive[j] = 0;
            atomicMin(&s_zmin, sl);
            atomicMax(&s_zmax, sl);
        }
        __syncthreads();
        if (tid == 0) {
            const int o = s_out++;
            keep[o] = i;
            keep_z[2 * o] = (double)(s_zmin - 1);
            keep_z[2 * o + 1] = (double)(s_zmax + 1);
            s_head = head + 1;
        }
        __syncthreads();
    }
    if (tid == 0) *n_keep = s_out;
}

static inline size_t al16(size_t v) { return (v + 15) / 16 * 16; }

}  // namespace mdt

extern "C" {

size_t mdt_wbc_workspace_bytes(int n, int n_patches) {
    if (n <= 0 || n_patches < 0) return 0;
    return mdt::al16((size_t)n) + mdt::al16((size_t)(n_patches > 0 ? n_patches : 1) * sizeof(int));
}

int mdt_wbc(const double *dets, const int *patch_id, const int *order, int n, int dim, int n_patches, double thresh, double n_ens, double *keep_scores,
            double *keep_coords, int *n_keep, void *ws, size_t ws_bytes, void *stream) {
    if (!n_keep || n < 0 || (dim != 2 && dim != 3) || n_patches < 0) return MDT_EINVAL;
    cudaStream_t st = mdt::as_stream(stream);
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(n_keep, 0, sizeof(int), st);
        return e == cudaSuccess ? MDT_OK : (int)e;
    }
    if (!dets || !patch_id || !order || !keep_scores || !keep_coords || !ws) return MDT_EINVAL;
    if (ws_bytes < mdt_wbc_workspace_bytes(n, n_patches)) return MDT_EWORKSPACE;
    unsigned char *alive = reinterpret_cast<unsigned char *>(ws);
    int *stamp = reinterpret_cast<int *>(alive + mdt::al16((size_t)n));
    if (dim == 3)
        mdt::wbc_kernel<3><<<1, mdt::kConsThreads, 0, st>>>(dets, patch_id, order, n, n_patches, thresh, n_ens, keep_scores, keep_coords, n_keep, alive, stamp);
    else
        mdt::wbc_kernel<2><<<1, mdt::kConsThreads, 0, st>>>(dets, patch_id, order, n, n_patches, thresh, n_ens, keep_scores, keep_coords, n_keep, alive, stamp);
    return mdt::launch_status();
}

size_t mdt_nms_2to3d_workspace_bytes(int n, int n_slices) {
    if (n <= 0 || n_slices < 0) return 0;
    return 2 * mdt::al16((size_t)n) + mdt::al16((size_t)(n_slices > 0 ? n_slices : 1) * sizeof(int));
}

int mdt_nms_2to3d(const double *dets, const int *order, int n, double thresh, int n_slices, long long *keep, double *keep_z, int *n_keep, void *ws,
                  size_t ws_bytes, void *stream) {
    if (!n_keep || n < 0 || n_slices < 0) return MDT_EINVAL;
    cudaStream_t st = mdt::as_stream(stream);
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(n_keep, 0, sizeof(int), st);
        return e == cudaSuccess ? MDT_OK : (int)e;
    }
    if (!dets || !order || !keep || !keep_z || !ws) return MDT_EINVAL;
    if (ws_bytes < mdt_nms_2to3d_workspace_bytes(n, n_slices)) return MDT_EWORKSPACE;
    unsigned char *alive = reinterpret_cast<unsigned char *>(ws);
    unsigned char *match = alive + mdt::al16((size_t)n);
    int *present = reinterpret_cast<int *>(match + mdt::al16((size_t)n));
    mdt::nms_2to3d_kernel<<<1, mdt::kConsThreads, 0, st>>>(dets, order, n, thresh, n_slices, keep, keep_z, n_keep, alive, match, present);
    return mdt::launch_status();
}

}  // extern "C"
