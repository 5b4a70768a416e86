// Greedy NMS (2D / 3D) for sm_100a: upper-triangle bitmask kernel with shared-memory box tiles + on-device greedy reduction.
//
// Reference semantics followed (paths relative to the reference root):
//   IoU arithmetic      cuda_functions/nms_3D/src/cuda/nms_kernel.cu:16-28 (2D: nms_2D/src/cuda/nms_kernel.cu:16-24): +1 extents, fp32, IEEE division
//   suppression test    nms_kernel.cu:61-77 : strict `>`; box i suppresses only boxes j > i (diagonal block starts at threadIdx+1)
//   greedy reduction    cuda_functions/nms_3D/src/nms_cuda.c:33-61 (host loop over a D2H copy of the mask) -> here a device kernel
//
// Arithmetic pinning: nvcc (default -fmad=true) compiles the reference's `Sa + Sb - interS` as fma(sb_hw, sb_d, Sa) - interS
// (checked in the sm_100a SASS of the unmodified reference file). We spell that sequence with explicit intrinsics so that neither
// our compiler flags nor the CPU oracle (oracle/nms_oracle.c uses fmaf) can drift from it.
#include "mdt_common.cuh"

namespace mdt {

constexpr int kTile = 64;  // boxes per mask word, as in the reference (threadsPerBlock = 64)

template <int DIM>
struct BoxF;  // floats per row
template <> struct BoxF<2> { static constexpr int n = 5; };
template <> struct BoxF<3> { static constexpr int n = 7; };

struct Box3 { float y1, x1, y2, x2, z1, z2, vol; };
struct Box2 { float y1, x1, y2, x2, vol; };

__device__ __forceinline__ float box_volume_a(const float *a, int dim) {
    // Sa = (a2-a0+1)*(a3-a1+1)*(a5-a4+1): two FMULs in the reference SASS
    float e0 = __fadd_rn(__fsub_rn(a[2], a[0]), 1.f);
    float e1 = __fadd_rn(__fsub_rn(a[3], a[1]), 1.f);
    float v = __fmul_rn(e0, e1);
    if (dim == 3) v = __fmul_rn(v, __fadd_rn(__fsub_rn(a[5], a[4]), 1.f));
    return v;
}

// predicate IoU(a,b) > thresh with `a` the row box (its volume Sa precomputed) and `b` the column box
template <int DIM>
__device__ __forceinline__ bool suppresses(const float *a, float Sa, const float *b, float thresh, bool fast_reject) {
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float width = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
    float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
    float inter = __fmul_rn(width, height);
    if (DIM == 3) {
        float front = fmaxf(a[4], b[4]), back = fminf(a[5], b[5]);
        float depth = fmaxf(__fadd_rn(__fsub_rn(back, front), 1.f), 0.f);
        inter = __fmul_rn(inter, depth);
    }
    // inter == 0  =>  IoU is +-0 or NaN, and `> thresh` is false for every thresh >= 0: skip the division (most pairs are disjoint)
    if (fast_reject && inter == 0.f) return false;
    const float bh = __fadd_rn(__fsub_rn(b[2], b[0]), 1.f), bw = __fadd_rn(__fsub_rn(b[3], b[1]), 1.f);
    float sum;  // Sa + Sb, contracted exactly like the reference build: 3D fma(bh*bw, bd, Sa); 2D fma(bh, bw, Sa)
    if (DIM == 3) sum = __fmaf_rn(__fmul_rn(bh, bw), __fadd_rn(__fsub_rn(b[5], b[4]), 1.f), Sa);
    else          sum = __fmaf_rn(bh, bw, Sa);
    float uni = __fsub_rn(sum, inter);
    return __fdiv_rn(inter, uni) > thresh;
}

// One CTA per row block (64 boxes held in registers, one per lane of each of the 4 column groups); the 4 groups of 64 threads walk the
// column tiles rb, rb+1, ... four at a time (upper triangle only unless `full`, which reproduces the reference `_nms` contract where the
// lower triangle is computed too, nms_kernel.cu:35 being commented out).  Column tiles are staged in shared memory with coalesced loads;
// the four groups write four ADJACENT mask words of each row in the same iteration, so every 32-byte sector of the mask is completed
// within one iteration.  Heaviest row blocks (most column tiles) are scheduled first.
constexpr int kMaskGroups = 4;

template <int DIM>
__global__ void __launch_bounds__(kTile *kMaskGroups) nms_mask_kernel(int n, float thresh, const float *__restrict__ boxes,
                                                                    unsigned long long *__restrict__ mask, int col_blocks, int full) {
    constexpr int F = BoxF<DIM>::n;
    const int row_blk = blockIdx.x;
    const int group = threadIdx.x / kTile, t = threadIdx.x % kTile;
    const int row_size = min(n - row_blk * kTile, kTile);
    __shared__ float tile[kMaskGroups][kTile * F];
    const int cur = row_blk * kTile + t;
    const bool row_live = t < row_size;
    float a[F];
#pragma unroll
    for (int k = 0; k < F - 1; ++k) a[k] = row_live ? boxes[(size_t)cur * F + k] : 0.f;
    const float Sa = box_volume_a(a, DIM);
    const bool fast = thresh >= 0.f;
    const int c_begin = full ? 0 : row_blk;
    for (int c0 = c_begin; c0 < col_blocks; c0 += kMaskGroups) {
        const int col_blk = c0 + group;
        const bool live = col_blk < col_blocks;
        const int col_size = live ? min(n - col_blk * kTile, kTile) : 0;
        for (int i = t; i < col_size * F; i += kTile) tile[group][i] = boxes[(size_t)col_blk * kTile * F + i];
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kTile) : "memory");   // group-local barrier
        if (live && row_live) {
            unsigned long long word = 0;
            const int start = (row_blk == col_blk) ? t + 1 : 0;   // a box may only suppress LOWER-scored boxes of its own block
            for (int i = start; i < col_size; ++i)
                if (suppresses<DIM>(a, Sa, tile[group] + i * F, thresh, fast)) word |= 1ULL << i;
            mask[(size_t)cur * col_blocks + col_blk] = word;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kTile) : "memory");
    }
}

// Greedy reduction on one CTA — the exact recurrence of nms_cuda.c:47-58 with remv (the suppression bitmap) in shared memory.
// Boxes are consumed 64 at a time:
//   (1) one thread resolves the 64 keep decisions of the block from the 64 diagonal mask words (registers / shared memory only);
//   (2) the 1024 threads, arranged as 64 rows x 16 word-lanes, OR the mask rows of the boxes just kept into remv: every thread streams
//       its strided share of a row with independent 8-byte loads (several in flight) and touches remv only for non-zero words.
// The diagonal words of the NEXT block do not depend on remv and are prefetched during (2).
constexpr int kScanThreads = 1024;
constexpr int kScanLanes = kScanThreads / kTile;   // 16 word-lanes per row

__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(int n, int col_blocks, const unsigned long long *__restrict__ mask,
                                                               int64_t *__restrict__ keep, int *__restrict__ num_out) {
    extern __shared__ unsigned long long remv[];  // [col_blocks]
    __shared__ unsigned long long diag[2][kTile];
    __shared__ unsigned long long s_kept;
    __shared__ int s_count;
    for (int j = threadIdx.x; j < col_blocks; j += kScanThreads) remv[j] = 0ULL;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const int r = threadIdx.x / kScanLanes, l = threadIdx.x % kScanLanes;
    __shared__ int s_next;
    int b = 0;
    while (b < col_blocks) {
        // skip ahead over blocks whose boxes are all suppressed already (remv[b..] is final for them: only kept boxes ever add bits):
        // one thread walks the shared-memory bitmap, no global access, one barrier per RUN of skipped blocks
        if (threadIdx.x == 0) {
            int nb = b;
            while (nb < col_blocks) {
                const int sz = min(n - nb * kTile, kTile);
                const unsigned long long all = sz == kTile ? ~0ULL : ((1ULL << sz) - 1ULL);
                if ((remv[nb] & all) != all) break;
                ++nb;
            }
            s_next = nb;
        }
        __syncthreads();
        b = s_next;
        if (b >= col_blocks) break;
        const int base = b * kTile;
        const int size = min(n - base, kTile);
        if ((int)threadIdx.x < size) diag[0][threadIdx.x] = mask[(size_t)(base + threadIdx.x) * col_blocks + b];
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long rm = remv[b], kept = 0ULL;
            int cnt = s_count;
            for (int i = 0; i < size; ++i) {
                if (!((rm >> i) & 1ULL)) {
                    kept |= 1ULL << i;
                    keep[cnt++] = base + i;
                    rm |= diag[0][i];
                }
            }
            s_kept = kept;
            s_count = cnt;
        }
        __syncthreads();
        const unsigned long long kept_now = s_kept;
        if (__popcll(kept_now) <= 8) {
            // few survivors in this block (the usual case once most boxes are suppressed): all 1024 threads stream each kept row, one word each
            unsigned long long k = kept_now;
            while (k) {
                const int i = __ffsll((long long)k) - 1;
                k &= k - 1;
                const unsigned long long *row = mask + (size_t)(base + i) * col_blocks;
                for (int j = b + 1 + threadIdx.x; j < col_blocks; j += kScanThreads) {
                    const unsigned long long v = row[j];
                    if (v) atomicOr(&remv[j], v);
                }
            }
        } else if ((kept_now >> r) & 1ULL) {
            // many survivors: 64 rows in parallel, 16 word-lanes each
            const unsigned long long *row = mask + (size_t)(base + r) * col_blocks;
#pragma unroll 4
            for (int j = b + 1 + l; j < col_blocks; j += kScanLanes) {
                const unsigned long long v = row[j];
                if (v) atomicOr(&remv[j], v);
            }
        }
        __syncthreads();
        ++b;
    }
    if (threadIdx.x == 0) *num_out = s_count;
}

template <int DIM>
static int launch_mask(int n, const float *boxes, unsigned long long *mask, float thresh, int full, cudaStream_t st) {
    if (n < 0 || (n > 0 && (!boxes || !mask))) return MDT_EINVAL;
    if (n == 0) return MDT_OK;
    const int cb = ceil_div(n, kTile);
    nms_mask_kernel<DIM><<<cb, kTile * kMaskGroups, 0, st>>>(n, thresh, boxes, mask, cb, full);
    return launch_status();
}

template <int DIM>
static int nms_fused(const float *boxes, int n, float thresh, void *ws, size_t ws_bytes, int64_t *keep, int *num_out, cudaStream_t st) {
    if (n < 0 || !num_out || (n > 0 && (!boxes || !keep || !ws))) return MDT_EINVAL;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(num_out, 0, sizeof(int), st);
        return e == cudaSuccess ? MDT_OK : (int)e;
    }
    if (ws_bytes < mdt_nms_workspace_bytes(n)) return MDT_EWORKSPACE;
    const int cb = ceil_div(n, kTile);
    const size_t smem = (size_t)cb * sizeof(unsigned long long);
    if (smem > 200 * 1024) return MDT_EUNSUPPORTED;  // N <= 1.6 M boxes
    auto *mask = reinterpret_cast<unsigned long long *>(ws);
    int rc = launch_mask<DIM>(n, boxes, mask, thresh, /*full=*/0, st);
    if (rc != MDT_OK) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    nms_scan_kernel<<<1, kScanThreads, smem, st>>>(n, cb, mask, keep, num_out);
    return launch_status();
}

}  // namespace mdt

extern "C" {

size_t mdt_nms_workspace_bytes(int boxes_num) {
    if (boxes_num <= 0) return 0;
    size_t cb = mdt::ceil_div(boxes_num, mdt::kTile);
    return (size_t)boxes_num * cb * sizeof(unsigned long long);
}

int mdt_nms_mask_3d(int n, const float *boxes, unsigned long long *mask, float thresh, void *stream) {
    return mdt::launch_mask<3>(n, boxes, mask, thresh, 1, mdt::as_stream(stream));
}
int mdt_nms_mask_2d(int n, const float *boxes, unsigned long long *mask, float thresh, void *stream) {
    return mdt::launch_mask<2>(n, boxes, mask, thresh, 1, mdt::as_stream(stream));
}
int mdt_nms_3d(const float *boxes, int n, float thresh, void *ws, size_t ws_bytes, int64_t *keep, int *num_out, void *stream) {
    return mdt::nms_fused<3>(boxes, n, thresh, ws, ws_bytes, keep, num_out, mdt::as_stream(stream));
}
int mdt_nms_2d(const float *boxes, int n, float thresh, void *ws, size_t ws_bytes, int64_t *keep, int *num_out, void *stream) {
    return mdt::nms_fused<2>(boxes, n, thresh, ws, ws_bytes, keep, num_out, mdt::as_stream(stream));
}

}  // extern "C"
