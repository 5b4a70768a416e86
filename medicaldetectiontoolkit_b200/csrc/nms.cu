// Greedy NMS (2D / 3D) for sm_100a: upper-triangle bitmask kernel with shared-memory box tiles + on-device greedy reduction.
//
// Reference semantics followed (paths relative to the reference root):
//   IoU arithmetic      cuda_functions/nms_3D/src/cuda/nms_kernel.cu:16-28 (2D: nms_2D/src/cuda/nms_kernel.cu:16-24): +1 extents, fp32, IEEE division
//   suppression test    nms_kernel.cu:61-77 : strict `>`; box i suppresses only boxes j > i (diagonal block starts at threadIdx+1)
//   greedy reduction    cuda_functions/nms_3D/src/nms_cuda.c:33-61 (host loop over a D2H copy of the mask) -> here a device kernel
//
// Arithmetic pinning: nvcc (default -fmad=true) compiles the reference's `Sa + Sb - interS` as fma(sb_hw, sb_d, Sa) - interS
// (checked in the sm_100a SASS of the unmodified reference file). We spell that sequence with explicit intrinsics so that neither
// our compiler flags nor the CPU oracle (oracle/mdt_oracle.c uses fmaf) can drift from it.
#include <stdlib.h>

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
// (Round 2 measured a two-sided multiply filter in front of the IEEE division — bit-identical keep lists, but 11.09 vs 10.57 ms at 100 k boxes:
// the extra compares cost more than the divisions they avoid; removed.  profiles/r02_nms_filter_ab.txt)
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
__device__ __forceinline__ bool suppresses_tile(const float *a, float Sa, const float4 *tile, int i, float thresh, bool fast) {
    float b[6];
    const float4 b0 = tile[(DIM == 3 ? 2 : 1) * i];   // one (3D: two) 128-bit broadcast loads per column box
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
    if (DIM == 3) { const float4 b1 = tile[2 * i + 1]; b[4] = b1.x; b[5] = b1.y; } else { b[4] = b[5] = 0.f; }
    return suppresses<DIM>(a, Sa, b, thresh, fast);
}

// y extent [min y1, max y2] of every 64-box tile (one warp per tile).  The mask kernel skips a (row block, column tile) pair whose extents are
// separated: then rn(min(a.y2, b.y2) - max(a.y1, b.y1)) <= rn(row_hi - col_lo) <= -1 for every pair (rounding is monotone, -1 is exact), so
// the kernel's own `width = max(rn(rn(right - left) + 1), 0)` is 0, inter is 0 and `inter / uni > thresh` is false for thresh >= 0: the word
// is written as 0 without touching the boxes.  Callers that translate independent NMS groups apart along y (one launch for all (batch
// element, class) groups, retina_unet.refine_detections) and order the boxes group by group skip every cross-group tile this way.
template <int DIM>
__global__ void __launch_bounds__(256) nms_tile_bounds_kernel(int n, const float *__restrict__ boxes, float2 *__restrict__ bounds, int col_blocks) {
    constexpr int F = BoxF<DIM>::n;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tile >= col_blocks) return;
    float lo = INFINITY, hi = -INFINITY;
    for (int i = lane; i < kTile; i += 32) {
        const int b = tile * kTile + i;
        if (b < n) { lo = fminf(lo, boxes[(size_t)b * F]); hi = fmaxf(hi, boxes[(size_t)b * F + 2]); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if (lane == 0) bounds[tile] = make_float2(lo, hi);
}

template <int DIM>
__global__ void __launch_bounds__(kTile *kMaskGroups) nms_mask_kernel(int n, float thresh, const float *__restrict__ boxes,
                                                                    unsigned long long *__restrict__ mask, int col_blocks, int full,
                                                                    const float2 *__restrict__ bounds) {
    constexpr int F = BoxF<DIM>::n;
    constexpr int V = DIM == 3 ? 2 : 1;   // float4 slots per staged box
    const int row_blk = blockIdx.x;
    const int group = threadIdx.x / kTile, t = threadIdx.x % kTile;
    const int row_size = min(n - row_blk * kTile, kTile);
    __shared__ float4 tile[kMaskGroups][kTile * V];
    const int cur = row_blk * kTile + t;
    const bool row_live = t < row_size;
    float a[F];
#pragma unroll
    for (int k = 0; k < F - 1; ++k) a[k] = row_live ? boxes[(size_t)cur * F + k] : 0.f;
    const float Sa = box_volume_a(a, DIM);
    const bool fast = thresh >= 0.f;
    const int c_begin = full ? 0 : row_blk;
    const bool can_skip = bounds != nullptr && fast;
    const float2 rbd = can_skip ? bounds[row_blk] : make_float2(0.f, 0.f);
    for (int c0 = c_begin; c0 < col_blocks; c0 += kMaskGroups) {
        const int col_blk = c0 + group;
        const bool live = col_blk < col_blocks;
        if (can_skip && live && col_blk != row_blk) {       // uniform over the 64 threads of the group: both group barriers are skipped together
            const float2 cbd = bounds[col_blk];
            if (__fsub_rn(rbd.y, cbd.x) <= -1.f || __fsub_rn(cbd.y, rbd.x) <= -1.f) {
                if (row_live) mask[(size_t)cur * col_blocks + col_blk] = 0ULL;
                continue;
            }
        }
        const int col_size = live ? min(n - col_blk * kTile, kTile) : 0;
        if (t < col_size) {   // coordinates only (the score column is not needed), as 16-byte vectors
            const float *src = boxes + (size_t)(col_blk * kTile + t) * F;
            tile[group][V * t] = make_float4(src[0], src[1], src[2], src[3]);
            if (DIM == 3) tile[group][V * t + 1] = make_float4(src[4], src[5], 0.f, 0.f);
        }
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kTile) : "memory");   // group-local barrier
        if (live && row_live) {
            unsigned long long word = 0;
            if (col_size == kTile && row_blk != col_blk) {
                // the bulk of the triangle, unrolled by 16 (ncu showed the rolled loop issue-bound at ~59 instructions per pair; ~25 here)
#pragma unroll 1
                for (int q = 0; q < kTile / 16; ++q) {   // 16 pairs per trip: constant bit positions and shared-memory offsets
                    const float4 *tq = tile[group] + q * 16 * V;
                    unsigned int bits = 0;
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        if (suppresses_tile<DIM>(a, Sa, tq, u, thresh, fast)) bits |= 1u << u;
                    word |= (unsigned long long)bits << (16 * q);
                }
            } else {
                const int start = (row_blk == col_blk) ? t + 1 : 0;   // a box may only suppress LOWER-scored boxes of its own block
                for (int i = start; i < col_size; ++i)
                    if (suppresses_tile<DIM>(a, Sa, tile[group], i, thresh, fast)) word |= 1ULL << i;
            }
            mask[(size_t)cur * col_blocks + col_blk] = word;
        }
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kTile) : "memory");
    }
}

// Greedy reduction on one CTA (first version; selected by MDT_NMS_SCAN=1 for A/B runs, the default is nms_scan_grid_kernel below) —
// the exact recurrence of nms_cuda.c:47-58 with remv (the suppression bitmap) in shared memory.  Boxes are consumed 64 at a time:
//   (1) one thread resolves the 64 keep decisions of the block from the 64 diagonal mask words (registers / shared memory only);
//   (2) the 1024 threads, arranged as 64 rows x 16 word-lanes, OR the mask rows of the boxes just kept into remv: every thread streams
//       its strided share of a row with independent 8-byte loads (several in flight) and touches remv only for non-zero words.
// Runs of blocks whose boxes are all suppressed already are skipped with one barrier per run.
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

// ---------------------------------------------------------------- grid-wide greedy reduction ----------------------------------------------------------------
// The single-CTA scan above is bound by the bytes ONE SM can keep in flight: at 100 k boxes / 51 k kept it streams 380 MB of kept
// rows at ~20 GB/s (ncu: profiles/r01_ncu_ops_summary.txt) and costs more than the mask kernel.  The grid version separates the two
// kinds of work of the recurrence:
//   serial phase (CTA 0)     a CHUNK of 16 blocks = 1024 boxes is decided from shared memory only: the 1024 x 16 words of the mask that
//                            couple the chunk's boxes with each other are staged once (rows that are still alive only), then every
//                            block that still has a live box costs one fully unrolled 64-step bit recurrence by one thread + one
//                            1024-thread OR of the kept rows into the chunk-local words; dead blocks cost nothing;
//   parallel phase (workers) the rows kept in the chunk are OR-ed into the global suppression bitmap for all words from the chunk
//                            AFTER the next one on; each worker CTA owns a contiguous word range, each lane one word (register
//                            accumulator, no atomics), warps split the kept rows: the whole GPU's load bandwidth, mask read once.
// CTA 0 ORs the kept rows into the NEXT chunk's 16 words itself, so it never waits for the ticket it has just published: the workers
// run one chunk behind (software pipeline).  Tickets go through a release/acquire sequence number; chunks in which nothing survives
// publish nothing.  Launched cooperatively (all CTAs co-resident: one per SM, 137 KB shared memory each); every spin is bounded and
// traps on time-out instead of hanging the GPU.
constexpr int kChunkBlocks = 16;
constexpr int kChunkRows = kChunkBlocks * kTile;   // 1024 = threads per CTA
constexpr int kChunkPitch = kChunkBlocks + 1;      // odd word pitch: conflict-free walks down a column
constexpr long long kSpinLimit = 6000000000LL;     // ~3 s of SM clocks

struct ScanCtl {        // in the workspace behind the bitmap; zeroed before every launch
    unsigned int seq;   // tickets published by CTA 0
    unsigned int done;  // worker CTAs that completed a ticket, summed over tickets
    int cnt0, cnt1;     // the ticket: keep[cnt0, cnt1) are the rows to OR in ...
    int word0;          // ... into words [word0, col_blocks); -1 = no more tickets
    int pad[3];
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spin until *p >= target (monotonic counters); traps instead of hanging the GPU if the protocol is ever broken
__device__ __forceinline__ void spin_until_ge(const unsigned int *p, unsigned int target) {
    const long long t0 = clock64();
    while (ld_acquire_u32(p) < target) {
        __nanosleep(20);
        if (clock64() - t0 > kSpinLimit) __trap();
    }
}

// worker `widx` of `workers`: OR keep[cnt0, cnt1)'s mask rows into its contiguous share of remv_g[word0, col_blocks)
__device__ __forceinline__ void scan_or_phase(int cnt0, int cnt1, int word0, int col_blocks, int widx, int workers,
                                              const unsigned long long *__restrict__ mask, const int64_t *keep, unsigned long long *remv_g,
                                              unsigned long long *s_acc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_cta = ceil_div(col_blocks - word0, workers);
    const int j0 = word0 + widx * per_cta, j1 = min(col_blocks, j0 + per_cta);
    for (int js = j0; js < j1; js += 32) {
        if (threadIdx.x < 32) s_acc[threadIdx.x] = 0ULL;
        __syncthreads();
        const int j = js + lane;
        if (j < j1) {
            unsigned long long acc = 0ULL;
            int k = cnt0 + warp;
            for (; k + 3 * 32 < cnt1; k += 4 * 32) {   // four independent rows in flight per lane
                const int64_t r0 = __ldcg(keep + k), r1 = __ldcg(keep + k + 32), r2 = __ldcg(keep + k + 64), r3 = __ldcg(keep + k + 96);
                const unsigned long long v0 = mask[(size_t)r0 * col_blocks + j], v1 = mask[(size_t)r1 * col_blocks + j];
                const unsigned long long v2 = mask[(size_t)r2 * col_blocks + j], v3 = mask[(size_t)r3 * col_blocks + j];
                acc |= (v0 | v1) | (v2 | v3);
            }
            for (; k < cnt1; k += 32) acc |= mask[(size_t)__ldcg(keep + k) * col_blocks + j];
            if (acc) atomicOr(&s_acc[lane], acc);
        }
        __syncthreads();
        if (threadIdx.x < 32 && js + (int)threadIdx.x < j1 && s_acc[threadIdx.x])   // this CTA is the only writer of these words for this ticket
            remv_g[js + threadIdx.x] = __ldcg(remv_g + js + threadIdx.x) | s_acc[threadIdx.x];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kChunkRows, 1) nms_scan_grid_kernel(int n, int col_blocks, const unsigned long long *__restrict__ mask,
                                                                     unsigned long long *remv_g, ScanCtl *ctl, int64_t *keep,
                                                                     int *__restrict__ num_out) {
    extern __shared__ unsigned long long s_chunk[];     // CTA 0: [kChunkRows][kChunkPitch] words of the current chunk
    __shared__ unsigned long long s_rm[kChunkBlocks];   // suppression words of the chunk's own blocks
    __shared__ unsigned long long s_next[kChunkBlocks]; // what this chunk's kept rows add to the NEXT chunk's words
    __shared__ unsigned long long s_acc[32];
    __shared__ unsigned long long s_kept;
    __shared__ int s_ticket[3];
    const int tid = threadIdx.x;
    const unsigned int W = gridDim.x - 1;   // worker CTAs

    if (blockIdx.x != 0) {   // ---------------- workers: wait for a ticket, OR, report
        for (unsigned int seen = 0;; ++seen) {
            if (tid == 0) {
                spin_until_ge(&ctl->seq, seen + 1);
                s_ticket[0] = *(volatile int *)&ctl->cnt0; s_ticket[1] = *(volatile int *)&ctl->cnt1; s_ticket[2] = *(volatile int *)&ctl->word0;
            }
            __syncthreads();
            const int cnt0 = s_ticket[0], cnt1 = s_ticket[1], word0 = s_ticket[2];
            if (word0 < 0) return;
            scan_or_phase(cnt0, cnt1, word0, col_blocks, (int)blockIdx.x - 1, (int)W, mask, keep, remv_g, s_acc);
            if (tid == 0) { __threadfence(); atomicAdd(&ctl->done, 1u); }
            __syncthreads();   // s_ticket is rewritten next round
        }
    }

    // ---------------- CTA 0: serial phase per chunk; all counters below are uniform across the CTA
    int count = 0;
    unsigned int tickets = 0, tickets_prev = 0, tickets_prev2 = 0;   // published so far / before the previous chunk / before the one before
    if (tid < kChunkBlocks) s_next[tid] = 0ULL;
    for (int c0 = 0; c0 < col_blocks; c0 += kChunkBlocks) {
        const int nblk = min(kChunkBlocks, col_blocks - c0);
        tickets_prev2 = tickets_prev;
        tickets_prev = tickets;
        // this chunk's words are final once the tickets of all chunks up to the one before the previous are complete (a ticket starts two
        // chunks ahead; the previous chunk's contribution is in s_next)
        if (W > 0 && tickets_prev2 > 0 && tid == 0) spin_until_ge(&ctl->done, W * tickets_prev2);
        __syncthreads();
        bool open = false;   // does block c0+tid still hold a box that is not suppressed?
        if (tid < nblk) {
            const unsigned long long w = __ldcg(remv_g + c0 + tid) | s_next[tid];
            s_rm[tid] = w;
            const int sz = min(n - (c0 + tid) * kTile, kTile);
            const unsigned long long all = sz == kTile ? ~0ULL : ((1ULL << sz) - 1ULL);
            open = (w & all) != all;
        }
        const int any_open = __syncthreads_or(open);
        if (tid < kChunkBlocks) s_next[tid] = 0ULL;
        if (!any_open) continue;   // nothing can be kept in this chunk: no staging, no ticket
        {   // stage the words that couple the chunk's live boxes with each other (one row per thread; words left of a row's own block are
            // never written by the mask kernel and never read below; rows already suppressed are never read either)
            const long long row = (long long)c0 * kTile + tid;
            const int own = tid / kTile;
            if (row < n && !((s_rm[own] >> (tid % kTile)) & 1ULL)) {
                const unsigned long long *src = mask + (size_t)row * col_blocks + c0;
#pragma unroll
                for (int w = 0; w < kChunkBlocks; ++w)
                    if (w >= own && w < nblk) s_chunk[tid * kChunkPitch + w] = src[w];
            }
        }
        __syncthreads();
        const int cnt_begin = count;
        for (int bl = 0; bl < nblk; ++bl) {
            const int base = (c0 + bl) * kTile;
            const int size = min(n - base, kTile);
            const unsigned long long all = size == kTile ? ~0ULL : ((1ULL << size) - 1ULL);
            const unsigned long long rm0 = s_rm[bl];            // final: ordered by the barrier that ended the previous iteration
            if ((rm0 & all) == all) continue;                   // every box of the block is suppressed already (uniform branch)
            if (tid == 0) {   // the reference's recurrence (nms_cuda.c:47-58) restricted to one block, fully unrolled: constant bit positions,
                              // the shared-memory loads do not depend on the chain
                unsigned long long rm = rm0 | ~all, kept = 0ULL;
                const unsigned long long *d = s_chunk + (size_t)(bl * kTile) * kChunkPitch + bl;
#pragma unroll
                for (int i = 0; i < kTile; ++i) {
                    const unsigned long long di = d[i * kChunkPitch];
                    if (!((rm >> i) & 1ULL)) { kept |= 1ULL << i; rm |= di; }
                }
                s_kept = kept;
            }
            __syncthreads();
            const unsigned long long kept = s_kept;
            if (tid < kTile && ((kept >> tid) & 1ULL)) keep[count + __popcll(kept & ((1ULL << tid) - 1ULL))] = base + tid;
            {   // kept rows -> the later words of this chunk (64 rows x 16 words = one element per thread)
                const int r = tid / kChunkBlocks, w = tid % kChunkBlocks;
                if (((kept >> r) & 1ULL) && w > bl && w < nblk) {
                    const unsigned long long v = s_chunk[(bl * kTile + r) * kChunkPitch + w];
                    if (v) atomicOr(&s_rm[w], v);
                }
            }
            count += __popcll(kept);
            __syncthreads();
        }
        const int word0 = c0 + nblk;
        if (count > cnt_begin && word0 < col_blocks) {
            const int wword0 = word0 + kChunkBlocks;   // the workers start behind the next chunk
            if (W > 0 && wword0 < col_blocks) {
                if (tid == 0) {
                    if (tickets > 0) spin_until_ge(&ctl->done, W * tickets);   // the ticket slot is free again (normally long since)
                    ctl->cnt0 = cnt_begin; ctl->cnt1 = count; ctl->word0 = wword0;
                    __threadfence();   // keep[] entries (ordered by the barriers above) and the ticket before the sequence number
                    st_release_u32(&ctl->seq, tickets + 1);
                }
                ++tickets;
            } else if (W == 0 && wword0 < col_blocks) {   // no workers were launched (tiny grids): do their share here
                scan_or_phase(cnt_begin, count, wword0, col_blocks, 0, 1, mask, keep, remv_g, s_acc);
            }
            // own share: the next chunk's words, 64 half-warps over the kept rows, one word per lane of a half-warp
            const int nw = min(kChunkBlocks, col_blocks - word0);
            const int h = tid / kChunkBlocks, l = tid % kChunkBlocks;
            if (l < nw) {
                unsigned long long acc = 0ULL;
                for (int k = cnt_begin + h; k < count; k += kChunkRows / kChunkBlocks)
                    acc |= mask[(size_t)__ldcg(keep + k) * col_blocks + word0 + l];
                if (acc) atomicOr(&s_next[l], acc);
            }
        }
    }
    if (tid == 0) {
        *num_out = count;
        if (W > 0) {
            if (tickets > 0) spin_until_ge(&ctl->done, W * tickets);
            ctl->word0 = -1;
            __threadfence();
            st_release_u32(&ctl->seq, tickets + 1);
        }
    }
}

template <int DIM>
static int launch_mask(int n, const float *boxes, unsigned long long *mask, float thresh, int full, cudaStream_t st, float2 *bounds = nullptr) {
    if (n < 0 || (n > 0 && (!boxes || !mask))) return MDT_EINVAL;
    if (n == 0) return MDT_OK;
    const int cb = ceil_div(n, kTile);
    if (bounds) {
        nms_tile_bounds_kernel<DIM><<<ceil_div(cb, 8), 256, 0, st>>>(n, boxes, bounds, cb);
        if (int rc = launch_status()) return rc;
    }
    nms_mask_kernel<DIM><<<cb, kTile * kMaskGroups, 0, st>>>(n, thresh, boxes, mask, cb, full, bounds);
    return launch_status();
}

static size_t scan_ctl_offset(int n) {   // workspace = mask [n][cb] | remv_g [cb] | ScanCtl
    const size_t cb = ceil_div(n, kTile);
    return (size_t)n * cb * sizeof(unsigned long long);
}

static int scan_variant() {   // MDT_NMS_SCAN=1 selects the single-CTA scan (kept for A/B measurements); read per call, it is cheap
    const char *e = getenv("MDT_NMS_SCAN");
    return (e && e[0] == '1') ? 1 : 2;
}

template <int DIM>
static int nms_fused(const float *boxes, int n, float thresh, void *ws, size_t ws_bytes, int64_t *keep, int *num_out, cudaStream_t st) {
    if (n < 0 || !num_out || (n > 0 && (!boxes || !keep || !ws))) return MDT_EINVAL;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(num_out, 0, sizeof(int), st);
        return e == cudaSuccess ? MDT_OK : (int)e;
    }
    if (ws_bytes < mdt_nms_workspace_bytes(n)) return MDT_EWORKSPACE;
    int cb = ceil_div(n, kTile);
    auto *mask = reinterpret_cast<unsigned long long *>(ws);
    // tile extents live behind the scan's control block (see mdt_nms_workspace_bytes)
    float2 *bounds = reinterpret_cast<float2 *>(reinterpret_cast<char *>(ws) + scan_ctl_offset(n) + (size_t)cb * sizeof(unsigned long long) + sizeof(ScanCtl));
    int rc = launch_mask<DIM>(n, boxes, mask, thresh, /*full=*/0, st, bounds);
    if (rc != MDT_OK) return rc;
    auto single_cta_scan = [&]() -> int {
        const size_t smem = (size_t)cb * sizeof(unsigned long long);
        if (smem > 200 * 1024) return MDT_EUNSUPPORTED;  // N <= 1.6 M boxes
        static bool attr_set[kMaxDevices] = {};
        if (!ensure_smem_attr(nms_scan_kernel, 200 * 1024, attr_set)) return MDT_EUNSUPPORTED;
        nms_scan_kernel<<<1, kScanThreads, smem, st>>>(n, cb, mask, keep, num_out);
        return launch_status();
    };
    if (scan_variant() == 1) return single_cta_scan();
    auto *remv_g = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(ws) + scan_ctl_offset(n));
    auto *ctl = reinterpret_cast<ScanCtl *>(remv_g + cb);
    cudaError_t e = cudaMemsetAsync(remv_g, 0, (size_t)cb * sizeof(unsigned long long) + sizeof(ScanCtl), st);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)kChunkRows * kChunkPitch * sizeof(unsigned long long);
    static bool grid_attr_set[kMaxDevices] = {};
    if (!ensure_smem_attr(nms_scan_grid_kernel, (int)smem, grid_attr_set)) return MDT_EUNSUPPORTED;
    // CTA 0 + workers, one CTA per SM at most (cooperative launch: all co-resident); a ticket starts two chunks ahead of the chunk that
    // issues it, so up to 32 blocks need no worker at all; ~8 bitmap words per worker and ticket at least
    int grid = cb > 2 * kChunkBlocks ? 1 + ceil_div(cb - 2 * kChunkBlocks, 8) : 1;
    // co-residency limit of THIS device (MIG / MPS slices and other GPUs have fewer SMs than the first device this process saw)
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nms_scan_grid_kernel, kChunkRows, smem) != cudaSuccess || sms * per_sm < 1) {
        (void)cudaGetLastError();
        return single_cta_scan();
    }
    if (grid > sms * per_sm) grid = sms * per_sm;
    void *args[] = {&n, &cb, &mask, &remv_g, &ctl, &keep, &num_out};
    e = cudaLaunchCooperativeKernel((const void *)nms_scan_grid_kernel, dim3(grid), dim3(kChunkRows), args, smem, st);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();          // e.g. cudaErrorCooperativeLaunchTooLarge under a reduced SM set: the single-CTA reduction always runs
        return single_cta_scan();
    }
    return launch_status();
}

}  // namespace mdt

extern "C" {

size_t mdt_nms_workspace_bytes(int boxes_num) {
    if (boxes_num <= 0) return 0;
    size_t cb = mdt::ceil_div(boxes_num, mdt::kTile);   // mask [n][cb] + the grid scan's bitmap [cb] and control block + tile extents [cb]
    return (size_t)boxes_num * cb * sizeof(unsigned long long) + cb * sizeof(unsigned long long) + sizeof(mdt::ScanCtl) + cb * sizeof(float2);
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
