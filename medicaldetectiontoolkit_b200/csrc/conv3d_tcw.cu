// conv3d fprop / dgrad on tcgen05, "tap-stacked" formulation (algo 2, W-contiguous convs whose lines fit one 128-row MMA tile).
//
// Replaces the cuDNN conv3d behind nn.Conv3d in the reference (utils/model_utils.py:762; models/backbone.py:27-206, heads in
// models/retina_unet.py:40-119, models/mrcnn.py:40-169).  Same arithmetic as conv3d_tc.cu (split-bf16 x3, fp32 TMEM accumulation),
// different GEMM shape — chosen from two measurements of the first kernel (profiles/r01_mma_rate.txt, r01_ncu_ops_summary.txt):
//   * a tcgen05.mma with both operands in shared memory costs max(N/2, (M + N)/4) cycles for K = 16 bf16: the M = 128 A rows alone are 32
//     cycles of shared-memory reads, so the N = 48..96 MMAs of the 18/36-channel layers ran at 40-55 % of the tensor rate;
//   * every (kd, kh, kw) tap re-streamed its weight tile and every (kd, kh) its activation line through L2 (630 KB per 128 voxels).
// Here the kw taps are stacked along N instead of being separate MMAs on row-shifted windows:
//      D[r, (kw, c)] = sum_k A[r, k] * W[kd, kh, kw][k, c]           (A = one source line: 128 voxels x K channels, fetched ONCE per (kd, line))
//      out[w, c]     = sum_kw D[w + kw - pw, (kw, c)]                (the shift along w moves into the epilogue: warp shuffles + a few edge rows)
// so one MMA has N = kw * C (108 -> 112 for 36 channels), and with the hi/lo weight planes stacked as well N = 224: tensor-bound, one A read
// per 224 columns.  A CTA owns TL = 2 adjacent output lines with one TMEM accumulator each: a source line is used by both (different kh), so
// activation lines are fetched (TL + KH - 1) / TL times per output line instead of KH times, and the (kd, kh) weight tiles flow through a FIFO
// ring in first-use order and serve both lines.  K is padded to 16 (36 -> 48: three K steps, not four) by splitting it into swizzle-width
// chunks (64 / 32 / 16 channels = 128B / 64B / 32B swizzle), each with its own tensor map over the same bf16 planes.
// Persistent CTAs (static round-robin over tiles); warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.  The producer runs
// ahead into the next tile while the epilogue drains; accumulator hand-over is per line (acc_full / acc_empty mbarriers).
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <vector>

#include "conv3d_common.cuh"
#include "conv3d_tc_plan.cuh"
#include "tc_common.cuh"

namespace mdt {
using namespace tc;

constexpr int kTcwThreads = 192;
constexpr int kTcwMaxChunks = 4;
constexpr int kTcwMaxLines = 16;
constexpr int kTcwMaxTL = 2;
constexpr int kTcwMaxKH = 8;
constexpr int kTcwMaxSA = 6, kTcwMaxSB = 8;
constexpr int kFlagFirst = 1, kFlagLast = 2, kFlagAccLast = 4;

// Position-independent schedule of one kd slice of a tile: source lines in ascending order, the (output line t, tap kh) pairs each line
// feeds, and the FIFO order of the weight tiles (first use).  Built on the host, identical for producer and consumer.
struct TcwSched {
    int nlines, ntiles;
    signed char line_rel[kTcwMaxLines];
    unsigned char npairs[kTcwMaxLines];
    unsigned char pair_t[kTcwMaxLines][kTcwMaxTL], pair_kh[kTcwMaxLines][kTcwMaxTL], pair_flags[kTcwMaxLines][kTcwMaxTL];
    unsigned char tile_order[kTcwMaxKH];   // kh -> position in the load order
    unsigned char load_kh[kTcwMaxKH];      // position -> kh
    unsigned char load_line[kTcwMaxKH];    // position -> schedule line before which the tile is fetched
};

struct TcwParams {
    int NB, RD, RH, RW, SD, SH;
    int KD, KH, KW, sd, sh, pd, ph, pw, dgrad;
    int Cn, CT, Cs, NW, stacked, planes;
    int TL, ACC, tmem_cols;
    int nchunk, ck0[kTcwMaxChunks], cw[kTcwMaxChunks], a_off[kTcwMaxChunks], b_off[kTcwMaxChunks], tm[kTcwMaxChunks];
    int a_stage_bytes, b_tile_bytes, SA, SB, a_tx, b_tx;
    int tiles_h;
    long long total_tiles;
    int relu;
    const float *bias, *residual;
    float *out;
    __nv_bfloat16 *out_split;   // optional: the result also as (hi, lo) bf16 planes in the canonical split layout [line][plane][w][Kg_out]
    int out_split_kg;
    int prof;
    TcwSched sch;
};

struct TcwMaps {
    CUtensorMap a[3], b[3];   // index 0 / 1 / 2 = 64 / 32 / 16-channel chunks (128B / 64B / 32B swizzle)
};

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float *v) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}

// role-level cycle counters of CTA 0 (MDT_TCW_PROF=1): [0] producer wait a_empty, [1] producer wait b_empty, [2] producer total,
// [3] mma wait a_full, [4] mma wait b_full, [5] mma wait acc_empty, [6] mma total, [7] epilogue wait acc_full, [8] epilogue total, [9] tiles
__device__ unsigned long long g_tcw_prof[16];
#define TCW_T0(flag) const long long _t0 = (flag) ? clock64() : 0
#define TCW_ACC(flag, var) do { if (flag) var += clock64() - _t0; } while (0)

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// source depth index feeding row-space depth rd through tap kd; false = this tap contributes nothing
__device__ __forceinline__ bool tcw_depth(const TcwParams &p, int rd, int kd, int &d_src) {
    if (!p.dgrad) {
        d_src = rd * p.sd - p.pd + kd;
    } else {
        const int td = rd + p.pd - kd;
        if (td < 0 || td % p.sd != 0) return false;
        d_src = td / p.sd;
    }
    return d_src >= 0 && d_src < p.SD;
}

__global__ void __launch_bounds__(kTcwThreads, 2)
conv_tcw_kernel(const __grid_constant__ TcwMaps maps, const __grid_constant__ TcwParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t a_full[kTcwMaxSA], a_empty[kTcwMaxSA], b_full[kTcwMaxSB], b_empty[kTcwMaxSB], acc_full[kTcwMaxTL], acc_empty[kTcwMaxTL];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_bias[256];
    __shared__ float s_edge[2][4][7][3][4];

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + (size_t)p.SA * p.a_stage_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * p.CT;
    const int ct = min(p.CT, p.Cn - n0);   // channels of this N tile

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < p.SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < p.TL; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        fence_barrier_init();
        for (int c = 0; c < p.nchunk; ++c) { prefetch_tmap(&maps.a[p.tm[c]]); prefetch_tmap(&maps.b[p.tm[c]]); }
    }
    for (int c = threadIdx.x; c < 256; c += kTcwThreads) s_bias[c] = (p.bias && c < ct) ? __ldg(p.bias + n0 + c) : 0.f;
    if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // =============================================================== TMA producer
        if (lane == 0) {
            uint32_t a_seq = 0, b_seq = 0;
            const bool prof = p.prof && blockIdx.x == 0 && blockIdx.y == 0;
            long long pw_a = 0, pw_b = 0;
            const long long p_start = prof ? clock64() : 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                long long u = tile;
                const int th = (int)(u % p.tiles_h); u /= p.tiles_h;
                const int rd = (int)(u % p.RD);
                const int nb = (int)(u / p.RD);
                const int rh0 = th * p.TL;
                const int line_base = p.dgrad ? rh0 / p.sh : rh0 * p.sh;
                for (int kd = 0; kd < p.KD; ++kd) {
                    int d_src;
                    if (!tcw_depth(p, rd, kd, d_src)) continue;
                    const int tap_base = (blockIdx.y * p.KD + kd) * p.KH;
                    int next_load = 0;
                    for (int i = 0; i < p.sch.nlines; ++i) {
                        while (next_load < p.sch.ntiles && p.sch.load_line[next_load] == i) {
                            const int kh = p.sch.load_kh[next_load];
                            const uint32_t slot = b_seq % p.SB;
                            { TCW_T0(prof); mbar_wait(&b_empty[slot], ((b_seq / p.SB) & 1) ^ 1); TCW_ACC(prof, pw_b); }
                            mbar_arrive_expect_tx(&b_full[slot], (uint32_t)p.b_tx);
                            uint8_t *bt = smem_b + (size_t)slot * p.b_tile_bytes;
                            for (int c = 0; c < p.nchunk; ++c)
                                for (int pl = 0; pl < p.planes; ++pl)
                                    tma_load_4d(bt + p.b_off[c] + (size_t)pl * p.NW * 2 * p.cw[c], &maps.b[p.tm[c]], &b_full[slot], p.ck0[c], 0, pl,
                                                tap_base + kh);
                            ++b_seq;
                            ++next_load;
                        }
                        const uint32_t slot = a_seq % p.SA;
                        { TCW_T0(prof); mbar_wait(&a_empty[slot], ((a_seq / p.SA) & 1) ^ 1); TCW_ACC(prof, pw_a); }
                        mbar_arrive_expect_tx(&a_full[slot], (uint32_t)p.a_tx);
                        uint8_t *as = smem_a + (size_t)slot * p.a_stage_bytes;
                        // one box = both planes of the source line: {chunk, 128 voxels, planes, 1, 1}; rows past the line end and lines outside
                        // the image are TMA zero fill (= the conv's zero padding)
                        for (int c = 0; c < p.nchunk; ++c)
                            tma_load_5d(as + p.a_off[c], &maps.a[p.tm[c]], &a_full[slot], p.ck0[c], 0, 0, line_base + p.sch.line_rel[i], nb * p.SD + d_src);
                        ++a_seq;
                    }
                }
            }
            if (prof) { g_tcw_prof[0] += pw_a; g_tcw_prof[1] += pw_b; g_tcw_prof[2] += clock64() - p_start; }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idescN = make_idesc_bf16(128, p.NW, 0, 0);
            const uint32_t idesc2N = make_idesc_bf16(128, 2 * p.NW, 0, 0);
            uint64_t dt[kTcwMaxChunks];
            for (int c = 0; c < p.nchunk; ++c) dt[c] = make_smem_desc(0, 16, 16u * p.cw[c], layout_type_for_swizzle_bytes(2 * p.cw[c]));
            const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
            uint32_t a_seq = 0, b_seq = 0, it = 0;
            const bool prof = p.prof && blockIdx.x == 0 && blockIdx.y == 0;
            long long mw_a = 0, mw_b = 0, mw_c = 0;
            const long long m_start = prof ? clock64() : 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                long long u = tile / p.tiles_h;
                const int rd = (int)(u % p.RD);
                int kd_last = -1, d_src;
                for (int kd = 0; kd < p.KD; ++kd)
                    if (tcw_depth(p, rd, kd, d_src)) kd_last = kd;
                uint32_t acc_started = 0;
                for (int kd = 0; kd < p.KD; ++kd) {
                    if (!tcw_depth(p, rd, kd, d_src)) continue;
                    const bool last_kd = kd == kd_last;
                    for (int i = 0; i < p.sch.nlines; ++i) {
                        const uint32_t slot_a = a_seq % p.SA;
                        { TCW_T0(prof); mbar_wait(&a_full[slot_a], (a_seq / p.SA) & 1); TCW_ACC(prof, mw_a); }
                        tc_fence_after();
                        const uint32_t a_base = sa0 + slot_a * p.a_stage_bytes;
                        const int np = p.sch.npairs[i];
                        for (int j = 0; j < np; ++j) {
                            const int t = p.sch.pair_t[i][j], kh = p.sch.pair_kh[i][j], fl = p.sch.pair_flags[i][j];
                            const uint32_t bs = b_seq + p.sch.tile_order[kh];
                            const uint32_t slot_b = bs % p.SB;
                            if (fl & kFlagFirst) { TCW_T0(prof); mbar_wait(&b_full[slot_b], (bs / p.SB) & 1); TCW_ACC(prof, mw_b); tc_fence_after(); }
                            uint32_t acc = (acc_started >> t) & 1u;
                            if (!acc) { TCW_T0(prof); mbar_wait(&acc_empty[t], (it & 1) ^ 1); TCW_ACC(prof, mw_c); tc_fence_after(); }
                            const uint32_t b_base = sb0 + slot_b * p.b_tile_bytes;
                            const uint32_t d_tmem = tmem + (uint32_t)(t * p.ACC);
                            for (int c = 0; c < p.nchunk; ++c) {
                                const uint32_t swz = 2u * p.cw[c];
                                const uint32_t a_hi = a_base + p.a_off[c], b_hi = b_base + p.b_off[c];
                                uint64_t da = dt[c] | (uint64_t)((a_hi >> 4) & 0x3FFF);
                                uint64_t dal = dt[c] | (uint64_t)(((a_hi + 128u * swz) >> 4) & 0x3FFF);
                                uint64_t db = dt[c] | (uint64_t)((b_hi >> 4) & 0x3FFF);
                                uint64_t dbl = dt[c] | (uint64_t)(((b_hi + (uint32_t)p.NW * swz) >> 4) & 0x3FFF);
                                const int ksteps = p.cw[c] >> 4;
                                for (int k = 0; k < ksteps; ++k) {
                                    if (p.stacked) {
                                        umma_bf16(d_tmem, da, db, idesc2N, acc);       // [0,NW) += hi*hi, [NW,2NW) += hi*lo
                                        umma_bf16(d_tmem, dal, db, idescN, 1);         // [0,NW) += lo*hi
                                    } else {
                                        umma_bf16(d_tmem, da, db, idescN, acc);
                                        if (p.planes > 1) {
                                            umma_bf16(d_tmem, da, dbl, idescN, 1);
                                            umma_bf16(d_tmem, dal, db, idescN, 1);
                                        }
                                    }
                                    acc = 1;
                                    da += 2; dal += 2; db += 2; dbl += 2;              // +32 bytes along K
                                }
                            }
                            acc_started |= 1u << t;
                            if (fl & kFlagLast) umma_commit(&b_empty[slot_b]);
                            if (last_kd && (fl & kFlagAccLast)) umma_commit(&acc_full[t]);
                        }
                        umma_commit(&a_empty[slot_a]);
                        ++a_seq;
                    }
                    b_seq += p.sch.ntiles;
                }
            }
            if (prof) { g_tcw_prof[3] += mw_a; g_tcw_prof[4] += mw_b; g_tcw_prof[5] += mw_c; g_tcw_prof[6] += clock64() - m_start; g_tcw_prof[9] += it; }
        }
    } else {
        // =============================================================== epilogue (warps 2..5 = TMEM lane quarters 2, 3, 0, 1)
        const int q = warp & 3;
        const int r = q * 32 + lane;          // accumulator row = source voxel index along the line; this thread produces output voxel w = r
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const bool two = p.stacked != 0;
        const int vecw = (p.Cn % 4 == 0 && n0 % 4 == 0) ? 4 : ((p.Cn % 2 == 0 && n0 % 2 == 0) ? 2 : 1);
        uint32_t it = 0, grp = 0;
        const bool prof = p.prof && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 64;
        long long ew = 0;
        const long long e_start = prof ? clock64() : 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            long long u = tile;
            const int th = (int)(u % p.tiles_h); u /= p.tiles_h;
            const int rd = (int)(u % p.RD);
            const int nb = (int)(u / p.RD);
            for (int t = 0; t < p.TL; ++t) {
                { TCW_T0(prof); mbar_wait(&acc_full[t], it & 1); TCW_ACC(prof, ew); }
                tc_fence_after();
                const int rh = th * p.TL + t;
                if (rh < p.RH) {
                    const uint32_t acc0 = lane_base + (uint32_t)(t * p.ACC);
                    const size_t line_off = (((size_t)nb * p.RD + rd) * p.RH + rh) * (size_t)p.RW;
                    const size_t row_off = (line_off + r) * (size_t)p.Cn + n0;
                    const bool row_ok = r < p.RW;
                    for (int c0 = 0; c0 < p.Cs; c0 += 4, ++grp) {
                        float v[7][4];
                        // ---- load the KW column groups of this row; the hi*lo half is added on the fly
#pragma unroll
                        for (int kw = 0; kw < 7; ++kw) {
                            if (kw < p.KW) {
                                tmem_ld4(acc0 + (uint32_t)(kw * p.Cs + c0), v[kw]);
                            }
                        }
                        float v2[7][4];
                        if (two) {
#pragma unroll
                            for (int kw = 0; kw < 7; ++kw)
                                if (kw < p.KW) tmem_ld4(acc0 + (uint32_t)(p.NW + kw * p.Cs + c0), v2[kw]);
                        }
                        tmem_ld_wait();
                        if (two) {
#pragma unroll
                            for (int kw = 0; kw < 7; ++kw)
                                if (kw < p.KW) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) v[kw][j] += v2[kw][j];
                                }
                        }
                        // ---- publish the rows a neighbouring warp needs: row r feeds output voxel r - s (s = shift of tap kw)
                        const int eb = grp & 1;
#pragma unroll
                        for (int kw = 0; kw < 7; ++kw) {
                            if (kw < p.KW) {
                                const int s = p.dgrad ? p.pw - kw : kw - p.pw;
                                if (s > 0 && lane < s) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) s_edge[eb][q][kw][lane][j] = v[kw][j];
                                } else if (s < 0 && lane >= 32 + s) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) s_edge[eb][q][kw][31 - lane][j] = v[kw][j];
                                }
                            }
                        }
                        epi_bar();
                        // ---- out[w] = sum_kw S[w + s_kw][kw]
                        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int kw = 0; kw < 7; ++kw) {
                            if (kw < p.KW) {
                                const int s = p.dgrad ? p.pw - kw : kw - p.pw;
                                const int src = lane + s;
                                float g[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) g[j] = __shfl_sync(0xffffffffu, v[kw][j], src & 31);
                                if (src >= 32) {
                                    if (q < 3) {
#pragma unroll
                                        for (int j = 0; j < 4; ++j) g[j] = s_edge[eb][q + 1][kw][src - 32][j];
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 4; ++j) g[j] = 0.f;
                                    }
                                } else if (src < 0) {
                                    if (q > 0) {
#pragma unroll
                                        for (int j = 0; j < 4; ++j) g[j] = s_edge[eb][q - 1][kw][-1 - src][j];
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 4; ++j) g[j] = 0.f;
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 4; ++j) o[j] += g[j];
                            }
                        }
                        // ---- bias / residual / ReLU, fp32 NDHWC store (+ optional bf16 split planes for the next conv)
                        if (row_ok && c0 < ct) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) o[j] += s_bias[c0 + j];
                            float *dst = p.out + row_off + c0;
                            const float *res = p.residual ? p.residual + row_off + c0 : nullptr;
                            if (vecw == 4 && c0 + 4 <= ct) {
                                if (res) { const float4 rr = __ldg(reinterpret_cast<const float4 *>(res)); o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w; }
                                if (p.relu) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f); }
                                *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                            } else if (vecw >= 2) {
#pragma unroll
                                for (int j = 0; j < 4; j += 2) {
                                    if (c0 + j + 2 <= ct) {
                                        if (res) { const float2 rr = __ldg(reinterpret_cast<const float2 *>(res + j)); o[j] += rr.x; o[j + 1] += rr.y; }
                                        if (p.relu) { o[j] = fmaxf(o[j], 0.f); o[j + 1] = fmaxf(o[j + 1], 0.f); }
                                        *reinterpret_cast<float2 *>(dst + j) = make_float2(o[j], o[j + 1]);
                                    } else if (c0 + j < ct) {
                                        if (res) o[j] += __ldg(res + j);
                                        if (p.relu) o[j] = fmaxf(o[j], 0.f);
                                        dst[j] = o[j];
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (c0 + j < ct) {
                                        if (res) o[j] += __ldg(res + j);
                                        if (p.relu) o[j] = fmaxf(o[j], 0.f);
                                        dst[j] = o[j];
                                    }
                            }
                            if (p.out_split) {
                                // canonical split layout of the OUTPUT tensor: [line][plane][w][Kg]; padded channels (>= Cn) were zeroed by the host
                                __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float x = (c0 + j < ct) ? o[j] : 0.f;
                                    hi[j] = __float2bfloat16_rn(x);
                                    lo[j] = __float2bfloat16_rn(x - __bfloat162float(hi[j]));
                                }
                                const size_t so = ((line_off / p.RW) * p.planes * (size_t)p.RW + r) * (size_t)p.out_split_kg + n0 + c0;
                                *reinterpret_cast<uint2 *>(p.out_split + so) = *reinterpret_cast<const uint2 *>(hi);
                                if (p.planes > 1) *reinterpret_cast<uint2 *>(p.out_split + so + (size_t)p.RW * p.out_split_kg) = *reinterpret_cast<const uint2 *>(lo);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[t]);
            }
        }
        if (prof) { g_tcw_prof[7] += ew; g_tcw_prof[8] += clock64() - e_start; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// weights [Cout, Cin, kd, kh, kw] fp32 -> bf16 planes [ntile][kd*KH + kh][plane][NW rows = (kw, c)][Kg]   (K contiguous)
//   fprop: n = cout, k = cin;   dgrad: n = cin, k = cout.   Rows / columns beyond the real extents are zero.
__global__ void __launch_bounds__(256) pack_weights_tcw_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ dst, int cout, int cin, int KD,
                                                              int KH, int KW, int NT, int CT, int Cs, int NW, int Kg, int planes, int dgrad) {
    const int T2 = KD * KH, T = T2 * KW;
    const long long per_plane = (long long)NW * Kg;
    const long long total = (long long)NT * T2 * per_plane;
    const int Nc = dgrad ? cin : cout, Kc = dgrad ? cout : cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % Kg);
        const int row = (int)((i / Kg) % NW);
        const int tap2 = (int)((i / per_plane) % T2);
        const int nt = (int)(i / (per_plane * T2));
        const int kw = row / Cs, c = row % Cs;
        const int n = nt * CT + c;
        float v = 0.f;
        if (kw < KW && c < CT && n < Nc && k < Kc) {
            const int tap = tap2 * KW + kw;
            v = dgrad ? w[((size_t)k * cin + n) * T + tap] : w[((size_t)n * cin + k) * T + tap];
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const long long o = (((long long)nt * T2 + tap2) * planes) * per_plane + (long long)row * Kg + k;
        dst[o] = hi;
        if (planes > 1) dst[o + per_plane] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

// ------------------------------------------------------------------------------------------------ host side
__global__ void split_rows_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, long long rows, int C, int Cp, int planes, int inter_w,
                                  const float *__restrict__ relu_of, float *__restrict__ masked_out, float *__restrict__ colsum);
int conv_tc_kpad(int channels);

struct TcwPlan {
    bool ok = false;
    int Kc, Kg, Nc, CT, NT, Cs, NW, stacked, TL, ACC, tmem_cols, SA, SB;
    int nchunk, ck0[kTcwMaxChunks], cw[kTcwMaxChunks], a_off[kTcwMaxChunks], b_off[kTcwMaxChunks], tm[kTcwMaxChunks];
    int a_stage, b_tile, smem_bytes, ctas_per_sm;
    int RD, RH, RW, SD, SH, SW;
    long long src_rows;
    TcwSched sch;
};

static bool tcw_build_sched(const ConvGeom &g, bool dgrad, int TL, TcwSched &s, int &live_max) {
    struct P { int rel, t, kh; };
    std::vector<P> ps;
    for (int t = 0; t < TL; ++t)
        for (int kh = 0; kh < g.kh; ++kh) {
            if (!dgrad) ps.push_back({t * g.sh - g.ph + kh, t, kh});
            else {
                const int x = t + g.ph - kh;
                if (((x % g.sh) + g.sh) % g.sh != 0) continue;
                ps.push_back({x / g.sh, t, kh});
            }
        }
    std::sort(ps.begin(), ps.end(), [](const P &a, const P &b) { return a.rel != b.rel ? a.rel < b.rel : a.t < b.t; });
    memset(&s, 0, sizeof(s));
    int first_line[kTcwMaxKH], last_line[kTcwMaxKH], last_pair_of_t[kTcwMaxTL][2];
    for (int k = 0; k < kTcwMaxKH; ++k) first_line[k] = last_line[k] = -1;
    for (int t = 0; t < kTcwMaxTL; ++t) last_pair_of_t[t][0] = last_pair_of_t[t][1] = -1;
    int nl = 0;
    for (size_t i = 0; i < ps.size(); ++i) {
        if (ps[i].rel < -128 || ps[i].rel > 127) return false;
        if (nl == 0 || s.line_rel[nl - 1] != ps[i].rel) {
            if (nl == kTcwMaxLines) return false;
            s.line_rel[nl] = (signed char)ps[i].rel;
            s.npairs[nl] = 0;
            ++nl;
        }
        const int l = nl - 1, j = s.npairs[l];
        if (j >= kTcwMaxTL) return false;
        s.pair_t[l][j] = (unsigned char)ps[i].t;
        s.pair_kh[l][j] = (unsigned char)ps[i].kh;
        s.pair_flags[l][j] = 0;
        ++s.npairs[l];
        if (first_line[ps[i].kh] < 0) first_line[ps[i].kh] = l;
        last_line[ps[i].kh] = l;
        last_pair_of_t[ps[i].t][0] = l;
        last_pair_of_t[ps[i].t][1] = j;
    }
    s.nlines = nl;
    for (int t = 0; t < TL; ++t) {
        if (last_pair_of_t[t][0] < 0) return false;   // an output line that no tap feeds (kernel smaller than the stride): not this kernel's case
        s.pair_flags[last_pair_of_t[t][0]][last_pair_of_t[t][1]] |= kFlagAccLast;
    }
    // weight tiles in first-use order; FIRST / LAST flags on the pairs
    int order = 0;
    bool seen[kTcwMaxKH] = {};
    for (int l = 0; l < nl; ++l)
        for (int j = 0; j < s.npairs[l]; ++j) {
            const int kh = s.pair_kh[l][j];
            if (!seen[kh]) {
                seen[kh] = true;
                s.tile_order[kh] = (unsigned char)order;
                s.load_kh[order] = (unsigned char)kh;
                s.load_line[order] = (unsigned char)l;
                s.pair_flags[l][j] |= kFlagFirst;
                ++order;
            }
        }
    s.ntiles = order;
    for (int kh = 0; kh < g.kh; ++kh) {
        if (last_line[kh] < 0) continue;
        const int l = last_line[kh];
        for (int j = s.npairs[l] - 1; j >= 0; --j)
            if (s.pair_kh[l][j] == kh) { s.pair_flags[l][j] |= kFlagLast; break; }
    }
    // FIFO release requires last uses in the same order as first uses
    for (int a = 0; a + 1 < order; ++a)
        if (last_line[s.load_kh[a]] > last_line[s.load_kh[a + 1]]) return false;
    live_max = 0;
    for (int l = 0; l < nl; ++l) {
        int live = 0;
        for (int kh = 0; kh < g.kh; ++kh)
            if (first_line[kh] >= 0 && first_line[kh] <= l && l <= last_line[kh]) ++live;
        live_max = std::max(live_max, live);
    }
    return true;
}

static TcwPlan make_tcw_plan(const ConvGeom &g, int pass, int planes) {
    TcwPlan pl;
    if (pass != 0 && pass != 1) return pl;
    if (g.sw != 1) return pl;
    if (const char *e = getenv("MDT_TCW")) { if (atoi(e) == 0) return pl; }
    const bool dgrad = pass == 1;
    pl.Kc = dgrad ? g.cout : g.cin;
    pl.Nc = dgrad ? g.cin : g.cout;
    pl.RD = dgrad ? g.d : g.od; pl.RH = dgrad ? g.h : g.oh; pl.RW = dgrad ? g.w : g.ow;
    pl.SD = dgrad ? g.od : g.d; pl.SH = dgrad ? g.oh : g.h; pl.SW = dgrad ? g.ow : g.w;
    if (pl.RW > 128 || pl.SW > 128 || pl.RW <= 64) return pl;      // narrower lines: conv3d_tc.cu packs several lines into one 128-row tile
    if (g.kw > 7 || g.kh > kTcwMaxKH || g.pw > 3 || g.kw - 1 - g.pw > 3 || g.pw > g.kw - 1) return pl;
    if (g.pd > g.kd - 1 || g.ph > g.kh - 1) return pl;
    if (dgrad && (g.kd < g.sd || g.kh < g.sh || g.pd < g.sd - 1 || g.ph < g.sh - 1)) return pl;   // some rows would receive no tap at all
    pl.Kg = conv_tc_kpad(pl.Kc);
    // K chunks: 64-channel (128B swizzle) chunks, then one 32- and / or one 16-channel chunk
    int k = 0, nc = 0;
    while (pl.Kg - k >= 64) { if (nc == kTcwMaxChunks) return pl; pl.ck0[nc] = k; pl.cw[nc] = 64; pl.tm[nc] = 0; ++nc; k += 64; }
    if (pl.Kg - k >= 32) { if (nc == kTcwMaxChunks) return pl; pl.ck0[nc] = k; pl.cw[nc] = 32; pl.tm[nc] = 1; ++nc; k += 32; }
    if (pl.Kg - k >= 16) { if (nc == kTcwMaxChunks) return pl; pl.ck0[nc] = k; pl.cw[nc] = 16; pl.tm[nc] = 2; ++nc; k += 16; }
    if (k != pl.Kg) return pl;
    pl.nchunk = nc;
    // N tiling: (kw, channel) columns of one tile must fit one MMA (N <= 256)
    const int cs_all = ceil_div(pl.Nc, 4) * 4;
    if (g.kw * cs_all <= 256) { pl.CT = pl.Nc; pl.NT = 1; }
    else { pl.CT = (256 / g.kw) & ~3; if (pl.CT < 4) return pl; pl.NT = ceil_div(pl.Nc, pl.CT); }
    pl.Cs = ceil_div(pl.CT, 4) * 4;
    pl.NW = ceil_div(g.kw * pl.Cs, 16) * 16;
    pl.stacked = (planes == 2 && 2 * pl.NW <= 256) ? 1 : 0;
    pl.ACC = pl.stacked ? 2 * pl.NW : pl.NW;
    pl.TL = (2 * pl.ACC <= 512 && pl.RH > 1) ? 2 : 1;
    if (dgrad && g.sh > 1) { if (g.sh != 2 || 2 * pl.ACC > 512) return pl; pl.TL = 2; }
    if (const char *e = getenv("MDT_TCW_TL")) { const int v = atoi(e); if (v == 1 && !(dgrad && g.sh > 1)) pl.TL = 1; }
    int live = 0;
    if (!tcw_build_sched(g, dgrad, pl.TL, pl.sch, live)) return pl;
    pl.tmem_cols = 32;
    while (pl.tmem_cols < pl.TL * pl.ACC) pl.tmem_cols <<= 1;
    int ao = 0, bo = 0;
    for (int c = 0; c < nc; ++c) {
        pl.a_off[c] = ao; pl.b_off[c] = bo;
        ao += planes * 128 * 2 * pl.cw[c];
        bo += planes * pl.NW * 2 * pl.cw[c];
    }
    pl.a_stage = (int)align_up((size_t)ao, 1024);
    pl.b_tile = (int)align_up((size_t)bo, 1024);
    // ring depths: weights need `live` slots (+ prefetch), activations >= 2
    const int budget = 212 * 1024;
    int SA = 4, SB = std::min(kTcwMaxSB, live + 2);
    auto bytes = [&](int sa, int sb) { return sa * pl.a_stage + sb * pl.b_tile; };
    while (bytes(SA, SB) > budget && SA > 3) --SA;
    while (bytes(SA, SB) > budget && SB > live + 1) --SB;
    while (bytes(SA, SB) > budget && SA > 2) --SA;
    while (bytes(SA, SB) > budget && SB > live) --SB;
    if (bytes(SA, SB) > budget || SB < live || SB < 1) return pl;
    if (const char *e = getenv("MDT_TCW_SA")) { const int v = atoi(e); if (v >= 2 && v <= kTcwMaxSA && bytes(v, SB) <= budget) SA = v; }
    if (const char *e = getenv("MDT_TCW_SB")) { const int v = atoi(e); if (v >= live && v <= kTcwMaxSB && bytes(SA, v) <= budget) SB = v; }
    pl.SA = SA; pl.SB = SB;
    pl.smem_bytes = bytes(SA, SB) + 1024;
    pl.ctas_per_sm = (pl.smem_bytes <= 108 * 1024 && pl.tmem_cols <= 256) ? 2 : 1;
    pl.src_rows = (long long)g.n * pl.SD * pl.SH * pl.SW;
    pl.ok = true;
    return pl;
}

// debug: read (and clear) the role-level cycle counters written under MDT_TCW_PROF=1; synchronises the device
int conv_tcw_read_prof(unsigned long long *out16) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyFromSymbol(out16, g_tcw_prof, sizeof(unsigned long long) * 16)) != cudaSuccess) return (int)e;
    unsigned long long z[16] = {};
    return (int)cudaMemcpyToSymbol(g_tcw_prof, z, sizeof(z));
}

bool conv_tcw_supported(const ConvGeom &g, int pass) { return make_tcw_plan(g, pass, 2).ok && tmap_encode_fn() != nullptr; }

static size_t tcw_weight_bytes(const ConvGeom &g, const TcwPlan &pl, int planes) {
    return align_up((size_t)pl.NT * g.kd * g.kh * planes * pl.NW * pl.Kg * 2, 1024);
}

size_t conv_tcw_workspace_bytes(const ConvGeom &g, int pass, int precision) {
    const int planes = precision == 1 ? 1 : 2;
    const TcwPlan pl = make_tcw_plan(g, pass, planes);
    if (!pl.ok) return 0;
    return align_up((size_t)planes * pl.src_rows * pl.Kg * 2, 1024) + tcw_weight_bytes(g, pl, planes) + 2048;
}

// presplit != nullptr: the A operand is already in canonical split form ([line][plane][w][Kg] bf16) and `src` is ignored.
// out_split != nullptr: additionally emit the result in canonical split form (Kg_out = conv_tc_kpad(Nc) channels; the caller zero-fills the
// padding channels once).
int conv_tcw_run(const ConvGeom &g, int pass, const float *src, const float *w, const float *bias, const float *residual, float *dst, int relu,
                 int precision, void *ws, size_t ws_bytes, cudaStream_t st, const __nv_bfloat16 *presplit, __nv_bfloat16 *out_split) {
    const int planes = precision == 1 ? 1 : 2;
    const TcwPlan pl = make_tcw_plan(g, pass, planes);
    if (!pl.ok) return MDT_EUNSUPPORTED;
    if (ws_bytes < conv_tcw_workspace_bytes(g, pass, precision)) return MDT_EWORKSPACE;
    const bool dgrad = pass == 1;
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
    __nv_bfloat16 *xs = presplit ? const_cast<__nv_bfloat16 *>(presplit) : reinterpret_cast<__nv_bfloat16 *>(base);
    __nv_bfloat16 *wp = reinterpret_cast<__nv_bfloat16 *>(base + align_up((size_t)planes * pl.src_rows * pl.Kg * 2, 1024));
    int rc;
    if (!presplit) {
        const long long total = pl.src_rows * (pl.Kg / 8);
        long long blocks = ceil_div<long long>(total, 256);
        if (blocks > (long long)num_sms() * 32) blocks = (long long)num_sms() * 32;
        split_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, xs, pl.src_rows, pl.Kc, pl.Kg, planes, pl.SW, nullptr, nullptr, nullptr);
        if ((rc = launch_status())) return rc;
    }
    {
        const long long wt = (long long)pl.NT * g.kd * g.kh * pl.NW * pl.Kg;
        pack_weights_tcw_kernel<<<(unsigned)ceil_div<long long>(wt, 256), 256, 0, st>>>(w, wp, g.cout, g.cin, g.kd, g.kh, g.kw, pl.NT, pl.CT, pl.Cs, pl.NW,
                                                                                       pl.Kg, planes, dgrad ? 1 : 0);
        if ((rc = launch_status())) return rc;
    }

    TcwParams p{};
    p.NB = g.n; p.RD = pl.RD; p.RH = pl.RH; p.RW = pl.RW; p.SD = pl.SD; p.SH = pl.SH;
    p.KD = g.kd; p.KH = g.kh; p.KW = g.kw; p.sd = g.sd; p.sh = g.sh; p.pd = g.pd; p.ph = g.ph; p.pw = g.pw; p.dgrad = dgrad ? 1 : 0;
    p.Cn = pl.Nc; p.CT = pl.CT; p.Cs = pl.Cs; p.NW = pl.NW; p.stacked = pl.stacked; p.planes = planes;
    p.TL = pl.TL; p.ACC = pl.ACC; p.tmem_cols = pl.tmem_cols;
    p.nchunk = pl.nchunk;
    p.a_tx = 0; p.b_tx = 0;
    for (int c = 0; c < pl.nchunk; ++c) {
        p.ck0[c] = pl.ck0[c]; p.cw[c] = pl.cw[c]; p.a_off[c] = pl.a_off[c]; p.b_off[c] = pl.b_off[c]; p.tm[c] = pl.tm[c];
        p.a_tx += planes * 128 * 2 * pl.cw[c];
        p.b_tx += planes * pl.NW * 2 * pl.cw[c];
    }
    p.a_stage_bytes = pl.a_stage; p.b_tile_bytes = pl.b_tile; p.SA = pl.SA; p.SB = pl.SB;
    p.tiles_h = ceil_div(pl.RH, pl.TL);
    p.total_tiles = (long long)g.n * pl.RD * p.tiles_h;
    p.relu = relu; p.bias = bias; p.residual = residual; p.out = dst;
    p.out_split = out_split; p.out_split_kg = conv_tc_kpad(pl.Nc);
    p.sch = pl.sch;
    if (const char *e = getenv("MDT_TCW_PROF")) p.prof = atoi(e);

    TcwMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int c = 0; c < pl.nchunk; ++c) {
        const int ti = pl.tm[c], cwid = pl.cw[c];
        // A: bf16 [N*SD][SH][plane][SW][Kg]; box {chunk, 128 voxels, planes, 1, 1}
        const uint64_t line = (uint64_t)pl.SW * pl.Kg * 2;
        const uint64_t dims[5] = {(uint64_t)pl.Kg, (uint64_t)pl.SW, (uint64_t)planes, (uint64_t)pl.SH, (uint64_t)g.n * pl.SD};
        const uint64_t strides[4] = {(uint64_t)pl.Kg * 2, line, line * planes, line * planes * pl.SH};
        const uint32_t box[5] = {(uint32_t)cwid, 128u, (uint32_t)planes, 1u, 1u};
        if (!encode_bf16_tmap(&maps.a[ti], xs, 5, dims, strides, box, 2 * cwid)) return MDT_EDRIVER;
        // B: bf16 [NT*KD*KH][plane][NW][Kg]; box {chunk, NW, 1, 1}
        const uint64_t bdims[4] = {(uint64_t)pl.Kg, (uint64_t)pl.NW, (uint64_t)planes, (uint64_t)pl.NT * g.kd * g.kh};
        const uint64_t bstr[3] = {(uint64_t)pl.Kg * 2, (uint64_t)pl.NW * pl.Kg * 2, (uint64_t)planes * pl.NW * pl.Kg * 2};
        const uint32_t bbox[4] = {(uint32_t)cwid, (uint32_t)pl.NW, 1u, 1u};
        if (!encode_bf16_tmap(&maps.b[ti], wp, 4, bdims, bstr, bbox, 2 * cwid)) return MDT_EDRIVER;
    }

    static bool attr[kMaxDevices] = {};
    if (!ensure_smem_attr(conv_tcw_kernel, 220 * 1024, attr)) return MDT_EDRIVER;
    long long gx = (long long)num_sms() * pl.ctas_per_sm;
    if (gx > p.total_tiles) gx = p.total_tiles;
    dim3 grid((unsigned)gx, (unsigned)pl.NT);
    conv_tcw_kernel<<<grid, kTcwThreads, (size_t)pl.smem_bytes, st>>>(maps, p);
    return launch_status();
}

}  // namespace mdt
