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
// so one MMA has N = kw * C columns (108 -> 112 for 36 channels; 128 with the hi/lo weight planes stacked as well for 18 channels).
// A CTA owns TL = 2 adjacent output lines with one TMEM accumulator each: a source line is used by both (different kh), so activation
// lines are fetched (TL + KH - 1) / TL times per output line instead of KH times, and the (kd, kh) weight tiles flow through a FIFO ring
// in first-use order and serve both lines.  K is padded to 16 (36 -> 48: three K steps, not four) by splitting it into swizzle-width
// chunks (64 / 32 / 16 channels = 128B / 64B / 32B swizzle), each with its own tensor map over the same bf16 planes.
// Persistent CTAs (static round-robin over tiles); warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.  Two accumulator
// sets when TMEM allows (2 * TL * ACC <= 512 columns): the epilogue of tile i overlaps the MMAs of tile i + 1.
// The single-thread issue loop and the per-warp epilogue are instruction-bound if written naively (first version, profiles/
// r02_tcw_v1_prof.txt: 300 cycles per MMA, 1900 cycles per 4-channel epilogue group): descriptors are base + precomputed offsets, ring
// slots come from multiply-high instead of divisions, the epilogue is specialised on (KW, TMEM load width).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv3d_common.cuh"
#include "conv3d_tc_plan.cuh"
#include "tc_common.cuh"

namespace mdt {
using namespace tc;

constexpr int kTcwThreads = 352;   // warp 0 producer, warps 1-2 MMA issuers (one per output-line accumulator), warps 3..10 epilogue (two per TMEM lane quarter)
constexpr int kTcwMaxChunks = 4;
constexpr int kTcwMaxOps = 12;      // K16 steps per (line, tap) pair: Kg <= 192
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
    unsigned int pair_word[kTcwMaxLines][kTcwMaxTL];   // t | kh << 8 | flags << 16 | (position of tile kh in the load order) << 24
    unsigned char load_kh[kTcwMaxKH];      // position -> kh
    unsigned char load_line[kTcwMaxKH];    // position -> schedule line before which the tile is fetched
};

struct TcwParams {
    int NB, RD, RH, RW, SD, SH;
    int KD, KH, KW, sd, sh, pd, ph, pw, dgrad;
    int Cn, CT, Cs, NW, stacked, planes, ldw;
    int TL, ACC, nbuf, tmem_cols;
    int nchunk, ck0[kTcwMaxChunks], cw[kTcwMaxChunks], cks[kTcwMaxChunks], a_off[kTcwMaxChunks], b_off[kTcwMaxChunks], tm[kTcwMaxChunks];   // cks: K16 steps issued
    int nops;
    unsigned int op_hi[kTcwMaxOps];                                            // high word of the shared-memory matrix descriptor (SBO, version, swizzle)
    unsigned short op_a[kTcwMaxOps], op_alo[kTcwMaxOps], op_b[kTcwMaxOps], op_blo[kTcwMaxOps];   // operand offsets inside a stage / tile, 16-byte units
    int a_stage_bytes, b_tile_bytes, SA, SB, a_tx, b_tx;
    unsigned int inv_sa, inv_sb;                                               // floor(2^32 / ring depth) + 1
    int tiles_h;
    long long total_tiles;
    int relu;
    const float *bias, *residual;
    float *out;
    __nv_bfloat16 *out_split;   // optional: the result also as (hi, lo) bf16 planes in the canonical split layout [line][plane][w][Kg_out]
    int out_split_kg;
    int prof;
    int skip;                   // diagnostics (MDT_TCW_SKIP): 1 = no MMAs issued (TMA pipeline alone), 2 = no TMA loads (MMA + epilogue alone), 3 = no epilogue work, 4 = neither TMA nor epilogue (MMA issue + tensor pipe alone); results are garbage
    TcwSched sch;
};

struct TcwMaps {
    CUtensorMap a[3], b[3];   // index 0 / 1 / 2 = 64 / 32 / 16-channel chunks (128B / 64B / 32B swizzle)
};

// role-level cycle counters of CTA 0 (MDT_TCW_PROF=1): [0] producer wait a_empty, [1] producer wait b_empty, [2] producer total,
// [3] mma wait a_full, [4] mma wait b_full, [5] mma wait acc_empty, [6] mma total, [7] epilogue wait acc_full, [8] epilogue total, [9] tiles
__device__ unsigned long long g_tcw_prof[16];
#define TCW_T0(flag) const long long _t0 = (flag) ? clock64() : 0
#define TCW_ACC(flag, var) do { if (flag) var += clock64() - _t0; } while (0)

template <int W>
__device__ __forceinline__ void tmem_ldw(uint32_t taddr, float *v);
template <>
__device__ __forceinline__ void tmem_ldw<4>(uint32_t taddr, float *v) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
template <>
__device__ __forceinline__ void tmem_ldw<2>(uint32_t taddr, float *v) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1);
}

__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// D[tmem] (+)= A * B with the descriptors given as (lo, hi) words
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

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

// One output line: TMEM accumulator -> (tap shifts along w) -> bias / residual / ReLU -> fp32 NDHWC (+ optional bf16 split planes).
// Thread = accumulator row r = source voxel index; it produces output voxel w = r:  out[w] = sum_kw S[w + s_kw][kw],  s_kw = kw - pw (fprop)
// or pw - kw (dgrad).  Rows held by a neighbouring warp travel through s_edge (at most 3 rows per side), all others by warp shuffles.
template <int KW, int CH, int LDW>
__device__ __forceinline__ void tcw_epilogue_line(const TcwParams &p, uint32_t acc0, int q, int lane, int ct, int n0, size_t line_idx, bool row_ok, int r,
                                                  int vecw, const float *s_bias, float (*s_edge)[4][7][3][8], uint32_t &grp, int half) {
    int sh[KW];
#pragma unroll
    for (int kw = 0; kw < KW; ++kw) sh[kw] = p.dgrad ? p.pw - kw : kw - p.pw;
    const bool two = p.stacked != 0;
    const size_t row_off = (line_idx * (size_t)p.RW + r) * (size_t)p.Cn + n0;
    // the two warps of a lane quarter take alternate channel groups; each half has its own edge buffers and named barrier
    for (int c0 = half * CH; c0 < p.Cs; c0 += 2 * CH, ++grp) {
        float v[KW][CH];
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) {
#pragma unroll
            for (int jj = 0; jj < CH; jj += LDW) {
                if (c0 + jj < p.Cs) {
                    tmem_ldw<LDW>(acc0 + (uint32_t)(kw * p.Cs + c0 + jj), &v[kw][jj]);
                } else {
#pragma unroll
                    for (int j = 0; j < LDW; ++j) v[kw][jj + j] = 0.f;
                }
            }
        }
        if (two) {
            float v2[KW][CH];
#pragma unroll
            for (int kw = 0; kw < KW; ++kw) {
#pragma unroll
                for (int jj = 0; jj < CH; jj += LDW) {
                    if (c0 + jj < p.Cs) {
                        tmem_ldw<LDW>(acc0 + (uint32_t)(p.NW + kw * p.Cs + c0 + jj), &v2[kw][jj]);
                    } else {
#pragma unroll
                        for (int j = 0; j < LDW; ++j) v2[kw][jj + j] = 0.f;
                    }
                }
            }
            tmem_ld_wait();
#pragma unroll
            for (int kw = 0; kw < KW; ++kw) {
#pragma unroll
                for (int j = 0; j < CH; ++j) v[kw][j] += v2[kw][j];
            }
        } else {
            tmem_ld_wait();
        }
        const int eb = grp & 1;
        if (KW > 1) {
#pragma unroll
            for (int kw = 0; kw < KW; ++kw) {
                const int s = sh[kw];
                if (s > 0) {
                    if (lane < s) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) s_edge[eb][q][kw][lane][j] = v[kw][j];
                    }
                } else if (s < 0) {
                    if (lane >= 32 + s) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) s_edge[eb][q][kw][31 - lane][j] = v[kw][j];
                    }
                }
            }
            epi_bar(1 + half);
        }
        float o[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) o[j] = 0.f;
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) {
            const int s = sh[kw];
            if (s == 0) {
#pragma unroll
                for (int j = 0; j < CH; ++j) o[j] += v[kw][j];
            } else {
                const int src = lane + s;
                float g[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) g[j] = __shfl_sync(0xffffffffu, v[kw][j], src & 31);
                if (src >= 32) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) g[j] = q < 3 ? s_edge[eb][q + 1][kw][src - 32][j] : 0.f;
                } else if (src < 0) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) g[j] = q > 0 ? s_edge[eb][q - 1][kw][-1 - src][j] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) o[j] += g[j];
            }
        }
        if (row_ok && c0 < ct) {
#pragma unroll
            for (int j = 0; j < CH; ++j) o[j] += s_bias[c0 + j];
            float *dst = p.out + row_off + c0;
            const float *res = p.residual ? p.residual + row_off + c0 : nullptr;
#pragma unroll
            for (int j4 = 0; j4 < CH; j4 += 4) {
                if (c0 + j4 < ct) {
                    float *o4 = o + j4;
                    if (vecw == 4 && c0 + j4 + 4 <= ct) {
                        if (res) { const float4 rr = __ldg(reinterpret_cast<const float4 *>(res + j4)); o4[0] += rr.x; o4[1] += rr.y; o4[2] += rr.z; o4[3] += rr.w; }
                        if (p.relu) { o4[0] = fmaxf(o4[0], 0.f); o4[1] = fmaxf(o4[1], 0.f); o4[2] = fmaxf(o4[2], 0.f); o4[3] = fmaxf(o4[3], 0.f); }
                        *reinterpret_cast<float4 *>(dst + j4) = make_float4(o4[0], o4[1], o4[2], o4[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; j += 2) {
                            if (vecw >= 2 && c0 + j4 + j + 2 <= ct) {
                                if (res) { const float2 rr = __ldg(reinterpret_cast<const float2 *>(res + j4 + j)); o4[j] += rr.x; o4[j + 1] += rr.y; }
                                if (p.relu) { o4[j] = fmaxf(o4[j], 0.f); o4[j + 1] = fmaxf(o4[j + 1], 0.f); }
                                *reinterpret_cast<float2 *>(dst + j4 + j) = make_float2(o4[j], o4[j + 1]);
                            } else {
#pragma unroll
                                for (int e = 0; e < 2; ++e)
                                    if (c0 + j4 + j + e < ct) {
                                        if (res) o4[j + e] += __ldg(res + j4 + j + e);
                                        if (p.relu) o4[j + e] = fmaxf(o4[j + e], 0.f);
                                        dst[j4 + j + e] = o4[j + e];
                                    }
                            }
                        }
                    }
                    if (p.out_split) {
                        // canonical split layout of the OUTPUT tensor: [line][plane][w][Kg]; channels >= Cn stay as the caller zero-filled them
                        __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float x = (c0 + j4 + j < ct) ? o4[j] : 0.f;
                            hi[j] = __float2bfloat16_rn(x);
                            lo[j] = __float2bfloat16_rn(x - __bfloat162float(hi[j]));
                        }
                        const size_t so = ((line_idx * p.planes) * (size_t)p.RW + r) * (size_t)p.out_split_kg + n0 + c0 + j4;
                        *reinterpret_cast<uint2 *>(p.out_split + so) = *reinterpret_cast<const uint2 *>(hi);
                        if (p.planes > 1) *reinterpret_cast<uint2 *>(p.out_split + so + (size_t)p.RW * p.out_split_kg) = *reinterpret_cast<const uint2 *>(lo);
                    }
                }
            }
        }
    }
    // padding channels of the split output (36 -> 48: channels 36..47): zeros, written by the last N tile
    if (p.out_split && row_ok && half == 0 && n0 + p.CT >= p.Cn) {
        const size_t so0 = ((line_idx * p.planes) * (size_t)p.RW + r) * (size_t)p.out_split_kg;
        for (int c = ((p.Cn + 3) / 4) * 4; c < p.out_split_kg; c += 4) {
            *reinterpret_cast<uint2 *>(p.out_split + so0 + c) = make_uint2(0u, 0u);
            if (p.planes > 1) *reinterpret_cast<uint2 *>(p.out_split + so0 + (size_t)p.RW * p.out_split_kg + c) = make_uint2(0u, 0u);
        }
    }
}

// All MMAs of one (source line, tap) pair, issued by the elected lane.  Every chunk is 64 channels wide (128B swizzle), so operand offsets are
// compile-time: op o -> chunk o / 4, K step o % 4.  NOPS = 0 selects the run-time loop.
template <int NOPS, int MODE>
__device__ __forceinline__ void tcw_issue_pair(uint32_t d_tmem, uint32_t a16, uint32_t b16, uint32_t chunk_a16, uint32_t chunk_b16, uint32_t alo16,
                                               uint32_t blo16, uint32_t hi, uint32_t idescN, uint32_t idesc2N, uint32_t acc, int nops_rt) {
    const uint32_t lbo = 1u << 16;
    const int n = NOPS > 0 ? NOPS : nops_rt;
#pragma unroll
    for (int o = 0; o < n; ++o) {
        const uint32_t oa = (uint32_t)(o >> 2) * chunk_a16 + (uint32_t)(o & 3) * 2u, ob = (uint32_t)(o >> 2) * chunk_b16 + (uint32_t)(o & 3) * 2u;
        const uint32_t a_hi = (a16 + oa) | lbo, b_hi = (b16 + ob) | lbo;
        if (MODE == 2) {
            umma2(d_tmem, a_hi, b_hi, hi, idesc2N, acc);                            // [0,NW) += hi*hi, [NW,2NW) += hi*lo
            umma2(d_tmem, (a16 + oa + alo16) | lbo, b_hi, hi, idescN, 1);           // [0,NW) += lo*hi
        } else {
            umma2(d_tmem, a_hi, b_hi, hi, idescN, acc);
            if (MODE == 3) {
                umma2(d_tmem, a_hi, (b16 + ob + blo16) | lbo, hi, idescN, 1);
                umma2(d_tmem, (a16 + oa + alo16) | lbo, b_hi, hi, idescN, 1);
            }
        }
        acc = 1;
    }
}

template <int MODE>
__device__ __forceinline__ void tcw_issue_pair_n(int nops, uint32_t d_tmem, uint32_t a16, uint32_t b16, uint32_t chunk_a16, uint32_t chunk_b16, uint32_t alo16,
                                                 uint32_t blo16, uint32_t hi, uint32_t idescN, uint32_t idesc2N, uint32_t acc) {
    switch (nops) {
        case 1: tcw_issue_pair<1, MODE>(d_tmem, a16, b16, chunk_a16, chunk_b16, alo16, blo16, hi, idescN, idesc2N, acc, nops); break;
        case 2: tcw_issue_pair<2, MODE>(d_tmem, a16, b16, chunk_a16, chunk_b16, alo16, blo16, hi, idescN, idesc2N, acc, nops); break;
        case 3: tcw_issue_pair<3, MODE>(d_tmem, a16, b16, chunk_a16, chunk_b16, alo16, blo16, hi, idescN, idesc2N, acc, nops); break;
        case 4: tcw_issue_pair<4, MODE>(d_tmem, a16, b16, chunk_a16, chunk_b16, alo16, blo16, hi, idescN, idesc2N, acc, nops); break;
        case 5: tcw_issue_pair<5, MODE>(d_tmem, a16, b16, chunk_a16, chunk_b16, alo16, blo16, hi, idescN, idesc2N, acc, nops); break;
        default: tcw_issue_pair<0, MODE>(d_tmem, a16, b16, chunk_a16, chunk_b16, alo16, blo16, hi, idescN, idesc2N, acc, nops); break;
    }
}

__global__ void __launch_bounds__(kTcwThreads, 1)
conv_tcw_kernel(const __grid_constant__ TcwMaps maps, const __grid_constant__ TcwParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t a_full[kTcwMaxSA], a_empty[kTcwMaxSA], b_full[kTcwMaxSB], b_empty[kTcwMaxSB], acc_full[2][kTcwMaxTL], acc_empty[2][kTcwMaxTL];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_bias[256 + 8];
    __shared__ float s_edge[2][2][4][7][3][8];   // [half][double buffer][quarter][kw][row][channel]

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + (size_t)p.SA * p.a_stage_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * p.CT;
    const int ct = min(p.CT, p.Cn - n0);   // channels of this N tile

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], p.TL); }      // every MMA issuer releases every stage
        for (int i = 0; i < p.SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], p.TL); }
        for (int b = 0; b < 2; ++b)
            for (int i = 0; i < p.TL; ++i) { mbar_init(&acc_full[b][i], 1); mbar_init(&acc_empty[b][i], 8); }
        fence_barrier_init();
        for (int c = 0; c < p.nchunk; ++c) { prefetch_tmap(&maps.a[p.tm[c]]); prefetch_tmap(&maps.b[p.tm[c]]); }
    }
    for (int c = threadIdx.x; c < 256 + 8; c += kTcwThreads) s_bias[c] = (p.bias && c < ct) ? __ldg(p.bias + n0 + c) : 0.f;
    if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // =============================================================== TMA producer: one elected lane, a single elect region (see the MMA issuer)
        if (elect_one()) {
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
                            const uint32_t qd = __umulhi(b_seq, p.inv_sb), slot = b_seq - qd * p.SB;
                            { TCW_T0(prof); mbar_wait(&b_empty[slot], (qd & 1) ^ 1); TCW_ACC(prof, pw_b); }
                            uint8_t *bt = smem_b + (size_t)slot * p.b_tile_bytes;
                            if (p.skip == 2 || p.skip == 4) mbar_arrive(&b_full[slot]);
                            else {
                                mbar_arrive_expect_tx(&b_full[slot], (uint32_t)p.b_tx);
                                for (int c = 0; c < p.nchunk; ++c)
                                    for (int pl = 0; pl < p.planes; ++pl)
                                        tma_load_4d(bt + p.b_off[c] + (size_t)pl * p.NW * 2 * p.cw[c], &maps.b[p.tm[c]], &b_full[slot], p.ck0[c], 0, pl,
                                                    tap_base + kh);
                            }
                            ++b_seq;
                            ++next_load;
                        }
                        const uint32_t qd = __umulhi(a_seq, p.inv_sa), slot = a_seq - qd * p.SA;
                        { TCW_T0(prof); mbar_wait(&a_empty[slot], (qd & 1) ^ 1); TCW_ACC(prof, pw_a); }
                        uint8_t *as = smem_a + (size_t)slot * p.a_stage_bytes;
                        // one box = both planes of the source line: {chunk, 128 voxels, planes, 1, 1}; rows past the line end and lines outside
                        // the image are TMA zero fill (= the conv's zero padding)
                        if (p.skip == 2 || p.skip == 4) mbar_arrive(&a_full[slot]);
                        else {
                            mbar_arrive_expect_tx(&a_full[slot], (uint32_t)p.a_tx);
                            for (int c = 0; c < p.nchunk; ++c)
                                tma_load_5d(as + p.a_off[c], &maps.a[p.tm[c]], &a_full[slot], p.ck0[c], 0, 0, line_base + p.sch.line_rel[i], nb * p.SD + d_src);
                        }
                        ++a_seq;
                    }
                }
            }
            if (prof) { g_tcw_prof[0] += pw_a; g_tcw_prof[1] += pw_b; g_tcw_prof[2] += clock64() - p_start; }
        }
        __syncwarp();
    } else if (warp <= 2) {
        // =============================================================== MMA issuers: warp 1 owns output line t = 0, warp 2 line t = 1 (idle when TL = 1).
        // Two issuing threads because one thread's fixed costs per pipeline event (a barrier wait ~230 cycles even when complete, a
        // tcgen05.commit ~80, profiles/r02_tcw_skip.txt + r02_mma_pipe_probe.txt) exceed the MMA time of the event: with one issuer the tensor pipe
        // sat idle 57 % of the time; two issuers overlap each other's waits.  Each walks the same schedule, issues only the pairs of its own
        // line and arrives on every release barrier (count = TL).
        // ONE elected lane runs the whole role inside a single elect
        // region.  Measured (tools/mma_pipe_probe.cu, profiles/r02_mma_pipe_probe.txt): every separate `if (elect) { tcgen05.mma ... }` region
        // costs ~220-300 dead cycles after its last MMA (+80 per tcgen05.commit) during which the queued MMAs drain — 48 such regions per tile
        // were 45 % of this kernel's time; inside one region MMAs issue back to back at the operand-bandwidth rate.
        const uint32_t my_t = (uint32_t)(warp - 1);
        if (my_t < (uint32_t)p.TL && elect_one()) {
            const uint32_t idescN = make_idesc_bf16(128, p.NW, 0, 0);
            const uint32_t idesc2N = make_idesc_bf16(128, 2 * p.NW, 0, 0);
            const uint32_t sa16 = smem_u32(smem_a) >> 4, sb16 = smem_u32(smem_b) >> 4;
            const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4, b_tile16 = (uint32_t)p.b_tile_bytes >> 4;
            const int mode = p.stacked ? 2 : (p.planes > 1 ? 3 : 1);
            // all chunks are 64 channels x 128B swizzle: chunk strides / plane offsets in 16-byte units, one descriptor high word
            // (a single narrower chunk — 32 or 16 channels, 64B / 32B swizzle — under MDT_TCW_NARROW: same formulas with its row width)
            const uint32_t swz = 2u * (uint32_t)p.cw[0];
            const uint32_t chunk_a16 = ((uint32_t)p.planes * 128u * swz) >> 4, chunk_b16 = ((uint32_t)(p.planes * p.NW) * swz) >> 4;
            const uint32_t alo16 = (128u * swz) >> 4, blo16 = ((uint32_t)p.NW * swz) >> 4;
            const uint32_t desc_hi = ((8u * swz) >> 4) | (1u << 14) | (layout_type_for_swizzle_bytes((int)swz) << 29);
            uint32_t a_seq = 0, b_seq = 0, it = 0;
            const bool prof = p.prof && blockIdx.x == 0 && blockIdx.y == 0 && my_t == 0;
            long long mw_a = 0, mw_b = 0, mw_c = 0;
            const long long m_start = prof ? clock64() : 0;
            for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int rd = (int)((tile / p.tiles_h) % p.RD);
                int kd_last = -1, d_src;
                for (int kd = 0; kd < p.KD; ++kd)
                    if (tcw_depth(p, rd, kd, d_src)) kd_last = kd;
                const uint32_t buf = p.nbuf > 1 ? (it & 1u) : 0u;
                const uint32_t bpar = p.nbuf > 1 ? ((it >> 1) & 1u) : (it & 1u);
                uint32_t acc_started = 0;
                for (int kd = 0; kd < p.KD; ++kd) {
                    if (!tcw_depth(p, rd, kd, d_src)) continue;
                    const bool last_kd = kd == kd_last;
                    for (int i = 0; i < p.sch.nlines; ++i) {
                        const uint32_t qa = __umulhi(a_seq, p.inv_sa), slot_a = a_seq - qa * p.SA;
                        { TCW_T0(prof); mbar_wait(&a_full[slot_a], qa & 1); TCW_ACC(prof, mw_a); }
                        tc_fence_after();
                        const uint32_t a16 = sa16 + slot_a * a_stage16;
                        const int np = p.sch.npairs[i];
                        for (int j = 0; j < np; ++j) {
                            const uint32_t pwd = p.sch.pair_word[i][j];
                            const uint32_t t = pwd & 0xffu, fl = (pwd >> 16) & 0xffu;
                            const uint32_t bs = b_seq + (pwd >> 24);
                            const uint32_t qb = __umulhi(bs, p.inv_sb), slot_b = bs - qb * p.SB;
                            if (t == my_t) {
                                { TCW_T0(prof); mbar_wait(&b_full[slot_b], qb & 1); TCW_ACC(prof, mw_b); tc_fence_after(); }
                                uint32_t acc = acc_started;
                                if (!acc) { TCW_T0(prof); mbar_wait(&acc_empty[buf][t], bpar ^ 1); TCW_ACC(prof, mw_c); tc_fence_after(); }
                                const uint32_t b16 = sb16 + slot_b * b_tile16;
                                const uint32_t d_tmem = tmem + (buf * (uint32_t)p.TL + t) * (uint32_t)p.ACC;
                                if (p.skip == 1) { if (!acc) umma2(d_tmem, a16 | (1u << 16), b16 | (1u << 16), desc_hi, idescN, 0); }
                                else if (mode == 2) tcw_issue_pair_n<2>(p.nops, d_tmem, a16, b16, chunk_a16, chunk_b16, alo16, blo16, desc_hi, idescN, idesc2N, acc);
                                else if (mode == 3) tcw_issue_pair_n<3>(p.nops, d_tmem, a16, b16, chunk_a16, chunk_b16, alo16, blo16, desc_hi, idescN, idesc2N, acc);
                                else tcw_issue_pair_n<1>(p.nops, d_tmem, a16, b16, chunk_a16, chunk_b16, alo16, blo16, desc_hi, idescN, idesc2N, acc);
                                acc_started = 1;
                                if (last_kd && (fl & kFlagAccLast)) umma_commit(&acc_full[buf][t]);
                            }
                            if (fl & kFlagLast) umma_commit(&b_empty[slot_b]);     // both issuers, at the tile's last use in the schedule
                        }
                        umma_commit(&a_empty[slot_a]);
                        ++a_seq;
                    }
                    b_seq += p.sch.ntiles;
                }
            }
            if (prof) { g_tcw_prof[3] += mw_a; g_tcw_prof[4] += mw_b; g_tcw_prof[5] += mw_c; g_tcw_prof[6] += clock64() - m_start; g_tcw_prof[9] += it; }
        }
        __syncwarp();
    } else {
        // =============================================================== epilogue (warps 3..10: TMEM lane quarter = warp % 4, two warps per quarter)
        const int q = warp & 3;
        const int half = (warp - 3) >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const int vecw = (p.Cn % 4 == 0 && n0 % 4 == 0) ? 4 : ((p.Cn % 2 == 0 && n0 % 2 == 0) ? 2 : 1);
        uint32_t it = 0, grp = 0;
        const bool prof = p.prof && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 96;
        long long ew = 0;
        const long long e_start = prof ? clock64() : 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            long long u = tile;
            const int th = (int)(u % p.tiles_h); u /= p.tiles_h;
            const int rd = (int)(u % p.RD);
            const int nb = (int)(u / p.RD);
            const uint32_t buf = p.nbuf > 1 ? (it & 1u) : 0u;
            const uint32_t bpar = p.nbuf > 1 ? ((it >> 1) & 1u) : (it & 1u);
            for (int t = 0; t < p.TL; ++t) {
                { TCW_T0(prof); mbar_wait(&acc_full[buf][t], bpar); TCW_ACC(prof, ew); }
                tc_fence_after();
                const int rh = th * p.TL + t;
                if (rh < p.RH && p.skip < 3) {
                    const uint32_t acc0 = lane_base + (buf * (uint32_t)p.TL + (uint32_t)t) * (uint32_t)p.ACC;
                    const size_t line_idx = ((size_t)nb * p.RD + rd) * p.RH + rh;
                    const bool row_ok = r < p.RW;
                    if (p.ldw == 4) {
                        if (p.KW == 3) tcw_epilogue_line<3, 8, 4>(p, acc0, q, lane, ct, n0, line_idx, row_ok, r, vecw, s_bias, s_edge[half], grp, half);
                        else if (p.KW == 7) tcw_epilogue_line<7, 4, 4>(p, acc0, q, lane, ct, n0, line_idx, row_ok, r, vecw, s_bias, s_edge[half], grp, half);
                        else tcw_epilogue_line<1, 8, 4>(p, acc0, q, lane, ct, n0, line_idx, row_ok, r, vecw, s_bias, s_edge[half], grp, half);
                    } else {
                        if (p.KW == 3) tcw_epilogue_line<3, 8, 2>(p, acc0, q, lane, ct, n0, line_idx, row_ok, r, vecw, s_bias, s_edge[half], grp, half);
                        else if (p.KW == 7) tcw_epilogue_line<7, 4, 2>(p, acc0, q, lane, ct, n0, line_idx, row_ok, r, vecw, s_bias, s_edge[half], grp, half);
                        else tcw_epilogue_line<1, 8, 2>(p, acc0, q, lane, ct, n0, line_idx, row_ok, r, vecw, s_bias, s_edge[half], grp, half);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf][t]);
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
    int Kc, Kg, Nc, CT, NT, Cs, NW, stacked, ldw, TL, ACC, nbuf, tmem_cols, SA, SB;
    int nchunk, ck0[kTcwMaxChunks], cw[kTcwMaxChunks], cks[kTcwMaxChunks], a_off[kTcwMaxChunks], b_off[kTcwMaxChunks], tm[kTcwMaxChunks];
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
    int pair_t[kTcwMaxLines][kTcwMaxTL], pair_kh[kTcwMaxLines][kTcwMaxTL], pair_fl[kTcwMaxLines][kTcwMaxTL];
    int first_line[kTcwMaxKH], last_line[kTcwMaxKH], last_pair_of_t[kTcwMaxTL][2], tile_order[kTcwMaxKH];
    for (int k = 0; k < kTcwMaxKH; ++k) first_line[k] = last_line[k] = -1, tile_order[k] = 0;
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
        pair_t[l][j] = ps[i].t; pair_kh[l][j] = ps[i].kh; pair_fl[l][j] = 0;
        ++s.npairs[l];
        if (first_line[ps[i].kh] < 0) first_line[ps[i].kh] = l;
        last_line[ps[i].kh] = l;
        last_pair_of_t[ps[i].t][0] = l;
        last_pair_of_t[ps[i].t][1] = j;
    }
    s.nlines = nl;
    for (int t = 0; t < TL; ++t) {
        if (last_pair_of_t[t][0] < 0) return false;   // an output line that no tap feeds (kernel smaller than the stride): not this kernel's case
        pair_fl[last_pair_of_t[t][0]][last_pair_of_t[t][1]] |= kFlagAccLast;
    }
    // weight tiles in first-use order; FIRST / LAST flags on the pairs
    int order = 0;
    bool seen[kTcwMaxKH] = {};
    for (int l = 0; l < nl; ++l)
        for (int j = 0; j < s.npairs[l]; ++j) {
            const int kh = pair_kh[l][j];
            if (!seen[kh]) {
                seen[kh] = true;
                tile_order[kh] = order;
                s.load_kh[order] = (unsigned char)kh;
                s.load_line[order] = (unsigned char)l;
                pair_fl[l][j] |= kFlagFirst;
                ++order;
            }
        }
    s.ntiles = order;
    for (int kh = 0; kh < g.kh; ++kh) {
        if (last_line[kh] < 0) continue;
        const int l = last_line[kh];
        for (int j = s.npairs[l] - 1; j >= 0; --j)
            if (pair_kh[l][j] == kh) { pair_fl[l][j] |= kFlagLast; break; }
    }
    // FIFO release requires last uses in the same order as first uses
    for (int a = 0; a + 1 < order; ++a)
        if (last_line[s.load_kh[a]] > last_line[s.load_kh[a + 1]]) return false;
    for (int l = 0; l < nl; ++l)
        for (int j = 0; j < s.npairs[l]; ++j)
            s.pair_word[l][j] = (unsigned)pair_t[l][j] | ((unsigned)pair_kh[l][j] << 8) | ((unsigned)pair_fl[l][j] << 16) | ((unsigned)tile_order[pair_kh[l][j]] << 24);
    live_max = 0;
    for (int l = 0; l < nl; ++l) {
        int live = 0;
        for (int kh = 0; kh < g.kh; ++kh)
            if (first_line[kh] >= 0 && first_line[kh] <= l && l <= last_line[kh]) ++live;
        live_max = std::max(live_max, live);
    }
    return true;
}

static int tcw_env(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

static TcwPlan make_tcw_plan(const ConvGeom &g, int pass, int planes) {
    TcwPlan pl;
    if (pass != 0 && pass != 1) return pl;
    if (g.sw != 1) return pl;
    if (tcw_env("MDT_TCW", 1) == 0) return pl;
    const bool dgrad = pass == 1;
    pl.Kc = dgrad ? g.cout : g.cin;
    pl.Nc = dgrad ? g.cin : g.cout;
    pl.RD = dgrad ? g.d : g.od; pl.RH = dgrad ? g.h : g.oh; pl.RW = dgrad ? g.w : g.ow;
    pl.SD = dgrad ? g.od : g.d; pl.SH = dgrad ? g.oh : g.h; pl.SW = dgrad ? g.ow : g.w;
    if (pl.RW > 128 || pl.SW > 128 || pl.RW <= 64) return pl;      // narrower lines: conv3d_tc.cu packs several lines into one 128-row tile
    if (g.kw != 1 && g.kw != 3 && g.kw != 7) return pl;            // epilogue instances
    if (g.kh > kTcwMaxKH || g.pw > 3 || g.kw - 1 - g.pw > 3 || g.pw > g.kw - 1) return pl;
    if (g.pd > g.kd - 1 || g.ph > g.kh - 1) return pl;
    if (dgrad && (g.kd < g.sd || g.kh < g.sh || g.pd < g.sd - 1 || g.ph < g.sh - 1)) return pl;   // some rows would receive no tap at all
    pl.Kg = conv_tc_kpad(pl.Kc);
    // K chunks.  Measured (profiles/r02_tcw_prof_v4.txt): MMAs whose operands sit in 64B / 32B-swizzled tiles run 2-5x slower than with 128B
    // swizzle, so every chunk is staged as 64 channels x 128B swizzle; a chunk that runs past Kg (36 -> Kg = 48) is zero-filled by the TMA
    // (no L2 traffic for it) and only its first Kg/16 K-steps are issued.
    int k = 0, nc = 0;
    const bool narrow = tcw_env("MDT_TCW_NARROW", 1) != 0;   // measured: 18 -> 18 k7 fprop 1.69 -> 1.41 ms (the zero-padded rows made it TMA-bound)
    while (k < pl.Kg) {
        if (nc == kTcwMaxChunks) return pl;
        const int rem = pl.Kg - k;
        pl.ck0[nc] = k; pl.cw[nc] = 64; pl.cks[nc] = std::min(4, rem / 16); pl.tm[nc] = 0;
        // a LAST chunk of exactly 32 or 16 channels may be staged at its own width (64B / 32B swizzle): half / quarter the shared-memory fill
        // of the zero-padded 128-byte rows (MDT_TCW_NARROW=0 switches back; A/B knob)
        if (narrow && nc == 0 && (rem == 32 || rem == 16)) { pl.cw[nc] = rem; pl.tm[nc] = rem == 32 ? 1 : 2; }
        ++nc; k += 64;
    }
    k = pl.Kg;
    if (k != pl.Kg || pl.Kg / 16 > kTcwMaxOps) return pl;
    pl.nchunk = nc;
    // N tiling: (kw, channel) columns of one tile must fit one MMA (N <= 256).  Channel stride inside the stacked column index: a multiple of
    // the TMEM load width (4 columns; 2 when the channel count is 2 mod 4 and the narrower stride saves MMA columns, e.g. 7 x 18 = 126 -> 128)
    const int cs4 = ceil_div(pl.Nc, 4) * 4, cs2 = ceil_div(pl.Nc, 2) * 2;
    if (g.kw * cs4 <= 256 || g.kw * cs2 <= 256) { pl.CT = pl.Nc; pl.NT = 1; }
    else { pl.CT = (256 / g.kw) & ~3; if (pl.CT < 4) return pl; pl.NT = ceil_div(pl.Nc, pl.CT); }
    pl.ldw = 4;
    pl.Cs = ceil_div(pl.CT, 4) * 4;
    if (pl.NT == 1 && (g.kw * cs4 > 256 || (ceil_div(g.kw * cs2, 16) < ceil_div(g.kw * cs4, 16) && tcw_env("MDT_TCW_LD2", 1)))) { pl.ldw = 2; pl.Cs = cs2; }
    pl.NW = ceil_div(g.kw * pl.Cs, 16) * 16;
    if (pl.NW > 256) return pl;
    // Where it wins (profiles/r02_tcw_vs_halo.txt): >= 112 stacked columns (36 -> 36/64, 64 -> 64/54 k3, 18 -> 18 k7: 1.07x .. 1.84x).  With the 64
    // stacked columns of the 18-channel 3x3x3 layers the halo-window kernel with 4-5 co-resident CTAs is faster (1.02 vs 1.16 ms).
    // MDT_TCW=2 forces this kernel wherever it is supported.
    if (tcw_env("MDT_TCW", 1) != 2 && !(pl.Kg <= 64 && pl.NW >= 112)) return pl;
    // accumulator layout.  "stacked": one MMA of N = 2*NW yields hi*hi and hi*lo side by side (fewer, wider MMAs: wins while 2*NW stays small);
    // prefer whatever leaves room for TWO accumulator sets (epilogue of tile i under the MMAs of tile i + 1)
    pl.TL = pl.RH > 1 ? 2 : 1;
    if (dgrad && g.sh > 1) { if (g.sh != 2 || pl.RH < 2) return pl; pl.TL = 2; }
    if (tcw_env("MDT_TCW_TL", 2) == 1 && !(dgrad && g.sh > 1)) pl.TL = 1;
    pl.stacked = (planes == 2 && pl.NW <= 64) ? 1 : 0;
    { const int v = tcw_env("MDT_TCW_STACK", -1); if (v == 0) pl.stacked = 0; if (v == 1 && planes == 2 && 2 * pl.NW <= 256) pl.stacked = 1; }
    pl.ACC = pl.stacked ? 2 * pl.NW : pl.NW;
    if (pl.TL * pl.ACC > 512) { if (dgrad && g.sh > 1) return pl; pl.TL = 1; }
    pl.nbuf = (2 * pl.TL * pl.ACC <= 512) ? 2 : 1;
    if (tcw_env("MDT_TCW_NBUF", 2) == 1) pl.nbuf = 1;
    int live = 0;
    if (!tcw_build_sched(g, dgrad, pl.TL, pl.sch, live)) return pl;
    pl.tmem_cols = 32;
    while (pl.tmem_cols < pl.nbuf * pl.TL * pl.ACC) pl.tmem_cols <<= 1;
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
    { const int v = tcw_env("MDT_TCW_SA", 0); if (v >= 2 && v <= kTcwMaxSA && bytes(v, SB) <= budget) SA = v; }
    { const int v = tcw_env("MDT_TCW_SB", 0); if (v >= live && v <= kTcwMaxSB && bytes(SA, v) <= budget) SB = v; }
    pl.SA = SA; pl.SB = SB;
    pl.smem_bytes = bytes(SA, SB) + 1024;
    pl.ctas_per_sm = 1;   // 320 threads x ~150 registers: one persistent CTA per SM
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
    p.Cn = pl.Nc; p.CT = pl.CT; p.Cs = pl.Cs; p.NW = pl.NW; p.stacked = pl.stacked; p.planes = planes; p.ldw = pl.ldw;
    p.TL = pl.TL; p.ACC = pl.ACC; p.nbuf = pl.nbuf; p.tmem_cols = pl.tmem_cols;
    p.nchunk = pl.nchunk;
    p.a_tx = 0; p.b_tx = 0; p.nops = 0;
    for (int c = 0; c < pl.nchunk; ++c) {
        p.ck0[c] = pl.ck0[c]; p.cw[c] = pl.cw[c]; p.cks[c] = pl.cks[c]; p.a_off[c] = pl.a_off[c]; p.b_off[c] = pl.b_off[c]; p.tm[c] = pl.tm[c];
        const int swz = 2 * pl.cw[c];
        p.a_tx += planes * 128 * swz;
        p.b_tx += planes * pl.NW * swz;
        // descriptor high word: SBO = 8 rows * swizzle span (16-byte units) | version 1 (bit 46) | layout type (bits 61..63)
        const unsigned hi = (unsigned)((8 * swz) >> 4) | (1u << 14) | (layout_type_for_swizzle_bytes(swz) << 29);
        for (int k16 = 0; k16 < pl.cks[c]; ++k16) {
            const int o = p.nops++;
            p.op_hi[o] = hi;
            p.op_a[o] = (unsigned short)((pl.a_off[c] + 32 * k16) >> 4);
            p.op_alo[o] = (unsigned short)((pl.a_off[c] + 128 * swz + 32 * k16) >> 4);
            p.op_b[o] = (unsigned short)((pl.b_off[c] + 32 * k16) >> 4);
            p.op_blo[o] = (unsigned short)((pl.b_off[c] + pl.NW * swz + 32 * k16) >> 4);
        }
    }
    p.a_stage_bytes = pl.a_stage; p.b_tile_bytes = pl.b_tile; p.SA = pl.SA; p.SB = pl.SB;
    p.inv_sa = (unsigned)(0x100000000ull / (unsigned)pl.SA) + 1u;
    p.inv_sb = (unsigned)(0x100000000ull / (unsigned)pl.SB) + 1u;
    p.tiles_h = ceil_div(pl.RH, pl.TL);
    p.total_tiles = (long long)g.n * pl.RD * p.tiles_h;
    p.relu = relu; p.bias = bias; p.residual = residual; p.out = dst;
    p.out_split = out_split; p.out_split_kg = conv_tc_kpad(pl.Nc);
    p.sch = pl.sch;
    p.prof = tcw_env("MDT_TCW_PROF", 0);
    p.skip = tcw_env("MDT_TCW_SKIP", 0);

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
    if (!ensure_smem_attr(conv_tcw_kernel, 214 * 1024, attr)) return MDT_EDRIVER;   // + 12 KB static (edge buffers, bias, barriers) <= 227 KB
    long long gx = (long long)num_sms() * pl.ctas_per_sm;
    if (gx > p.total_tiles) gx = p.total_tiles;
    dim3 grid((unsigned)gx, (unsigned)pl.NT);
    conv_tcw_kernel<<<grid, kTcwThreads, (size_t)pl.smem_bytes, st>>>(maps, p);
    return launch_status();
}

}  // namespace mdt
