// CTA-pair (tcgen05 cta_group::2) variant of the implicit-GEMM conv kernel of conv3d_tc.cu — EXPERIMENTAL, off unless MDT_TC_PAIR=1.
//
// Status: written against the PTX ISA and the measurements of round 1; compiles for sm_100a, NOT yet run on hardware (the assumptions it
// rests on are exactly what tools/mma2_probe.cu checks: M = 256 MMAs over two CTAs' shared memory, B split by rows between the CTAs at
// equal shared-memory offsets, TMA completions counted on the leader's mbarrier, multicast commits).  The default path never reaches it.
//
// Why: ncu of conv_tc_kernel (36->36 k3 at 2x128^3) shows the tensor pipe 44 % active with DRAM at 9 %.  One thread issues an MMA every
// >= 60 cycles and each MMA of a 128-row tile re-reads its B tile from shared memory ((M + N) K operand bytes for M N K MACs: the small
// N of these layers makes the kernel shared-memory- and issue-bound).  A CTA pair halves both costs per tile: one instruction covers
// M = 256 rows (two tiles: two adjacent output lines), and each CTA supplies only HALF of B.
//
// Mapping of the split-bf16 N-stacking onto the pair (rank r = %cluster_ctarank, NT = columns per tile):
//   MMA 1   D[256 x 2NT] += A_hi[256 x K] * [B_hi ; B_lo]^T      rank r stages plane r of the weight tile (NT rows) at offset X
//   MMA 2   D[256 x NT]  += A_lo[256 x K] * B_hi^T               rank r stages rows [r NT/2, (r+1) NT/2) of B_hi at offset Y
// so each CTA loads 1.5 NT weight rows per tap instead of 2 NT, and A (own halo line, both planes) exactly as before.
// Barriers: full[s] lives in the leader and counts the bytes of BOTH CTAs (peer loads use the .cta_group::2 TMA form with the leader's
// barrier address); empty[s] and accum_full exist in both CTAs and are arrived on by multicast tcgen05.commit from the leader.
// The two CTAs must walk identical stage sequences, so taps are skipped only when the source PLANE (d) is out of range — identical for
// both lines of a pair; out-of-range source LINES (h) are fetched as TMA out-of-bounds zeros instead.
#include "conv3d_tc_plan.cuh"

namespace mdt {
using namespace tc;

namespace {

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols) {   // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this shared-memory offset in both CTAs once every MMA issued so far has completed
__device__ __forceinline__ void umma2_commit_both(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// TMA loads into this CTA's shared memory whose completion bytes are counted on a barrier given by its shared::cluster address
__device__ __forceinline__ void tma_load_4d_pair(void *dst, const CUtensorMap *m, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void *dst, const CUtensorMap *m, uint32_t bar_cluster, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// source plane / line feeding row-line (rd, rh0) through tap (kd, kh).  false = the PLANE is outside the source (same answer for both CTAs
// of a pair: they share rd); an out-of-range line is returned as is and becomes TMA zero fill.
__device__ __forceinline__ bool pair_step_coords(const TcConvParams &p, int rd, int rh0, int kd, int kh, int &d_src, int &h_src) {
    if (!p.dgrad) {
        d_src = rd * p.sd - p.pd + kd;
        h_src = rh0 * p.sh - p.ph + kh;
    } else {
        const int td = rd + p.pd - kd;
        if (td < 0 || td % p.sd != 0) return false;
        d_src = td / p.sd;
        h_src = rh0 + p.ph - kh;   // sh == 1 on this path
    }
    return d_src >= 0 && d_src < p.SD;
}

// Halo mode, one stage, split-bf16 only (see conv_tc_pair_wanted).  Warp roles as in conv_tc_kernel: 0 TMA producer, 1 MMA issuer (leader
// CTA only) + TMEM allocation, 2-5 epilogue.
__global__ void __launch_bounds__(kTcThreads, 4)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBx, const __grid_constant__ CUtensorMap tmBy,
                    const TcConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full, empty, accum_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t s_have_acc, s_n1;
    __shared__ float s_bias[128];

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    int t = blockIdx.x;   // consecutive blocks = the two CTAs of a pair: neighbouring tiles of the same (nb, rd) plane
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    const int rd = t % p.RD;
    const int nb = t / p.RD;
    const int rw0 = tw * p.BW, rh0 = th;   // BH == 1
    const int n0 = blockIdx.y * p.NT;
    const int chunk_elems = p.swz >> 1;
    const int ngroups = (p.nchunks + p.CPS - 1) / p.CPS;
    const int y_off = p.TPS * p.b_plane_bytes;      // inside a chunk's B area: [X: TPS x NT rows][Y: TPS x NT/2 rows]
    const int y_tap_bytes = p.b_plane_bytes / 2;

    if (threadIdx.x == 0) {
        mbar_init(&full, 1);
        mbar_init(&empty, 1);
        mbar_init(&accum_full, 1);
        s_have_acc = 1;
        fence_barrier_init();
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmBx);
        prefetch_tmap(&tmBy);
    }
    if (threadIdx.x >= 64) {
        const int c = threadIdx.x - 64;
        s_bias[c] = (p.bias && n0 + c < p.Cn) ? __ldg(p.bias + n0 + c) : 0.f;
    }
    const int acc1_cols = 2 * p.NT;
    const int Q = p.Q;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < Q * acc1_cols) tmem_cols <<= 1;
    if (warp == 1) tmem_alloc2(&tmem_base_s, tmem_cols);
    tc_fence_before();
    cluster_sync_all();   // barriers initialised and TMEM allocated in BOTH CTAs before anyone signals or issues
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // =============================================================== TMA producer (both CTAs)
        if (lane == 0) {
            const uint32_t leader_full = map_to_cta(smem_u32(&full), 0);
            uint32_t ph = 1;   // parity to wait for on `empty`
            for (int kd = 0; kd < p.KD; ++kd)
                for (int kh = 0; kh < p.KH; ++kh) {
                    int d_src, h_src;
                    if (!pair_step_coords(p, rd, rh0, kd, kh, d_src, h_src)) continue;
                    for (int g = 0; g < ngroups; ++g) {
                        const int c_lo = g * p.CPS, cn = min(p.CPS, p.nchunks - c_lo);
                        for (int kw0 = 0; kw0 < p.KW; kw0 += p.TPS) {
                            mbar_wait(&empty, ph);
                            ph ^= 1;
                            const bool load_a = kw0 == 0;   // the halo stays in place while the later tap groups stream their weights
                            if (rank == 0)   // bytes of both CTAs: A (2 planes) + X (NT rows per tap) + Y (NT/2 rows per tap), per chunk
                                mbar_arrive_expect_tx(&full, 2u * (uint32_t)(cn * ((load_a ? 2 * p.a_tx_bytes : 0) + p.TPS * (p.b_plane_bytes + y_tap_bytes))));
                            const int w_start = p.dgrad ? rw0 + p.pw - (p.KW - 1) : rw0 - p.pw;
                            const int tap0 = (kd * p.KH + kh) * p.KW + kw0;
                            for (int c = 0; c < cn; ++c) {
                                uint8_t *ab = smem + (size_t)c * p.a_chunk_bytes;
                                uint8_t *bb = smem + p.a_region_bytes + (size_t)c * p.b_chunk_bytes;
                                if (load_a) tma_load_5d_pair(ab, &tmA, leader_full, (c_lo + c) * chunk_elems, w_start, 0, h_src, nb * p.SD + d_src);
                                tma_load_4d_pair(bb, &tmBx, leader_full, (c_lo + c) * chunk_elems, n0, (int)rank, tap0);
                                tma_load_4d_pair(bb + y_off, &tmBy, leader_full, (c_lo + c) * chunk_elems, n0 + (int)rank * (p.NT / 2), 0, tap0);
                            }
                        }
                    }
                }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer: the leader issues, the peer only replays the counters
        if (lane == 0) {
            const uint32_t idesc1 = make_idesc_bf16(256, p.NT, 0, 0);
            const uint32_t idesc2 = make_idesc_bf16(256, 2 * p.NT, 0, 0);
            const uint64_t dtmpl = make_smem_desc(0, 16, 8u * p.swz, layout_type_for_swizzle_bytes(p.swz));
            const int ksteps = p.swz / 32;
            const uint32_t smem_base = smem_u32(smem);
            uint32_t ph = 0, n1 = 0, q1 = 0;
            for (int kd = 0; kd < p.KD; ++kd)
                for (int kh = 0; kh < p.KH; ++kh) {
                    int d_src, h_src;
                    if (!pair_step_coords(p, rd, rh0, kd, kh, d_src, h_src)) continue;
                    for (int g = 0; g < ngroups; ++g) {
                        const int cn = min(p.CPS, p.nchunks - g * p.CPS);
                        for (int kw0 = 0; kw0 < p.KW; kw0 += p.TPS) {
                            const int ktaps = min(p.TPS, p.KW - kw0);
                            if (rank != 0) { n1 += (uint32_t)(ktaps * cn * ksteps); continue; }
                            mbar_wait(&full, ph);
                            ph ^= 1;
                            tc_fence_after();
                            for (int k = 0; k < ktaps; ++k) {
                                const int shift = p.dgrad ? p.KW - 1 - (kw0 + k) : kw0 + k;
                                for (int c = 0; c < cn; ++c) {
                                    const uint32_t a_hi = smem_base + c * p.a_chunk_bytes + shift * p.swz;
                                    const uint32_t b_x = smem_base + p.a_region_bytes + c * p.b_chunk_bytes + k * p.b_plane_bytes;
                                    const uint32_t b_y = smem_base + p.a_region_bytes + c * p.b_chunk_bytes + y_off + k * y_tap_bytes;
                                    uint64_t da = dtmpl | (uint64_t)((a_hi >> 4) & 0x3FFF);
                                    uint64_t dal = dtmpl | (uint64_t)(((a_hi + p.a_plane_bytes) >> 4) & 0x3FFF);
                                    uint64_t dbx = dtmpl | (uint64_t)((b_x >> 4) & 0x3FFF);
                                    uint64_t dby = dtmpl | (uint64_t)((b_y >> 4) & 0x3FFF);
                                    for (int j = 0; j < ksteps; ++j) {
                                        umma2_bf16(tmem + q1 * acc1_cols, da, dbx, idesc2, n1 >= (uint32_t)Q);   // [hi*hi | hi*lo]
                                        umma2_bf16(tmem + q1 * acc1_cols, dal, dby, idesc1, 1);                  // hi*hi columns += lo*hi
                                        ++n1;
                                        if (++q1 == (uint32_t)Q) q1 = 0;
                                        da += 2; dal += 2; dbx += 2; dby += 2;   // +32 bytes along K
                                    }
                                }
                            }
                            umma2_commit_both(&empty);
                        }
                    }
                }
            *(volatile uint32_t *)&s_n1 = n1 < (uint32_t)Q ? n1 : (uint32_t)Q;
            __threadfence_block();
            if (n1) {
                if (rank == 0) umma2_commit_both(&accum_full);
            } else {   // no tap contributed (same for both CTAs): bias / zero only
                *(volatile uint32_t *)&s_have_acc = 0;
                mbar_arrive(&accum_full);
            }
        }
    }
    if (warp >= 2) {
        // =============================================================== epilogue (each CTA: its own 128 TMEM lanes = its own output line)
        mbar_wait(&accum_full, 0);
        tc_fence_after();
        const bool have_acc = (*(volatile uint32_t *)&s_have_acc) != 0u;
        const uint32_t n1_used = *(volatile uint32_t *)&s_n1;
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int rh = rh0, rw = rw0 + r;
        const bool valid = rh < p.RH && rw < p.RW;
        const size_t row_off = ((((size_t)nb * p.RD + rd) * p.RH + rh) * p.RW + rw) * (size_t)p.Cn;
        const bool vec4 = (p.Cn % 4 == 0);
        for (int c0 = 0; c0 < p.NT; c0 += 16) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
            if (have_acc) {
                const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
                float u[16];
                for (uint32_t a = 0; a < n1_used; ++a) {
                    tmem_ld16(lane_base + a * acc1_cols + c0, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += u[j];
                    tmem_ld16(lane_base + a * acc1_cols + p.NT + c0, u);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += u[j];
                }
            }
            if (!valid) continue;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const int col = n0 + c0 + j;
                if (col >= p.Cn) break;
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = v[j + k] + s_bias[c0 + j + k];
                if (vec4) {
                    if (p.residual) {
                        const float4 rr = __ldg(reinterpret_cast<const float4 *>(p.residual + row_off + col));
                        o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
                    }
                    if (p.relu) { o[0] = fmaxf(o[0], 0.f); o[1] = fmaxf(o[1], 0.f); o[2] = fmaxf(o[2], 0.f); o[3] = fmaxf(o[3], 0.f); }
                    *reinterpret_cast<float4 *>(p.out + row_off + col) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (col + k >= p.Cn) break;
                        float x = o[k];
                        if (p.residual) x += __ldg(p.residual + row_off + col + k);
                        if (p.relu) x = fmaxf(x, 0.f);
                        p.out[row_off + col + k] = x;
                    }
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();   // the peer's shared memory and TMEM stay alive until both CTAs are done
    if (warp == 1) tmem_dealloc2(tmem, tmem_cols);
}

}  // namespace

bool conv_tc_pair_wanted(const ConvGeom &g, const TcPlan &pl, const TcConvParams &p, int pass) {
    const char *e = getenv("MDT_TC_PAIR");
    if (!e || e[0] != '1') return false;
    if (!pl.halo || p.planes != 2 || p.D != 1) return false;            // halo lines, split-bf16, single stage
    if (p.NT % 16 || p.NT > 128) return false;                          // N = 2 NT <= 256, NT / 2 rows = whole 8-row swizzle atoms
    if ((p.tiles_w * p.tiles_h) % 2) return false;                      // a pair must not straddle two (nb, rd) planes
    if (pass == 1 && g.sh != 1) return false;                           // strided dgrad lines alternate between contributing and not
    return true;
}

int conv_tc_pair_launch(const ConvGeom &g, const TcPlan &pl, const TcConvParams &p, const CUtensorMap &tmA, const void *packed_weights, int T,
                        cudaStream_t st) {
    // weight maps with per-CTA boxes: X = one plane x NT rows x TPS taps, Y = plane 0 x NT/2 rows x TPS taps
    CUtensorMap tmBx, tmBy;
    const uint64_t bdims[4] = {(uint64_t)pl.Kp, (uint64_t)pl.Np, 2u, (uint64_t)T};
    const uint64_t bstr[3] = {(uint64_t)pl.Kp * 2, (uint64_t)pl.Np * pl.Kp * 2, 2ull * pl.Np * pl.Kp * 2};
    const uint32_t box_x[4] = {(uint32_t)(pl.swz / 2), (uint32_t)pl.NT, 1u, (uint32_t)p.TPS};
    const uint32_t box_y[4] = {(uint32_t)(pl.swz / 2), (uint32_t)(pl.NT / 2), 1u, (uint32_t)p.TPS};
    void *wp = const_cast<void *>(packed_weights);
    if (!encode_bf16_tmap(&tmBx, wp, 4, bdims, bstr, box_x, pl.swz) || !encode_bf16_tmap(&tmBy, wp, 4, bdims, bstr, box_y, pl.swz)) return MDT_EDRIVER;
    const size_t smem = (size_t)p.stage_bytes + 1024;
    static bool attr[kMaxDevices] = {};
    if (!ensure_smem_attr(conv_tc_pair_kernel, 220 * 1024, attr)) return MDT_EDRIVER;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((long long)g.n * pl.RD * p.tiles_h * p.tiles_w), pl.n_tiles_n);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_pair_kernel, tmA, tmBx, tmBy, p);
    if (e != cudaSuccess) return (int)e;
    return launch_status();
}

}  // namespace mdt
