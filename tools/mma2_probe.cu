// mma2_probe — hardware probe for the CTA-pair (cta_group::2) assumptions the next conv kernel wants to rely on.  NOT part of the
// library; build + run on a B200:   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/mma2_probe tools/mma2_probe.cu && /tmp/mma2_probe
//
// A cluster of two CTAs computes D[256 x N] = A[256 x K] * B[N x K]^T (bf16 in, fp32 out):
//   * CTA r holds A rows [128 r, 128 r + 128) and B rows [r N/2, (r+1) N/2) at the SAME shared-memory offsets (K-major, SWIZZLE_128B, TMA);
//   * the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256; each CTA's TMEM receives its own 128 rows x N columns;
//   * tcgen05.commit ... multicast::cluster tells both CTAs that the accumulator is complete.
// Questions answered (printed as PASS/FAIL + numbers):
//   sync=0  operands announced to the leader by a cluster barrier            -> are the M = 256 / split-B semantics as assumed?
//   sync=1  the peer's TMA loads complete_tx on the LEADER's mbarrier         -> can the producer/consumer ring stay barrier-only?
//   rate    cycles per cta_group::2 MMA issued back to back by one thread for the conv shapes (N = 96 / 48 / 128 / 256, K = 16),
//           to compare with profiles/r01_mma_rate.txt (cta_group::1: >= 60 cycles per MMA)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../medicaldetectiontoolkit_b200/csrc/tc_common.cuh"

using namespace mdt;
using namespace mdt::tc;

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
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols) {
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
// arrive on the mbarrier at the same shared-memory offset in every CTA of `mask` once all MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// TMA load into this CTA's shared memory whose completion is counted on an mbarrier given as a shared::cluster address (may be the peer's)
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
                 : "memory");
}

struct Probe2 {
    int N, K;     // N multiple of 32 (N/2 rows per CTA, 8-row swizzle atoms), K multiple of 64
    int sync;     // 0 cluster barrier, 1 cross-CTA complete_tx
    int reps;     // > 0: rate mode, issue `reps` accumulating MMAs of K = 16 and report cycles
};

constexpr int kSw = 128, kChunk = 64;   // SWIZZLE_128B: 64 bf16 per row

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                                const __grid_constant__ CUtensorMap tmB, Probe2 p, float *out,
                                                                                long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_ctarank();
    const int nchunks = p.K / kChunk;
    const uint32_t a_chunk_bytes = 128 * kSw, b_chunk_bytes = (p.N / 2) * kSw;
    uint8_t *sA = smem;
    uint8_t *sB = smem + nchunks * a_chunk_bytes;   // multiples of 16 KB: stays 1024-byte aligned
    if (threadIdx.x == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc2(&tmem_base, 256);
    tc_fence_before();
    cluster_sync_all();   // barriers initialised and TMEM allocated in both CTAs
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    const uint32_t my_bytes = nchunks * (a_chunk_bytes + b_chunk_bytes);
    if (threadIdx.x == 0) {
        if (p.sync == 0) {
            mbar_arrive_expect_tx(&bar_full, my_bytes);
            for (int c = 0; c < nchunks; ++c) tma_load_2d(sA + c * a_chunk_bytes, &tmA, &bar_full, c * kChunk, (int)rank * 128);
            for (int c = 0; c < nchunks; ++c) tma_load_2d(sB + c * b_chunk_bytes, &tmB, &bar_full, c * kChunk, (int)rank * (p.N / 2));
            mbar_wait(&bar_full, 0);
        } else {
            const uint32_t leader_bar = map_to_cta(smem_u32(&bar_full), 0);
            if (rank == 0) mbar_arrive_expect_tx(&bar_full, 2 * my_bytes);   // both CTAs' bytes are counted on the leader's barrier
            for (int c = 0; c < nchunks; ++c) tma_load_2d_pair(sA + c * a_chunk_bytes, &tmA, leader_bar, c * kChunk, (int)rank * 128);
            for (int c = 0; c < nchunks; ++c) tma_load_2d_pair(sB + c * b_chunk_bytes, &tmB, leader_bar, c * kChunk, (int)rank * (p.N / 2));
            if (rank == 0) mbar_wait(&bar_full, 0);
        }
    }
    if (p.sync == 0) cluster_sync_all();   // the leader may read the peer's operands only after the peer has seen its loads land
    if (threadIdx.x == 0 && rank == 0) {
        tc_fence_after();
        const uint32_t lt = layout_type_for_swizzle_bytes(kSw);
        const uint32_t idesc = make_idesc_bf16(256, p.N, 0, 0);
        const uint32_t sbo = 8 * kSw;
        if (p.reps == 0) {
            int acc = 0;
            for (int c = 0; c < nchunks; ++c)
                for (int k = 0; k < kChunk / 16; ++k) {
                    const uint32_t a_addr = smem_u32(sA + c * a_chunk_bytes) + k * 32, b_addr = smem_u32(sB + c * b_chunk_bytes) + k * 32;
                    umma2_bf16(tmem, make_smem_desc(a_addr, 16, sbo, lt, 0), make_smem_desc(b_addr, 16, sbo, lt, 0), idesc, acc);
                    acc = 1;
                }
        } else {
            const uint64_t da = make_smem_desc(smem_u32(sA), 16, sbo, lt, 0), db = make_smem_desc(smem_u32(sB), 16, sbo, lt, 0);
            const long long t0 = clock64();
            for (int i = 0; i < p.reps; ++i) umma2_bf16(tmem, da, db, idesc, i > 0);
            const long long t1 = clock64();
            cycles[0] = t1 - t0;   // issue time; completion time is measured by the host around the whole launch
        }
        umma2_commit(&bar_done, 0b11);
    }
    __syncthreads();
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    if (p.reps == 0) {
        for (int c0 = 0; c0 < p.N; c0 += 8) {
            float v[8];
            tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 8; ++j) out[(size_t)(rank * 128 + warp * 32 + lane) * p.N + c0 + j] = v[j];
        }
    } else if (threadIdx.x == 0 && rank == 0) {
        cycles[1] = clock64();
    }
    tc_fence_before();
    cluster_sync_all();   // nobody deallocates while the peer still reads TMEM / the leader's MMAs still read the peer's shared memory
    if (warp == 0) tmem_dealloc2(tmem, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static int run(const char *name, Probe2 p) {
    const int M = 256;
    std::vector<float> A((size_t)M * p.K), B((size_t)p.N * p.K);
    srand(99 + p.N + p.K + p.sync);
    for (auto &v : A) v = bf((rand() % 2001 - 1000) / 1000.f);
    for (auto &v : B) v = bf((rand() % 2001 - 1000) / 1000.f);
    std::vector<__nv_bfloat16> hA(A.size()), hB(B.size());
    for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2bfloat16(A[i]);
    for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2bfloat16(B[i]);
    __nv_bfloat16 *dA, *dB; float *dOut; long long *dCyc;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, (size_t)M * p.N * 4); cudaMalloc(&dCyc, 16);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dOut, 0xff, (size_t)M * p.N * 4);
    cudaMemset(dCyc, 0, 16);
    CUtensorMap tmA, tmB;
    uint64_t dimsA[2] = {(uint64_t)p.K, (uint64_t)M}, strA[1] = {(uint64_t)p.K * 2};
    uint32_t boxA[2] = {(uint32_t)kChunk, 128};
    uint64_t dimsB[2] = {(uint64_t)p.K, (uint64_t)p.N}, strB[1] = {(uint64_t)p.K * 2};
    uint32_t boxB[2] = {(uint32_t)kChunk, (uint32_t)(p.N / 2)};
    if (!(encode_bf16_tmap(&tmA, dA, 2, dimsA, strA, boxA, kSw) && encode_bf16_tmap(&tmB, dB, 2, dimsB, strB, boxB, kSw))) {
        printf("%-28s ENCODE_FAILED\n", name);
        return 1;
    }
    const size_t smem = 160 * 1024;
    cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe2_kernel<<<2, 128, smem>>>(tmA, tmB, p, dOut, dCyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s CUDA_ERROR %s\n", name, cudaGetErrorString(e)); return 2; }
    int rc = 0;
    if (p.reps == 0) {
        std::vector<float> out((size_t)M * p.N);
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < p.N; ++n) {
                double r = 0;
                for (int k = 0; k < p.K; ++k) r += (double)A[(size_t)m * p.K + k] * B[(size_t)n * p.K + k];
                maxerr = fmax(maxerr, fabs(r - out[(size_t)m * p.N + n]));
                maxref = fmax(maxref, fabs(r));
            }
        printf("%-28s max_abs_err %.3e  max_ref %.3f  %s\n", name, maxerr, maxref, maxerr < 1e-3 * maxref ? "PASS" : "FAIL");
        rc = maxerr < 1e-3 * maxref ? 0 : 3;
    } else {
        long long cyc[2];
        cudaMemcpy(cyc, dCyc, 16, cudaMemcpyDeviceToHost);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %d MMAs (M=256, N=%d, K=16): issue %.1f cycles/MMA, kernel %.3f ms (incl. launch) = %.1f TFLOP/s\n", name, p.reps, p.N,
               (double)cyc[0] / p.reps, ms, 2.0 * 256 * p.N * 16 * p.reps / (ms * 1e9));
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dOut); cudaFree(dCyc);
    return rc;
}

int main() {
    char name[64];
    for (int sync : {0, 1})
        for (int N : {64, 96, 128, 256}) {
            snprintf(name, sizeof name, "pair_sync%d_N%d_K128", sync, N);
            run(name, Probe2{N, 128, sync, 0});
        }
    for (int N : {32, 64, 96, 128, 256}) {
        snprintf(name, sizeof name, "pair_rate_N%d", N);
        run(name, Probe2{N, 64, 0, 20000});
    }
    return 0;
}
