// tc_probe — hardware probe for the tcgen05 / TMA assumptions the conv kernels rely on (run once on the B200; results go to profiles/).
// One CTA computes D[128 x N] = A[128 x K] * B[N x K]^T (bf16 in, fp32 out) from TMA-loaded, hardware-swizzled tiles and checks it
// against a host reference, for:
//   kmajor  sw in {128, 64, 32}          K-major operands, K split into swizzle-span chunks (what fprop/dgrad use)
//   shift   row offset s in {1, 2, 5}    A descriptor starts s rows into a taller tile (halo reuse), base_offset 0 vs (s & 7)
//   mnmajor sw in {128, 64, 32}          both operands MN-major: A^T [K x 128], B^T [K x N] tiles (what wgrad uses)
//   mnshift sw in {128, 64}              MN-major B whose N index runs over 3 row-shifted windows of one halo tile (LBO = one row):
//                                        D[m][s*chunk + c] = sum_k A[m][k] * Bt[k + s][c]   (wgrad: all kw taps in one MMA)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../medicaldetectiontoolkit_b200/csrc/tc_common.cuh"

using namespace mdt;
using namespace mdt::tc;

struct ProbeParams {
    int N, K;            // K multiple of 16
    int sw;              // swizzle bytes: chunk = sw/2 elements
    int mode;            // 0 kmajor, 1 shift, 2 mnmajor
    int shift, base_off; // mode 1
    int a_rows;          // rows loaded for A (128, or 128+8 for shift)
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ProbeParams p,
                                                    float *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int chunk = p.sw / 2;                 // elements per swizzle span
    const int nchunks_k = p.K / chunk;          // K-major: chunks along K
    // K-major: A chunk buffers [a_rows][chunk] each a_rows*sw bytes; B chunk buffers [N][chunk]
    // MN-major: A buffers: for each 'chunk'-wide slice of M (128/chunk slices): [K rows][chunk]; B: N/chunk slices of [K][chunk]
    const uint32_t a_chunk_bytes = (p.mode >= 2 ? p.K : p.a_rows) * p.sw;
    const uint32_t b_chunk_bytes = (p.mode == 3 ? p.K + 8 : p.mode == 2 ? p.K : p.N) * p.sw;
    const int a_nbuf = p.mode >= 2 ? 128 / chunk : nchunks_k;
    const int b_nbuf = p.mode == 3 ? 1 : p.mode == 2 ? p.N / chunk : nchunks_k;
    uint8_t *sA = smem;
    uint8_t *sB = smem + ((a_nbuf * a_chunk_bytes + 1023) / 1024) * 1024;
    if (threadIdx.x == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_full, a_nbuf * a_chunk_bytes + b_nbuf * b_chunk_bytes);
        for (int c = 0; c < a_nbuf; ++c) tma_load_2d(sA + c * a_chunk_bytes, &tmA, &bar_full, c * chunk, 0);
        for (int c = 0; c < b_nbuf; ++c) tma_load_2d(sB + c * b_chunk_bytes, &tmB, &bar_full, c * chunk, 0);
        mbar_wait(&bar_full, 0);
        tc_fence_after();
        const uint32_t lt = layout_type_for_swizzle_bytes(p.sw);
        if (p.mode < 2) {
            const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
            const uint32_t sbo = 8 * p.sw;
            int acc = 0;
            for (int c = 0; c < nchunks_k; ++c)
                for (int k = 0; k < chunk / 16; ++k) {
                    const uint32_t a_addr = smem_u32(sA + c * a_chunk_bytes) + p.shift * p.sw + k * 32;
                    const uint32_t b_addr = smem_u32(sB + c * b_chunk_bytes) + k * 32;
                    umma_bf16(tmem, make_smem_desc(a_addr, 16, sbo, lt, p.base_off), make_smem_desc(b_addr, 16, sbo, lt, 0), idesc, acc);
                    acc = 1;
                }
        } else {
            // MN-major: per K16 step the descriptor covers [16 k rows][M]; k rows pitch = sw bytes, 8-row groups SBO = 8*sw,
            // MN slices of `chunk` elements at LBO = a_chunk_bytes
            const uint32_t idesc = make_idesc_bf16(128, p.N, 1, 1);
            int acc = 0;
            for (int k = 0; k < p.K / 16; ++k) {
                const uint32_t a_addr = smem_u32(sA) + k * 16 * p.sw;
                const uint32_t b_addr = smem_u32(sB) + k * 16 * p.sw;
                const uint32_t b_lbo = p.mode == 3 ? (uint32_t)p.sw : b_chunk_bytes;   // mode 3: next N chunk = next ROW of the halo tile
                umma_bf16(tmem, make_smem_desc(a_addr, a_chunk_bytes, 8 * p.sw, lt, 0), make_smem_desc(b_addr, b_lbo, 8 * p.sw, lt, 0), idesc, acc);
                acc = 1;
            }
        }
        umma_commit(&bar_done);
    }
    __syncthreads();
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < p.N; c0 += 8) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out[(size_t)(warp * 32 + lane) * p.N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static int run(const char *name, ProbeParams p) {
    const int M = 128, rowsA = p.a_rows;
    const int chunk_h = p.sw / 2;
    std::vector<float> A((size_t)rowsA * p.K), B((size_t)(p.mode == 3 ? chunk_h * (p.K + 8) : p.N * p.K));
    srand(1234 + p.sw + p.mode * 7 + p.shift);
    for (auto &v : A) v = bf((rand() % 2001 - 1000) / 1000.f);
    for (auto &v : B) v = bf((rand() % 2001 - 1000) / 1000.f);
    // device layouts: K-major: A[rowsA][K], B[N][K];  MN-major: At[K][M], Bt[K][N]
    std::vector<__nv_bfloat16> hA, hB;
    if (p.mode < 2) {
        hA.resize(A.size()); hB.resize(B.size());
        for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2bfloat16(A[i]);
        for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2bfloat16(B[i]);
    } else {
        hA.resize((size_t)p.K * M);
        for (int m = 0; m < M; ++m) for (int k = 0; k < p.K; ++k) hA[(size_t)k * M + m] = __float2bfloat16(A[(size_t)m * p.K + k]);
        if (p.mode == 2) {
            hB.resize((size_t)p.K * p.N);
            for (int n = 0; n < p.N; ++n) for (int k = 0; k < p.K; ++k) hB[(size_t)k * p.N + n] = __float2bfloat16(B[(size_t)n * p.K + k]);
        } else {   // mode 3: Bt halo [K + 8][chunk], stored as is
            hB.resize(B.size());
            for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2bfloat16(B[i]);
        }
    }
    __nv_bfloat16 *dA, *dB; float *dOut;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, (size_t)M * p.N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dOut, 0xff, (size_t)M * p.N * 4);
    CUtensorMap tmA, tmB;
    const int chunk = p.sw / 2;
    bool ok;
    if (p.mode < 2) {
        uint64_t dimsA[2] = {(uint64_t)p.K, (uint64_t)rowsA}, strA[1] = {(uint64_t)p.K * 2};
        uint32_t boxA[2] = {(uint32_t)chunk, (uint32_t)rowsA};
        uint64_t dimsB[2] = {(uint64_t)p.K, (uint64_t)p.N}, strB[1] = {(uint64_t)p.K * 2};
        uint32_t boxB[2] = {(uint32_t)chunk, (uint32_t)p.N};
        ok = encode_bf16_tmap(&tmA, dA, 2, dimsA, strA, boxA, p.sw) && encode_bf16_tmap(&tmB, dB, 2, dimsB, strB, boxB, p.sw);
    } else {
        uint64_t dimsA[2] = {(uint64_t)M, (uint64_t)p.K}, strA[1] = {(uint64_t)M * 2};
        uint32_t boxA[2] = {(uint32_t)chunk, (uint32_t)p.K};
        uint64_t dimsB[2] = {(uint64_t)(p.mode == 3 ? chunk : p.N), (uint64_t)(p.mode == 3 ? p.K + 8 : p.K)}, strB[1] = {(uint64_t)(p.mode == 3 ? chunk : p.N) * 2};
        uint32_t boxB[2] = {(uint32_t)chunk, (uint32_t)(p.mode == 3 ? p.K + 8 : p.K)};
        ok = encode_bf16_tmap(&tmA, dA, 2, dimsA, strA, boxA, p.sw) && encode_bf16_tmap(&tmB, dB, 2, dimsB, strB, boxB, p.sw);
    }
    if (!ok) { printf("%-28s ENCODE_FAILED\n", name); return 1; }
    const size_t smem = 160 * 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(tmA, tmB, p, dOut);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s CUDA_ERROR %s\n", name, cudaGetErrorString(e)); return 2; }
    std::vector<float> out((size_t)M * p.N);
    cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < p.N; ++n) {
            double r = 0;
            if (p.mode == 3) { const int sft = n / chunk, c = n % chunk; for (int k = 0; k < p.K; ++k) r += (double)A[(size_t)m * p.K + k] * B[(size_t)(k + sft) * chunk + c]; }
            else for (int k = 0; k < p.K; ++k) r += (double)A[(size_t)(m + p.shift) * p.K + k] * B[(size_t)n * p.K + k];
            maxerr = fmax(maxerr, fabs(r - out[(size_t)m * p.N + n]));
            maxref = fmax(maxref, fabs(r));
        }
    printf("%-28s max_abs_err %.3e  max_ref %.3f  %s\n", name, maxerr, maxref, maxerr < 1e-3 * maxref ? "PASS" : "FAIL");
    cudaFree(dA); cudaFree(dB); cudaFree(dOut);
    return maxerr < 1e-3 * maxref ? 0 : 3;
}

int main() {
    int bad = 0;
    char name[64];
    for (int sw : {128, 64, 32}) {
        snprintf(name, sizeof name, "kmajor_sw%d_N64_K64", sw);
        bad += run(name, ProbeParams{64, 64, sw, 0, 0, 0, 128}) != 0;
        snprintf(name, sizeof name, "kmajor_sw%d_N48_K128", sw);
        bad += run(name, ProbeParams{48, 128, sw, 0, 0, 0, 128}) != 0;
    }
    for (int sw : {128, 32})
        for (int s : {1, 2, 5, 8})
            for (int bo : {0, 1}) {
                                const int bov = bo ? ((s * sw) >> 7) & 7 : 0;
                if (bo && bov == 0) continue;
                snprintf(name, sizeof name, "shift_sw%d_s%d_bo%d", sw, s, bov);
                run(name, ProbeParams{64, 64, sw, 1, s, bov, 136});  // informational: decides how halo reuse is addressed
            }
    for (int sw : {128, 64, 32}) {
        snprintf(name, sizeof name, "mnmajor_sw%d_N64_K64", sw);
        run(name, ProbeParams{64, 64, sw, 2, 0, 0, 128});
        snprintf(name, sizeof name, "mnmajor_sw%d_N128_K32", sw);
        run(name, ProbeParams{128, 32, sw, 2, 0, 0, 128});
    }
    for (int sw : {128, 64}) {
        snprintf(name, sizeof name, "mnshift_sw%d_N%d_K64", sw, 3 * sw / 2);
        run(name, ProbeParams{3 * sw / 2, 64, sw, 3, 0, 0, 128});
        snprintf(name, sizeof name, "mnshift_sw%d_N%d_K128", sw, 3 * sw / 2);
        run(name, ProbeParams{3 * sw / 2, 128, sw, 3, 0, 0, 128});
    }
    printf("kmajor failures: %d\n", bad);
    return 0;
}
