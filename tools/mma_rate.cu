// mma_rate — how fast can ONE thread drive tcgen05.mma for small N?  Issues `iters` MMAs (M=128, K=16, bf16, SW128 K-major operands
// already resident in shared memory) from one elected thread and reports cycles per MMA for N in {32,64,128,256}, with 1/2/4
// independent accumulators (round robin) and 1..4 co-resident CTAs per SM.  Decides the conv kernel design (profiles/r01_mma_rate.txt).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../medicaldetectiontoolkit_b200/csrc/tc_common.cuh"
using namespace mdt;
using namespace mdt::tc;

__global__ void __launch_bounds__(128) rate_kernel(int N, int nacc, int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    uint32_t cols = 32;
    while ((int)cols < N * nacc) cols <<= 1;
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024, 2), db = make_smem_desc(smem_u32(smem + 16384), 16, 1024, 2);
        const long long t0 = clock64();
        int q = 0;
        for (int i = 0; i < iters; ++i) {
            umma_bf16(tmem + q * N, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1);
            if (++q == nacc) q = 0;
        }
        const long long t_issue = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t_issue - t0; out[1] = t1 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, cols);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    printf("%5s %5s %8s | %10s %10s  (cycles per MMA: issue-loop only / until all complete)\n", "N", "nacc", "ctas/SM", "issue", "complete");
    for (int N : {32, 64, 128, 256})
        for (int nacc : {1, 2, 4}) {
            if (N * nacc > 512) continue;
            for (int per_sm : {1, 2, 4}) {
                if (N * nacc * per_sm > 512) continue;
                const size_t smem = 50 * 1024;   // 4 CTAs fit
                rate_kernel<<<148 * per_sm, 128, smem>>>(N, nacc, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("%5d %5d %8d | %10.1f %10.1f\n", N, nacc, per_sm, (double)h[0] / iters, (double)h[1] / iters);
            }
        }
    return 0;
}
