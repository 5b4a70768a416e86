// mma_pipe_probe — what breaks the tcgen05.mma pipeline?  One warp per CTA issues groups of G MMAs (M = 128, K = 16, bf16, SW128 K-major operands
// resident in shared memory; warp-uniform control flow, elected lane) and between groups optionally: tcgen05.commit to an mbarrier (nobody
// waits), a switch to a second accumulator, a different A buffer, tcgen05.fence::after_thread_sync.  Reports cycles per MMA until all complete.
// Answers why the tap-stacked conv kernel (conv3d_tcw.cu) spends ~500 cycles per pipeline stage beyond its MMA time (profiles/r02_*).
#include <cstdio>
#include <cstdlib>
#include "../medicaldetectiontoolkit_b200/csrc/tc_common.cuh"
using namespace mdt;
using namespace mdt::tc;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}

struct Cfg { int N, G, groups, commit, sw_acc, sw_buf, fence, nb; };

__global__ void __launch_bounds__(128) probe_kernel(Cfg c, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bars[8], done;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (4 * 16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_init(&done, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc_bf16(128, c.N, 0, 0);
        const uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29), lbo = 1u << 16;
        const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + 4 * 16384) >> 4;
        const long long t0 = clock64();
        for (int g = 0; g < c.groups; ++g) {
            if (c.fence) tc_fence_after();
            const uint32_t d = tmem + (c.sw_acc ? (uint32_t)(g & 1) * (uint32_t)c.N : 0u);
            const uint32_t a = a16 + (c.sw_buf ? (uint32_t)(g & 3) * 1024u : 0u);
            if (elect_one()) {
                for (int i = 0; i < c.G; ++i) umma2(d, (a + 2u * (i & 3)) | lbo, (b16 + 2u * (i & 3)) | lbo, hi, idesc, 1);
                for (int k = 0; k < c.commit; ++k) umma_commit(&bars[(g + k) & 7]);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&done);
        __syncwarp();
        mbar_wait(&done, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    printf("%4s %4s %6s %6s %6s %6s | %s\n", "N", "G", "commit", "swacc", "swbuf", "fence", "cycles per MMA (until complete)   cycles per group");
    const int total = 3600;
    for (int N : {64, 112, 128, 224})
        for (int G : {1, 3, 9, 36, 3600})
            for (int variant = 0; variant < 6; ++variant) {
                Cfg c{N, G, total / G, 0, 0, 0, 0, 0};
                if (variant == 1) c.commit = 1;
                if (variant == 2) { c.commit = 1; c.sw_acc = 1; }
                if (variant == 3) { c.commit = 1; c.sw_acc = 1; c.sw_buf = 1; c.fence = 1; }
                if (variant == 4) { c.commit = 2; c.sw_acc = 1; c.sw_buf = 1; c.fence = 1; }
                if (variant == 5) { c.commit = 0; c.sw_acc = 1; c.sw_buf = 1; }
                if (G == 3600 && variant > 0) continue;
                if (2 * N > 512) c.sw_acc = 0;
                probe_kernel<<<148, 128, 110 * 1024>>>(c, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("%4d %4d %6d %6d %6d %6d | %8.1f %10.1f\n", N, G, c.commit, c.sw_acc, c.sw_buf, c.fence, (double)h / (c.groups * c.G), (double)h / c.groups);
            }
    return 0;
}
