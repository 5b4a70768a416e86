// mma_major_probe — what does one tcgen05.mma (K = 16, bf16) cost as a function of WHERE and HOW its operands lie?
// Operand majors (K-major vs MN-major shared-memory layouts, 128-byte swizzle), A from shared memory vs from tensor memory, M = 64 / 128,
// N = 64 .. 256, and the wgrad kernel's overlapped-chunk trick (LBO = one row).  One elected thread issues 3600 back-to-back MMAs on resident
// operands; cycles per MMA until the last one completes.  Decides whether the weight-gradient kernel (both operands MN-major, measured
// ~188 cycles for M = 128, N = 192) is bound by the shared-memory read path of transposed operands.
#include <cstdio>
#include <cstdlib>
#include "../medicaldetectiontoolkit_b200/csrc/tc_common.cuh"
using namespace mdt;
using namespace mdt::tc;

struct Cfg { int M, N, amaj, bmaj, a_tmem, overlap, total, a_off, b_off; };

__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(128) probe_kernel(Cfg c, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (224 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&done, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc_bf16(c.M, c.N, c.a_tmem ? 0 : c.amaj, c.bmaj);
        const uint32_t a_addr = smem_u32(smem + c.a_off), b_addr = smem_u32(smem + c.b_off);
        // K-major SW128: 8-row groups 1024 bytes apart, K16 step = 32 bytes inside the swizzled row.
        // MN-major SW128: 64-element chunks `lbo` apart, 8-K-row groups 1024 bytes apart, K16 step = 16 rows = 2048 bytes.
        const uint32_t lbo_b = c.overlap ? 128u : 16384u;
        const uint64_t adesc0 = c.amaj ? make_smem_desc(a_addr, 16384, 1024, 2) : make_smem_desc(a_addr, 16, 1024, 2);
        const uint64_t bdesc0 = c.bmaj ? make_smem_desc(b_addr, lbo_b, 1024, 2) : make_smem_desc(b_addr, 16, 1024, 2);
        const uint64_t astep = c.amaj ? 128 : 2, bstep = c.bmaj ? 128 : 2;
        const long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < c.total; ++i) {
                const uint64_t s = (uint64_t)(i & 3);
                if (c.a_tmem) umma_ts(tmem, tmem + 256 + (uint32_t)s * 8u, bdesc0 + s * bstep, idesc);
                else umma_ss(tmem, adesc0 + s * astep, bdesc0 + s * bstep, idesc);
            }
            umma_commit(&done);
        }
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
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    printf("%4s %4s %8s %8s %8s %8s | %s\n", "M", "N", "A", "B", "A-from", "overlap", "cycles per MMA");
    const char *mj[2] = {"K-major", "MN-major"};
    for (int M : {64, 128})
        for (int N : {64, 128, 192, 256})
            for (int variant = 0; variant < 7; ++variant) {
                Cfg c{M, N, 0, 0, 0, 0, 3600, 0, 32 * 1024};
                if (variant == 1) c.bmaj = 1;
                if (variant == 2) c.amaj = 1;
                if (variant == 3) { c.amaj = 1; c.bmaj = 1; }
                if (variant == 4) { c.amaj = 1; c.bmaj = 1; c.overlap = 1; }
                if (variant == 5) { c.a_tmem = 1; c.bmaj = 0; }
                if (variant == 6) { c.a_tmem = 1; c.bmaj = 1; }
                probe_kernel<<<148, 128, 226 * 1024>>>(c, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("%4d %4d %8s %8s %8s %8d | %8.1f\n", M, N, c.a_tmem ? "-" : mj[c.amaj], mj[c.bmaj], c.a_tmem ? "tmem" : "smem", c.overlap, (double)h / c.total);
            }
    // does the PLACEMENT of the two operands in shared memory matter?  (tools/mma_pipe_probe had A at 0 and B at 64 KiB and measured
    // max(N/2, (M+N)/4); the table above has B at 32 KiB and measures N/2 + M/4 + 11)
    printf("\nplacement sweep, M = 128, both K-major: cycles per MMA\n%8s %8s | %8s %8s %8s\n", "A at", "B at", "N=64", "N=128", "N=192");
    for (int a_off : {0, 16 * 1024, 64 * 1024})
        for (int b_kib : {16, 32, 48, 64, 80, 96, 112, 128, 144, 160, 192}) {
            const int b_off = a_off + b_kib * 1024;
            if (b_off + 32 * 1024 > 224 * 1024) continue;
            printf("%8d %8d |", a_off, b_off);
            for (int N : {64, 128, 192}) {
                Cfg c{128, N, 0, 0, 0, 0, 3600, a_off, b_off};
                probe_kernel<<<148, 128, 226 * 1024>>>(c, d);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
                long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf(" %8.1f", (double)h / c.total);
            }
            printf("\n");
        }
    return 0;
}
