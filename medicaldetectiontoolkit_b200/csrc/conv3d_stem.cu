// Direct kernels for the network stem (Cin <= 4, e.g. the 1-channel image -> 18 feature maps 3x3x3 conv of models/backbone.py:48):
// with one input channel the GEMM view has K = 27 — padding it to tensor-core shape wastes 94 % of every MMA and the operand split writes
// 16x the input — so this layer is bandwidth-bound SIMT work: each output voxel reads its 27*Cin neighbours (L1/L2 hits) and writes Cout
// floats.  fprop: one thread per voxel, weights broadcast from shared memory, the warp's 32 x Cout outputs staged through shared memory
// for fully coalesced stores.  wgrad: one block per output line, dy line + x halo lines in shared memory, each thread owns a few
// (cout, tap) accumulators, block-level partials combined with one atomicAdd per element per block.
#include "conv3d_common.cuh"

namespace mdt {

constexpr int kStemMaxCout = 32, kStemMaxTaps = 27, kStemMaxCin = 4;

__global__ void __launch_bounds__(256) stem_fprop_kernel(ConvGeom g, const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                                                        float *__restrict__ y, int relu) {
    __shared__ float s_w[kStemMaxTaps * kStemMaxCin * kStemMaxCout];   // [tap][ci][co]
    __shared__ float s_b[kStemMaxCout];
    __shared__ float s_out[8][32 * kStemMaxCout];
    const int T = g.kd * g.kh * g.kw;
    for (int i = threadIdx.x; i < T * g.cin * g.cout; i += blockDim.x) {
        const int co = i % g.cout, ci = (i / g.cout) % g.cin, t = i / (g.cout * g.cin);
        s_w[i] = w[((size_t)co * g.cin + ci) * T + t];
    }
    if ((int)threadIdx.x < g.cout) s_b[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
    __syncthreads();
    const long long M = (long long)g.n * g.od * g.oh * g.ow;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long base = (long long)blockIdx.x * 256; base < M; base += (long long)gridDim.x * 256) {
        const long long m = base + threadIdx.x;
        float acc[kStemMaxCout];
#pragma unroll
        for (int c = 0; c < kStemMaxCout; ++c) acc[c] = 0.f;
        if (m < M) {
            int ow = m % g.ow; long long r = m / g.ow;
            int oh = r % g.oh; r /= g.oh;
            int od = r % g.od;
            const int n = (int)(r / g.od);
            for (int kd = 0; kd < g.kd; ++kd) {
                const int d = od * g.sd - g.pd + kd;
                if (d < 0 || d >= g.d) continue;
                for (int kh = 0; kh < g.kh; ++kh) {
                    const int h = oh * g.sh - g.ph + kh;
                    if (h < 0 || h >= g.h) continue;
                    for (int kw = 0; kw < g.kw; ++kw) {
                        const int ww = ow * g.sw - g.pw + kw;
                        if (ww < 0 || ww >= g.w) continue;
                        const float *px = x + ((((long long)n * g.d + d) * g.h + h) * g.w + ww) * g.cin;
                        const float *pw = s_w + ((kd * g.kh + kh) * g.kw + kw) * g.cin * g.cout;
                        for (int ci = 0; ci < g.cin; ++ci) {
                            const float xv = __ldg(px + ci);
#pragma unroll
                            for (int c = 0; c < kStemMaxCout; ++c)
                                if (c < g.cout) acc[c] = fmaf(xv, pw[ci * g.cout + c], acc[c]);
                        }
                    }
                }
            }
        }
        // stage the warp's 32 x cout outputs and store them as one contiguous run
        float *so = s_out[warp];
#pragma unroll
        for (int c = 0; c < kStemMaxCout; ++c)
            if (c < g.cout) {
                float v = acc[c] + s_b[c];
                so[lane * g.cout + c] = relu ? fmaxf(v, 0.f) : v;
            }
        __syncwarp();
        const long long wbase = base + warp * 32;
        const long long nvalid = (M - wbase) < 32 ? (M - wbase) : 32;
        if (nvalid > 0)
            for (int i = lane; i < (int)nvalid * g.cout; i += 32) y[wbase * g.cout + i] = so[i];
        __syncwarp();
    }
}

// one block per (n, od, oh) output line; requires sw == 1
__global__ void __launch_bounds__(256) stem_wgrad_kernel(ConvGeom g, const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ dw,
                                                        long long lines_per_block) {
    extern __shared__ float smem[];
    const int T = g.kd * g.kh * g.kw;
    const int XW = g.ow + g.kw - 1;                       // halo line length
    float *s_dy = smem;                                   // [ow][cout]
    float *s_x = smem + (size_t)g.ow * g.cout;            // [kd*kh][XW][cin]
    const int npairs = g.cout * g.cin * T;
    constexpr int kPer = 8;                               // (cout, ci, tap) accumulators per thread: supports up to 2048 weights
    float acc[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) acc[q] = 0.f;
    const long long nlines = (long long)g.n * g.od * g.oh;
    const long long l0 = (long long)blockIdx.x * lines_per_block, l1 = min(nlines, l0 + lines_per_block);
    for (long long line = l0; line < l1; ++line) {
        const int oh = line % g.oh; long long r = line / g.oh;
        const int od = r % g.od;
        const int n = (int)(r / g.od);
        __syncthreads();
        for (int i = threadIdx.x; i < g.ow * g.cout; i += blockDim.x) s_dy[i] = __ldg(dy + line * g.ow * g.cout + i);
        for (int i = threadIdx.x; i < g.kd * g.kh * XW * g.cin; i += blockDim.x) {
            const int ci = i % g.cin; int t = i / g.cin;
            const int xw = t % XW; t /= XW;
            const int kh = t % g.kh, kd = t / g.kh;
            const int d = od * g.sd - g.pd + kd, h = oh * g.sh - g.ph + kh, ww = xw - g.pw;
            s_x[i] = (d >= 0 && d < g.d && h >= 0 && h < g.h && ww >= 0 && ww < g.w)
                         ? __ldg(x + ((((long long)n * g.d + d) * g.h + h) * g.w + ww) * g.cin + ci) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int pidx = threadIdx.x + q * 256;
            if (pidx >= npairs) break;
            const int t = pidx % T, ci = (pidx / T) % g.cin, co = pidx / (T * g.cin);
            const int kw = t % g.kw, kdh = t / g.kw;
            const float *px = s_x + ((size_t)kdh * XW + kw) * g.cin + ci;
            float a = 0.f;
            for (int v = 0; v < g.ow; ++v) a = fmaf(s_dy[v * g.cout + co], px[(size_t)v * g.cin], a);
            acc[q] += a;
        }
    }
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int pidx = threadIdx.x + q * 256;
        if (pidx < npairs && acc[q] != 0.f) atomicAdd(dw + pidx, acc[q]);   // dw layout [co][ci][tap] == pidx order
    }
}

bool conv_stem_supported(const ConvGeom &g, int pass) {
    if (g.cin > kStemMaxCin || g.cout > kStemMaxCout || g.kd * g.kh * g.kw > kStemMaxTaps) return false;
    if (pass == 0) return true;
    if (pass == 2) {
        const size_t smem = ((size_t)g.ow * g.cout + (size_t)g.kd * g.kh * (g.ow + g.kw - 1) * g.cin) * sizeof(float);
        return g.sw == 1 && g.cout * g.cin * g.kd * g.kh * g.kw <= 2048 && smem <= 96 * 1024;
    }
    return false;   // dgrad of the stem is never needed (the image has no gradient); generic kernels cover it if it is
}

int conv_stem_fprop(const ConvGeom &g, const float *x, const float *w, const float *bias, float *y, int relu, cudaStream_t st) {
    const long long M = (long long)g.n * g.od * g.oh * g.ow;
    long long blocks = ceil_div<long long>(M, 256);
    if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
    stem_fprop_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, x, w, bias, y, relu);
    return launch_status();
}

int conv_stem_wgrad(const ConvGeom &g, const float *x, const float *dy, float *dw, float *db, cudaStream_t st) {
    const int T = g.kd * g.kh * g.kw;
    cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g.cout * g.cin * T, st);
    if (e != cudaSuccess) return (int)e;
    const long long nlines = (long long)g.n * g.od * g.oh;
    long long blocks = (long long)num_sms() * 8;
    if (blocks > nlines) blocks = nlines;
    const size_t smem = ((size_t)g.ow * g.cout + (size_t)g.kd * g.kh * (g.ow + g.kw - 1) * g.cin) * sizeof(float);
    static bool attr[kMaxDevices] = {};
    if (!ensure_smem_attr(stem_wgrad_kernel, 96 * 1024, attr)) return MDT_EDRIVER;
    stem_wgrad_kernel<<<(unsigned)blocks, 256, smem, st>>>(g, x, dy, dw, ceil_div<long long>(nlines, blocks));
    int rc = launch_status();
    if (rc) return rc;
    if (db) return conv_bias_grad(g, dy, db, st);
    return MDT_OK;
}

}  // namespace mdt
