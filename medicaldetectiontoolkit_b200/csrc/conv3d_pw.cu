// Pointwise (1x1x1, stride 1, no padding) conv3d: fprop / dgrad / wgrad as fp32 STREAMING kernels (algo 4).
//
// Why not the tensor cores: a 1x1x1 conv with the toolkit's channel counts (18 -> 36, 18 -> 72, 36 -> 144, 36 -> 2; reference
// models/backbone.py:27-60 ResBlock conv1/conv3, :148-156 P*_conv1 laterals, models/retina_unet.py final_conv) does
// Cin * Cout / (4 * (Cin + Cout)) = 3 .. 7 FMA per byte it has to move: it is HBM-bound.  The tcgen05 path pays for it twice: an operand-split
// pass (fp32 -> 2 bf16 planes padded to 16 channels = MORE bytes than the fp32 rows) and a 128 x 48 x 32 MMA per 128 voxels whose epilogue, not
// its math, sets the time (profiles/r02_layer_bench.txt: 18 -> 36 at 2x128^3 1.14 ms = 2.3 GB / 2.0 TB/s).  Here each pass reads its fp32
// rows once, multiplies in exact fp32 FMAs with the weights resident in shared memory, and writes once.
//
//   pw_gemm_kernel : out[v, n] = sum_k in[v, k] * B[k, n]  (+ bias[n]) (+ res[v, n]) (ReLU) (+ the (hi, lo) bf16 split planes of out for a
//                    tcgen05 consumer, layout of split_rows_kernel).  fprop: in = x, B = w^T.  dgrad: in = dy (ReLU-masked on load by the
//                    forward output, masked copy optionally written out as the residual's gradient), B = w.
//   pw_wgrad_kernel: dw[co, ci] = sum_v dy[v, co] * x[v, ci], db[co] = sum_v dy[v, co]; persistent CTAs keep 4 x 4 register tiles over their
//                    share of the voxels, dump one partial per CTA, pw_wgrad_reduce_kernel adds the partials in a fixed order (deterministic).
//
// Data layout: channels-last rows ([voxel][channel] fp32), a tile of TV consecutive voxels is one contiguous span of global memory: all global
// accesses are flat, coalesced 64-bit accesses; the per-voxel (row) view exists only in shared memory (row stride an odd number of 16-byte
// chunks: conflict-free 128-bit row reads).
#include "conv3d_common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace mdt {
int conv_tc_kpad(int channels);

constexpr int kPwThreads = 128;
constexpr int kPwMaxQ = 12;        // output-channel quads (4 channels) a thread accumulates per chunk
constexpr int kPwMaxMacs = 6144;   // Cin * Cout the kernels accept (algo 4 forced)
constexpr int kPwAutoMacs = 2592;  // ... and up to which `auto` prefers them: measured against the tcgen05 path in tools/pw_bench.py
                                   // (profiles/r02_pw_bench.txt): 18 -> 36/72, 72 -> 18, 36 -> 2 win, 36 -> 144 (3 output chunks, one CTA per SM) loses
constexpr int kPwU = 8;          // independent global loads in flight per thread in the flat passes
constexpr int kPwWgTile = 128;     // voxels per wgrad tile
constexpr int kPwWgMaxTpt = 3;     // 4 x 4 tiles per wgrad thread

struct PwGemmParams {
    const float *in, *w, *bias, *res, *relu_of;
    float *out, *masked_out;
    __nv_bfloat16 *out_split;
    long long V, tiles;
    int K, N, KQ;            // contraction channels, output channels, ceil(K / 4)
    int SI, SO;              // shared-memory row strides (16-byte chunks, odd)
    int chunks, NP4;         // output chunks of NQ quads; NP4 = chunks * NQ
    int w_kn;                // 1: w is [K][N] (dgrad), 0: w is [N][K] (fprop)
    int relu, line_w, out_kg, planes;
    int alias;               // output staging reuses the (consumed) input tile: single chunk and rows no wider than the input's
    int res_smem;            // residual rows are staged in shared memory by cp.async (else read from global memory in the epilogue)
    int rs_stride;           // floats per residual row in shared memory (N rounded up to 4: keeps the buffers after it 16-byte aligned)
    int vec_in, vec_out;     // 64-bit global accesses allowed (even channel count + aligned pointers)
};

__device__ __forceinline__ void pw_fma4(float (&acc)[4], float x, const float4 &w) {
    acc[0] = fmaf(x, w.x, acc[0]);
    acc[1] = fmaf(x, w.y, acc[1]);
    acc[2] = fmaf(x, w.z, acc[2]);
    acc[3] = fmaf(x, w.w, acc[3]);
}

// position of a thread's u-th element in a flat pass over rows of C floats: advanced incrementally (no per-element division)
struct PwCursor {
    int v, k, dv, dk, C;
    __device__ __forceinline__ void init(int first, int step, int C_) {
        C = C_;
        v = first / C; k = first - v * C;
        dv = step / C; dk = step - dv * C;
    }
    __device__ __forceinline__ void next() {
        v += dv; k += dk;
        if (k >= C) { k -= C; ++v; }
    }
};

__device__ __forceinline__ uint32_t pw_smem_u32(const void *q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(pw_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(pw_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Pipeline: the input rows of tile i+1 and the residual rows of tile i are in flight (cp.async, straight into their shared-memory layout)
// while tile i is multiplied; nothing in the steady state waits on a global load it issued itself.  (The first version loaded, multiplied and
// stored in sequence: 2.2 TB/s, latency-bound, profiles/r02_pw_bench.txt.)
template <int VPT, int NQ>
__global__ void __launch_bounds__(kPwThreads) pw_gemm_kernel(const PwGemmParams p) {
    extern __shared__ __align__(16) float4 pw_smem[];
    constexpr int TV = kPwThreads * VPT;
    const int tid = threadIdx.x;
    const int K = p.K, N = p.N;
    const int SI4 = p.SI * 4, SO4 = p.SO * 4;
    float4 *Bs = pw_smem;                                   // [KQ * 4][NP4]
    float4 *bias_s = Bs + (size_t)p.KQ * 4 * p.NP4;         // [NP4]
    float4 *xs = bias_s + p.NP4;                            // [2][TV][SI]
    float *rs_f = reinterpret_cast<float *>(xs + (size_t)2 * TV * p.SI);          // [TV * N] residual rows (flat), if any
    float4 *ys_own = reinterpret_cast<float4 *>(rs_f + (p.res_smem ? (size_t)TV * p.rs_stride : 0));   // [TV][SO] unless aliased to the input tile

    {   // weights (transposed to [k][n] if needed) and bias, zero padded
        float *Bf = reinterpret_cast<float *>(Bs);
        const int np = p.NP4 * 4, tot = p.KQ * 4 * np;
        for (int i = tid; i < tot; i += kPwThreads) {
            const int k = i / np, n = i - k * np;
            Bf[i] = (k < K && n < N) ? __ldg(p.w + (p.w_kn ? (size_t)k * N + n : (size_t)n * K + k)) : 0.f;
        }
        float *bf = reinterpret_cast<float *>(bias_s);
        for (int i = tid; i < np; i += kPwThreads) bf[i] = (p.bias && i < N) ? __ldg(p.bias + i) : 0.f;
    }
    // the flat passes touch the same tile-relative positions on every tile: cursors are set up once
    const int ein = p.vec_in ? 2 : 1, eout = p.vec_out ? 2 : 1;
    PwCursor cin0, cout0, csp0;
    cin0.init(tid * ein, kPwThreads * ein, K);
    cout0.init(tid * eout, kPwThreads * eout, N);
    const int groups = p.out_kg / 8;
    csp0.init(tid, kPwThreads, groups);
    const bool async_in = p.relu_of == nullptr;   // a ReLU mask has to be applied to the values: those tiles are loaded through registers

    // input rows of tile t -> buffer b (cp.async; rows past the end of the tensor and padding channels are zeroed with plain stores)
    auto issue_in = [&](long long t, int b) {
        const long long v0 = t * TV;
        const int nvalid = (int)min((long long)TV, p.V - v0);
        float *xb = reinterpret_cast<float *>(xs + (size_t)b * TV * p.SI);
        const float *src = p.in + v0 * K;
        const int tot = nvalid * K;
        PwCursor c = cin0;
        if (p.vec_in) {
            for (int f = tid * 2; f < tot; f += kPwThreads * 2) { cp_async8(xb + c.v * SI4 + c.k, src + f); c.next(); }
        } else {
            for (int f = tid; f < tot; f += kPwThreads) { cp_async4(xb + c.v * SI4 + c.k, src + f); c.next(); }
        }
    };
    auto zero_pad = [&](int nvalid, int b) {
        if (K < p.KQ * 4 || nvalid < TV) {
            float *xb = reinterpret_cast<float *>(xs + (size_t)b * TV * p.SI);
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const int v = tid + j * kPwThreads;
                float *row = xb + v * SI4;
                for (int k = (v < nvalid ? K : 0); k < p.KQ * 4; ++k) row[k] = 0.f;
            }
        }
    };
    auto issue_res = [&](long long t) {
        const long long v0 = t * TV;
        const int nvalid = (int)min((long long)TV, p.V - v0);
        const float *src = p.res + v0 * N;
        const int tot = nvalid * N;
        if (p.vec_out) {
            for (int f = tid * 2; f < tot; f += kPwThreads * 2) cp_async8(rs_f + f, src + f);
        } else {
            for (int f = tid; f < tot; f += kPwThreads) cp_async4(rs_f + f, src + f);
        }
    };

    int buf = 0;
    if (async_in && (long long)blockIdx.x < p.tiles) issue_in(blockIdx.x, 0);
    cp_async_commit();

    for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x, buf ^= 1) {
        const long long v0 = t * TV;
        const int nvalid = (int)min((long long)TV, p.V - v0);
        float4 *xb = xs + (size_t)buf * TV * p.SI;
        float *xb_f = reinterpret_cast<float *>(xb);
        float4 *ys = p.alias ? xb : ys_own;
        float *ys_f = reinterpret_cast<float *>(ys);
        __syncthreads();   // previous tile's staging (the other input buffer, the residual rows) fully consumed; weights visible
        if (async_in && t + gridDim.x < p.tiles) issue_in(t + gridDim.x, buf ^ 1);
        if (p.res_smem) issue_res(t);
        cp_async_commit();
        if (!async_in) {   // ---- masked input (dgrad through a fused ReLU): flat coalesced pass through registers, kPwU loads in flight
            const float *src = p.in + v0 * K;
            const float *msk = p.relu_of + v0 * K;
            float *mo = p.masked_out ? p.masked_out + v0 * K : nullptr;
            const int tot = nvalid * K;
            PwCursor c = cin0;
            if (p.vec_in) {
                for (int f0 = tid * 2; f0 < tot; f0 += kPwThreads * 2 * kPwU) {
                    float2 a[kPwU], y[kPwU];
#pragma unroll
                    for (int u = 0; u < kPwU; ++u) {
                        const int f = f0 + u * kPwThreads * 2;
                        if (f < tot) {
                            a[u] = __ldg(reinterpret_cast<const float2 *>(src + f));
                            y[u] = __ldg(reinterpret_cast<const float2 *>(msk + f));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kPwU; ++u) {
                        const int f = f0 + u * kPwThreads * 2;
                        if (f < tot) {
                            if (!(y[u].x > 0.f)) a[u].x = 0.f;
                            if (!(y[u].y > 0.f)) a[u].y = 0.f;
                            if (mo) *reinterpret_cast<float2 *>(mo + f) = a[u];
                            *reinterpret_cast<float2 *>(xb_f + c.v * SI4 + c.k) = a[u];
                        }
                        c.next();
                    }
                }
            } else {
                for (int f0 = tid; f0 < tot; f0 += kPwThreads * kPwU) {
                    float a[kPwU], y[kPwU];
#pragma unroll
                    for (int u = 0; u < kPwU; ++u) {
                        const int f = f0 + u * kPwThreads;
                        if (f < tot) {
                            a[u] = __ldg(src + f);
                            y[u] = __ldg(msk + f);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kPwU; ++u) {
                        const int f = f0 + u * kPwThreads;
                        if (f < tot) {
                            if (!(y[u] > 0.f)) a[u] = 0.f;
                            if (mo) mo[f] = a[u];
                            xb_f[c.v * SI4 + c.k] = a[u];
                        }
                        c.next();
                    }
                }
            }
        }
        zero_pad(nvalid, buf);
        cp_async_wait<1>();   // everything but the group committed above: this tile's input rows have landed
        __syncthreads();

        for (int c = 0; c < p.chunks; ++c) {
            float acc[VPT][NQ][4];
#pragma unroll
            for (int j = 0; j < VPT; ++j)
#pragma unroll
                for (int q = 0; q < NQ; ++q) acc[j][q][0] = acc[j][q][1] = acc[j][q][2] = acc[j][q][3] = 0.f;
            const float4 *brow = Bs + c * NQ;
            const float4 *xrow = xb + tid * p.SI;
            const int xstep = kPwThreads * p.SI, bstep = 4 * p.NP4;
            for (int kq = 0; kq < p.KQ; ++kq, brow += bstep) {
                float4 xv[VPT];
#pragma unroll
                for (int j = 0; j < VPT; ++j) xv[j] = xrow[j * xstep + kq];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const float4 w4 = brow[kk * p.NP4 + q];
#pragma unroll
                        for (int j = 0; j < VPT; ++j) {
                            const float x = kk == 0 ? xv[j].x : kk == 1 ? xv[j].y : kk == 2 ? xv[j].z : xv[j].w;
                            pw_fma4(acc[j][q], x, w4);
                        }
                    }
                }
            }
            if (p.alias) __syncthreads();   // every thread is done reading the input tile before its rows are overwritten
#pragma unroll
            for (int j = 0; j < VPT; ++j)
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    ys[(tid + j * kPwThreads) * p.SO + c * NQ + q] = make_float4(acc[j][q][0], acc[j][q][1], acc[j][q][2], acc[j][q][3]);
        }
        cp_async_wait<0>();   // residual rows of this tile (and, long since, the next tile's input)
        __syncthreads();

        {   // ---- epilogue: flat coalesced pass over nvalid * N floats
            float *dst = p.out + v0 * N;
            const float *bf = reinterpret_cast<const float *>(bias_s);
            const int tot = nvalid * N;
            const bool back = p.out_split != nullptr, res = p.res != nullptr, rsm = p.res_smem != 0;
            const float *resg = res ? p.res + v0 * N : nullptr;
            PwCursor c = cout0;
            if (p.vec_out) {
                for (int f = tid * 2; f < tot; f += kPwThreads * 2) {
                    float *sp = ys_f + c.v * SO4 + c.k;
                    float2 a = *reinterpret_cast<const float2 *>(sp);
                    const float2 b2 = *reinterpret_cast<const float2 *>(bf + c.k);
                    a.x += b2.x;
                    a.y += b2.y;
                    if (res) {
                        const float2 r = rsm ? *reinterpret_cast<const float2 *>(rs_f + f) : __ldg(reinterpret_cast<const float2 *>(resg + f));
                        a.x += r.x;
                        a.y += r.y;
                    }
                    if (p.relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); }
                    *reinterpret_cast<float2 *>(dst + f) = a;
                    if (back) *reinterpret_cast<float2 *>(sp) = a;
                    c.next();
                }
            } else {
                for (int f = tid; f < tot; f += kPwThreads) {
                    float *sp = ys_f + c.v * SO4 + c.k;
                    float a = *sp + bf[c.k];
                    if (res) a += rsm ? rs_f[f] : __ldg(resg + f);
                    if (p.relu) a = fmaxf(a, 0.f);
                    dst[f] = a;
                    if (back) *sp = a;
                    c.next();
                }
            }
            if (back) {   // (hi, lo) bf16 planes of the result, [line][plane][w][out_kg] (split_rows_kernel's layout, padding channels zero)
                __syncthreads();
                const int tot2 = nvalid * groups;
                const long long plane_stride = (long long)p.line_w * p.out_kg;
                const long long line0 = v0 / p.line_w;
                const int w0 = (int)(v0 - line0 * p.line_w);
                PwCursor g = csp0;
                for (int i = tid; i < tot2; i += kPwThreads) {
                    const int c0 = g.k * 8;
                    const float *sp = ys_f + g.v * SO4 + c0;
                    float a[8];
                    if (c0 + 8 <= N) {
                        const float4 a0 = *reinterpret_cast<const float4 *>(sp), a1 = *reinterpret_cast<const float4 *>(sp + 4);
                        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) a[k] = (c0 + k < N) ? sp[k] : 0.f;
                    }
                    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        hi[k] = __float2bfloat16_rn(a[k]);
                        lo[k] = __float2bfloat16_rn(a[k] - __bfloat162float(hi[k]));
                    }
                    int w = w0 + g.v, lr = 0;
                    if (w >= p.line_w) { lr = w / p.line_w; w -= lr * p.line_w; }
                    const long long off = ((line0 + lr) * p.planes * p.line_w + w) * p.out_kg + c0;
                    *reinterpret_cast<uint4 *>(p.out_split + off) = *reinterpret_cast<const uint4 *>(hi);
                    if (p.planes > 1) *reinterpret_cast<uint4 *>(p.out_split + off + plane_stride) = *reinterpret_cast<const uint4 *>(lo);
                    g.next();
                }
            }
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------------------------------- wgrad
struct PwWgradParams {
    const float *x, *dy, *relu_of;
    float *masked_out;       // optional fp32 copy of the ReLU-masked dy (when no dgrad launch wrote it)
    float *partial;          // [grid][Cout * Cin + Cout]
    long long V, tiles;
    int Cin, Cout, IQ, CQ;   // channel quads
    int TG, G;               // threads per voxel group (each owns TPT 4x4 tiles), voxel groups
    int vec_x, vec_y;
};

template <int TPT>
__global__ void __launch_bounds__(kPwThreads) pw_wgrad_kernel(const PwWgradParams p) {
    extern __shared__ __align__(16) float4 pw_smem[];
    const int tid = threadIdx.x;
    const int IQ4 = p.IQ * 4, CQ4 = p.CQ * 4;
    const int tile_q = kPwWgTile * (p.IQ + p.CQ);          // float4 per buffer: x rows [tile][IQ], then dy rows [tile][CQ]
    const int grp = tid / p.TG, r = tid - grp * p.TG;
    const int ntile4 = p.CQ * p.IQ;
    const bool active = grp < p.G;
    int cq[TPT], iq[TPT];
    bool ok[TPT];
#pragma unroll
    for (int j = 0; j < TPT; ++j) {
        const int id = r + j * p.TG;
        ok[j] = active && id < ntile4;
        cq[j] = ok[j] ? id / p.IQ : 0;
        iq[j] = ok[j] ? id - cq[j] * p.IQ : 0;
    }
    float acc[TPT][4][4], bacc[TPT][4];
#pragma unroll
    for (int j = 0; j < TPT; ++j)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            bacc[j][a] = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[j][a][b] = 0.f;
        }

    PwCursor cx0, cy0;
    const int ex = p.vec_x ? 2 : 1, ey = p.vec_y ? 2 : 1;
    cx0.init(tid * ex, kPwThreads * ex, p.Cin);
    cy0.init(tid * ey, kPwThreads * ey, p.Cout);
    // padding channels of both buffers stay zero for the whole kernel (the loads below only write real channels)
    for (int i = tid; i < 2 * tile_q * 4; i += kPwThreads) reinterpret_cast<float *>(pw_smem)[i] = 0.f;
    __syncthreads();
    const bool async_y = p.relu_of == nullptr;   // a ReLU mask is applied to the values: those rows go through registers

    auto issue = [&](long long t, int b) {
        const long long v0 = t * kPwWgTile;
        const int nvalid = (int)min((long long)kPwWgTile, p.V - v0);
        float *xb = reinterpret_cast<float *>(pw_smem + (size_t)b * tile_q);
        float *db = xb + (size_t)kPwWgTile * IQ4;
        {
            const float *src = p.x + v0 * p.Cin;
            const int tot = nvalid * p.Cin;
            PwCursor c = cx0;
            if (p.vec_x) {
                for (int f = tid * 2; f < tot; f += kPwThreads * 2) { cp_async8(xb + c.v * IQ4 + c.k, src + f); c.next(); }
            } else {
                for (int f = tid; f < tot; f += kPwThreads) { cp_async4(xb + c.v * IQ4 + c.k, src + f); c.next(); }
            }
        }
        if (async_y) {
            const float *src = p.dy + v0 * p.Cout;
            const int tot = nvalid * p.Cout;
            PwCursor c = cy0;
            if (p.vec_y) {
                for (int f = tid * 2; f < tot; f += kPwThreads * 2) { cp_async8(db + c.v * CQ4 + c.k, src + f); c.next(); }
            } else {
                for (int f = tid; f < tot; f += kPwThreads) { cp_async4(db + c.v * CQ4 + c.k, src + f); c.next(); }
            }
        }
    };

    int buf = 0;
    if ((long long)blockIdx.x < p.tiles) issue(blockIdx.x, 0);
    cp_async_commit();
    for (long long t = blockIdx.x; t < p.tiles; t += gridDim.x, buf ^= 1) {
        const long long v0 = t * kPwWgTile;
        const int nvalid = (int)min((long long)kPwWgTile, p.V - v0);
        const float4 *xs = pw_smem + (size_t)buf * tile_q;
        const float4 *ds = xs + (size_t)kPwWgTile * p.IQ;
        __syncthreads();   // the other buffer's rows are consumed
        if (t + gridDim.x < p.tiles) issue(t + gridDim.x, buf ^ 1);
        cp_async_commit();
        if (!async_y) {
            float *ds_f = reinterpret_cast<float *>(pw_smem + (size_t)buf * tile_q) + (size_t)kPwWgTile * IQ4;
            const float *sy = p.dy + v0 * p.Cout;
            const float *msk = p.relu_of + v0 * p.Cout;
            float *mo = p.masked_out ? p.masked_out + v0 * p.Cout : nullptr;
            const int toty = nvalid * p.Cout;
            PwCursor c = cy0;
            if (p.vec_y) {
                for (int f0 = tid * 2; f0 < toty; f0 += kPwThreads * 2 * kPwU) {
                    float2 a[kPwU], y[kPwU];
#pragma unroll
                    for (int u = 0; u < kPwU; ++u) {
                        const int f = f0 + u * kPwThreads * 2;
                        if (f < toty) {
                            a[u] = __ldg(reinterpret_cast<const float2 *>(sy + f));
                            y[u] = __ldg(reinterpret_cast<const float2 *>(msk + f));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kPwU; ++u) {
                        const int f = f0 + u * kPwThreads * 2;
                        if (f < toty) {
                            if (!(y[u].x > 0.f)) a[u].x = 0.f;
                            if (!(y[u].y > 0.f)) a[u].y = 0.f;
                            if (mo) *reinterpret_cast<float2 *>(mo + f) = a[u];
                            *reinterpret_cast<float2 *>(ds_f + c.v * CQ4 + c.k) = a[u];
                        }
                        c.next();
                    }
                }
            } else {
                for (int f = tid; f < toty; f += kPwThreads) {
                    float a = __ldg(sy + f);
                    if (!(__ldg(msk + f) > 0.f)) a = 0.f;
                    if (mo) mo[f] = a;
                    ds_f[c.v * CQ4 + c.k] = a;
                    c.next();
                }
            }
        }
        cp_async_wait<1>();
        __syncthreads();
        if (active) {
            // row pointers are hoisted and 4 voxels are loaded before their FMAs: the loop is bound by shared-memory latency otherwise
            const float4 *dp[TPT], *xp[TPT];
#pragma unroll
            for (int j = 0; j < TPT; ++j) { dp[j] = ds + cq[j]; xp[j] = xs + iq[j]; }
            const int G = p.G, sd = G * p.CQ, sx = G * p.IQ;
            int v = grp;
            int od = grp * p.CQ, ox = grp * p.IQ;
            for (; v + 3 * G < nvalid; v += 4 * G, od += 4 * sd, ox += 4 * sx) {
#pragma unroll
                for (int j = 0; j < TPT; ++j) {
                    if (!ok[j]) continue;
                    float4 d[4], x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { d[u] = dp[j][od + u * sd]; x[u] = xp[j][ox + u * sx]; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        pw_fma4(acc[j][0], d[u].x, x[u]);
                        pw_fma4(acc[j][1], d[u].y, x[u]);
                        pw_fma4(acc[j][2], d[u].z, x[u]);
                        pw_fma4(acc[j][3], d[u].w, x[u]);
                        if (iq[j] == 0) { bacc[j][0] += d[u].x; bacc[j][1] += d[u].y; bacc[j][2] += d[u].z; bacc[j][3] += d[u].w; }
                    }
                }
            }
            for (; v < nvalid; v += G, od += sd, ox += sx) {
#pragma unroll
                for (int j = 0; j < TPT; ++j) {
                    if (!ok[j]) continue;
                    const float4 d = dp[j][od];
                    const float4 x = xp[j][ox];
                    pw_fma4(acc[j][0], d.x, x);
                    pw_fma4(acc[j][1], d.y, x);
                    pw_fma4(acc[j][2], d.z, x);
                    pw_fma4(acc[j][3], d.w, x);
                    if (iq[j] == 0) { bacc[j][0] += d.x; bacc[j][1] += d.y; bacc[j][2] += d.z; bacc[j][3] += d.w; }
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- CTA reduction over the voxel groups (fixed order), then one partial per CTA
    __syncthreads();
    const int cin4 = p.IQ * 4, cout4 = p.CQ * 4;
    const int per = cout4 * cin4 + cout4;
    float *red = reinterpret_cast<float *>(pw_smem);        // [G][per]
    if (active) {
        float *mine = red + (size_t)grp * per;
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
            if (!ok[j]) continue;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                *reinterpret_cast<float4 *>(mine + (size_t)(cq[j] * 4 + a) * cin4 + iq[j] * 4) =
                    make_float4(acc[j][a][0], acc[j][a][1], acc[j][a][2], acc[j][a][3]);
                if (iq[j] == 0) mine[cout4 * cin4 + cq[j] * 4 + a] = bacc[j][a];
            }
        }
    }
    __syncthreads();
    float *out = p.partial + (size_t)blockIdx.x * (p.Cout * p.Cin + p.Cout);
    for (int e = tid; e < per; e += kPwThreads) {
        float s = 0.f;
        for (int g = 0; g < p.G; ++g) s += red[(size_t)g * per + e];
        if (e < cout4 * cin4) {
            const int co = e / cin4, ci = e - co * cin4;
            if (co < p.Cout && ci < p.Cin) out[co * p.Cin + ci] = s;
        } else {
            const int co = e - cout4 * cin4;
            if (co < p.Cout) out[p.Cout * p.Cin + co] = s;
        }
    }
}

// dw / db = sum over the CTAs' partials, in a fixed order (deterministic): a block owns 32 consecutive outputs (lane = output: coalesced
// 128-byte reads), its 8 warps stride over the partials, a shared-memory pass adds the 8 sub-sums
__global__ void __launch_bounds__(256) pw_wgrad_reduce_kernel(const float *__restrict__ partial, int nblk, int nw, int nb, float *__restrict__ dw,
                                                              float *__restrict__ db) {
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    const int per = nw + nb;
    float s0 = 0.f, s1 = 0.f;
    if (e < per) {
        int b = warp;
        for (; b + 8 < nblk; b += 16) {
            s0 += __ldg(partial + (size_t)b * per + e);
            s1 += __ldg(partial + (size_t)(b + 8) * per + e);
        }
        if (b < nblk) s0 += __ldg(partial + (size_t)b * per + e);
    }
    red[warp][lane] = s0 + s1;
    __syncthreads();
    if (warp == 0 && e < per) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][lane];
        if (e < nw) dw[e] = s;
        else if (db) db[e - nw] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------------------------------ host
static int odd_chunks(int q) { return (q & 1) ? q : q + 1; }

struct PwGemmPlan {
    bool ok = false;
    int K = 0, N = 0, KQ = 0, NQ = 0, chunks = 0, SI = 0, SO = 0, vpt = 0, alias = 0, res_smem = 0;
    size_t smem = 0;
};

static PwGemmPlan pw_gemm_plan(int K, int N, bool with_res = true) {
    PwGemmPlan pl;
    pl.K = K; pl.N = N;
    pl.KQ = ceil_div(K, 4);
    const int nq = ceil_div(N, 4);
    pl.chunks = ceil_div(nq, kPwMaxQ);
    pl.NQ = ceil_div(nq, pl.chunks);
    pl.SI = odd_chunks(pl.KQ);
    pl.SO = odd_chunks(pl.NQ * pl.chunks);
    pl.alias = (pl.chunks == 1 && pl.SO <= pl.SI) ? 1 : 0;
    const size_t fixed = ((size_t)pl.KQ * 4 + 1) * pl.NQ * pl.chunks * 16;
    // preference: 2 voxels per thread with >= 3 CTAs per SM, else 1 voxel per thread, else 1 voxel and the residual rows read from global
    // memory in the epilogue instead of being staged (wide outputs on small maps)
    for (int attempt = 0; attempt < 3 && !pl.ok; ++attempt) {
        const int vpt = attempt == 0 ? 2 : 1;
        const bool stage_res = with_res && attempt < 2;
        const size_t tv = (size_t)kPwThreads * vpt;
        pl.smem = fixed + 2 * tv * pl.SI * 16 + (stage_res ? tv * (size_t)ceil_div(N, 4) * 16 : 0) + (pl.alias ? 0 : tv * (size_t)pl.SO * 16);
        pl.vpt = vpt;
        pl.res_smem = stage_res ? 1 : 0;
        pl.ok = pl.smem <= (size_t)(vpt == 2 ? 110 : 200) * 1024;   // 2 voxels per thread as long as 2 CTAs fit an SM
    }
    return pl;
}

static bool pw_geometry(const ConvGeom &g) {
    return g.kd == 1 && g.kh == 1 && g.kw == 1 && g.sd == 1 && g.sh == 1 && g.sw == 1 && g.pd == 0 && g.ph == 0 && g.pw == 0 &&
           (long long)g.cin * g.cout <= kPwMaxMacs;
}

static bool pw_wgrad_ok(const ConvGeom &g) {
    const int tiles4 = ceil_div(g.cin, 4) * ceil_div(g.cout, 4);
    return tiles4 <= kPwThreads * kPwWgMaxTpt;
}

bool conv_pw_preferred(const ConvGeom &g) { return (long long)g.cin * g.cout <= kPwAutoMacs; }

bool conv_pw_supported(const ConvGeom &g, int pass) {
    if (!pw_geometry(g)) return false;
    if (pass == 2) return pw_wgrad_ok(g);
    return pass == 0 ? pw_gemm_plan(g.cin, g.cout).ok : pw_gemm_plan(g.cout, g.cin).ok;
}

template <int VPT, int NQ>
static int pw_launch(const PwGemmParams &p, const PwGemmPlan &pl, cudaStream_t st) {
    static bool attr[kMaxDevices] = {};
    if (!ensure_smem_attr(pw_gemm_kernel<VPT, NQ>, 200 * 1024, attr)) return MDT_EDRIVER;
    const long long per_sm = std::max<long long>(1, std::min<long long>(8, (220 * 1024) / (long long)(pl.smem + 1024)));
    const unsigned grid = (unsigned)std::min<long long>(p.tiles, (long long)num_sms() * per_sm);
    pw_gemm_kernel<VPT, NQ><<<grid, kPwThreads, pl.smem, st>>>(p);
    return launch_status();
}

template <int VPT>
static int pw_dispatch_q(const PwGemmParams &p, const PwGemmPlan &pl, cudaStream_t st) {
    switch (pl.NQ) {
        case 1: return pw_launch<VPT, 1>(p, pl, st);
        case 2: return pw_launch<VPT, 2>(p, pl, st);
        case 3: return pw_launch<VPT, 3>(p, pl, st);
        case 4: return pw_launch<VPT, 4>(p, pl, st);
        case 5: return pw_launch<VPT, 5>(p, pl, st);
        case 6: return pw_launch<VPT, 6>(p, pl, st);
        case 7: return pw_launch<VPT, 7>(p, pl, st);
        case 8: return pw_launch<VPT, 8>(p, pl, st);
        case 9: return pw_launch<VPT, 9>(p, pl, st);
        case 10: return pw_launch<VPT, 10>(p, pl, st);
        case 11: return pw_launch<VPT, 11>(p, pl, st);
        case 12: return pw_launch<VPT, 12>(p, pl, st);
    }
    return MDT_EUNSUPPORTED;
}

static bool aligned8(const void *q) { return (reinterpret_cast<uintptr_t>(q) & 7) == 0; }

static int pw_gemm_run(const ConvGeom &g, int K, int N, const float *in, const float *w, int w_kn, const float *bias, const float *res,
                       const float *relu_of, float *out, float *masked_out, int relu, int precision, void *out_split, cudaStream_t st) {
    const PwGemmPlan pl = pw_gemm_plan(K, N, res != nullptr);
    if (!pl.ok) return MDT_EUNSUPPORTED;
    PwGemmParams p{};
    p.in = in; p.w = w; p.bias = bias; p.res = res; p.relu_of = relu_of; p.out = out; p.masked_out = masked_out;
    p.out_split = reinterpret_cast<__nv_bfloat16 *>(out_split);
    p.V = (long long)g.n * g.d * g.h * g.w;
    p.tiles = ceil_div<long long>(p.V, (long long)kPwThreads * pl.vpt);
    p.K = K; p.N = N; p.KQ = pl.KQ; p.SI = pl.SI; p.SO = pl.SO; p.chunks = pl.chunks; p.NP4 = pl.NQ * pl.chunks;
    p.w_kn = w_kn; p.relu = relu; p.line_w = g.w; p.out_kg = conv_tc_kpad(N); p.planes = precision == 1 ? 1 : 2;
    p.alias = pl.alias; p.rs_stride = ceil_div(N, 4) * 4; p.res_smem = pl.res_smem;
    p.vec_in = (K % 2 == 0) && aligned8(in) && aligned8(relu_of) && aligned8(masked_out);
    p.vec_out = (N % 2 == 0) && aligned8(out) && aligned8(res);
    return pl.vpt == 2 ? pw_dispatch_q<2>(p, pl, st) : pw_dispatch_q<1>(p, pl, st);
}

int conv_pw_fprop(const ConvGeom &g, const float *x, const float *w, const float *bias, const float *residual, float *y, int relu, int precision,
                  void *y_split, cudaStream_t st) {
    if (!conv_pw_supported(g, 0)) return MDT_EUNSUPPORTED;
    return pw_gemm_run(g, g.cin, g.cout, x, w, 0, bias, residual, nullptr, y, nullptr, relu, precision, y_split, st);
}

int conv_pw_dgrad(const ConvGeom &g, const float *dy, const float *relu_of, const float *w, float *dx, float *dy_masked_out, cudaStream_t st) {
    if (!conv_pw_supported(g, 1)) return MDT_EUNSUPPORTED;
    return pw_gemm_run(g, g.cout, g.cin, dy, w, 1, nullptr, nullptr, relu_of, dx, dy_masked_out, 0, 0, nullptr, st);
}

static size_t pw_wgrad_smem(const ConvGeom &g) {
    const int IQ = ceil_div(g.cin, 4), CQ = ceil_div(g.cout, 4);
    const int tiles4 = IQ * CQ, tpt = ceil_div(tiles4, kPwThreads), TG = ceil_div(tiles4, tpt), G = std::max(1, kPwThreads / TG);
    const size_t tile_bytes = 2 * (size_t)kPwWgTile * (IQ + CQ) * 16;   // double-buffered x and dy rows
    const size_t red_bytes = (size_t)G * ((size_t)CQ * 4 * IQ * 4 + CQ * 4) * 4;
    return std::max(tile_bytes, red_bytes);
}

static unsigned pw_wgrad_grid(const ConvGeom &g) {
    const long long V = (long long)g.n * g.d * g.h * g.w;
    const long long per_sm = std::max<long long>(1, std::min<long long>(4, (220 * 1024) / (long long)(pw_wgrad_smem(g) + 1024)));
    return (unsigned)std::min<long long>(ceil_div<long long>(V, kPwWgTile), (long long)num_sms() * per_sm);
}

size_t conv_pw_workspace_bytes(const ConvGeom &g, int pass) {
    if (pass != 2) return 0;
    return (size_t)pw_wgrad_grid(g) * ((size_t)g.cout * g.cin + g.cout) * sizeof(float);
}

int conv_pw_wgrad(const ConvGeom &g, const float *x, const float *dy, const float *relu_of, float *dw, float *db, void *ws, size_t ws_bytes,
                  cudaStream_t st, float *dy_masked_out) {
    if (!conv_pw_supported(g, 2)) return MDT_EUNSUPPORTED;
    if (!ws || ws_bytes < conv_pw_workspace_bytes(g, 2)) return MDT_EWORKSPACE;
    PwWgradParams p{};
    p.x = x; p.dy = dy; p.relu_of = relu_of; p.masked_out = dy_masked_out;
    p.partial = reinterpret_cast<float *>(ws);
    p.V = (long long)g.n * g.d * g.h * g.w;
    p.tiles = ceil_div<long long>(p.V, kPwWgTile);
    p.Cin = g.cin; p.Cout = g.cout; p.IQ = ceil_div(g.cin, 4); p.CQ = ceil_div(g.cout, 4);
    const int tiles4 = p.IQ * p.CQ;
    const int tpt = ceil_div(tiles4, kPwThreads);
    p.TG = ceil_div(tiles4, tpt);
    p.G = max(1, kPwThreads / p.TG);
    p.vec_x = (g.cin % 2 == 0) && aligned8(x);
    p.vec_y = (g.cout % 2 == 0) && aligned8(dy) && aligned8(relu_of) && aligned8(dy_masked_out);
    const size_t smem = pw_wgrad_smem(g);
    if (smem > 200 * 1024) return MDT_EUNSUPPORTED;
    const unsigned grid = pw_wgrad_grid(g);
    int rc;
    static bool a1[kMaxDevices] = {}, a2[kMaxDevices] = {}, a3[kMaxDevices] = {};
    if (tpt == 1) {
        if (!ensure_smem_attr(pw_wgrad_kernel<1>, 200 * 1024, a1)) return MDT_EDRIVER;
        pw_wgrad_kernel<1><<<grid, kPwThreads, smem, st>>>(p);
    } else if (tpt == 2) {
        if (!ensure_smem_attr(pw_wgrad_kernel<2>, 200 * 1024, a2)) return MDT_EDRIVER;
        pw_wgrad_kernel<2><<<grid, kPwThreads, smem, st>>>(p);
    } else {
        if (!ensure_smem_attr(pw_wgrad_kernel<3>, 200 * 1024, a3)) return MDT_EDRIVER;
        pw_wgrad_kernel<3><<<grid, kPwThreads, smem, st>>>(p);
    }
    if ((rc = launch_status())) return rc;
    const int nw = g.cout * g.cin, nb = g.cout;
    pw_wgrad_reduce_kernel<<<ceil_div(nw + nb, 32), 256, 0, st>>>(p.partial, (int)grid, nw, nb, dw, db);
    return launch_status();
}

}  // namespace mdt
