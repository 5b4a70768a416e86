// RoIAlign (TF-style crop_and_resize, one tri-/bi-linear sample per output bin) forward + backward for sm_100a, 2D and 3D.
//
// Reference semantics followed (paths relative to the reference root):
//   sampling coordinates  cuda_functions/roi_align_3D/roi_align/src/cuda/crop_and_resize_kernel.cu:52-106 ("fixed" half-bin-centred, clamped to the map)
//   neighbours / lerp     :110-147   forward interpolation order (x, then y, then z)
//   backward weights      :256-301   8 (2D: 4) scatter-adds into a zero-initialised image gradient
//   bad box_ind           :43-47 + crop_and_resize_gpu.c:27: that crop is all zeros
//   2D twin               cuda_functions/roi_align_2D/roi_align/src/cuda/crop_and_resize_kernel.cu:10-194
//
// B200 design: the reference runs one thread per output scalar over an NCDHW map, so the 8 gathers of a warp hit 8x32 different
// sectors. Here the fast path works on channels-last maps (the layout the tcgen05 conv kernels produce): one thread owns one bin x 4
// channels, so every corner read / gradient scatter of a warp is a run of consecutive 16-byte vectors (fully coalesced 128-bit
// ld.global.nc / red.global.add.v4.f32). A stride-generic scalar kernel keeps the reference's NCDHW contract working with no copy.
#include "mdt_common.cuh"

namespace mdt {

struct RoiGeom {
    int num_boxes, batch, H, W, Z, ch, cw, cz, C;
    int64_t is[5];  // image strides  (b, c, y, x, z) in elements
    int64_t os[5];  // crop strides   (n, c, y, x, z) in elements
};

// sampling coordinate of output index `o` along one axis (crop extent `crop`, map extent `size`); expression form kept as in the
// reference so that default nvcc contraction produces the same roundings
__device__ __forceinline__ float sample_coord(float lo, float hi, int o, int crop, int size) {
    const float scale = (crop > 1) ? (hi - lo) * (size) / (crop) : 0;
    float t = (crop > 1) ? lo * (size) + o * scale + scale / 2 - 0.5 : 0.5 * (lo + hi) * (size);
    if (t > size - 1) t = size - 1;
    if (t < 0) t = 0;
    return t;
}

struct Tap {
    int lo, hi;
    float lerp;
};
__device__ __forceinline__ Tap make_tap(float in) {
    Tap t;
    t.lo = (int)floorf(in);
    t.hi = (int)ceilf(in);
    t.lerp = in - t.lo;
    return t;
}

template <int DIM>
__device__ __forceinline__ bool bin_taps(const RoiGeom &g, const float *__restrict__ boxes, const int *__restrict__ box_ind, int n, int y, int x,
                                         int z, Tap &ty, Tap &tx, Tap &tz, int &b_in) {
    b_in = box_ind[n];
    if (b_in < 0 || b_in >= g.batch) return false;
    const float *bx = boxes + (size_t)n * (2 * DIM);
    ty = make_tap(sample_coord(bx[0], bx[2], y, g.ch, g.H));
    tx = make_tap(sample_coord(bx[1], bx[3], x, g.cw, g.W));
    if (DIM == 3) tz = make_tap(sample_coord(bx[4], bx[5], z, g.cz, g.Z));
    else { tz.lo = tz.hi = 0; tz.lerp = 0.f; }
    return true;
}

// ---------------------------------------------------------------- generic strided scalar kernels ----------------------------------------------------------------
// IdxT: the flat thread->element index is decomposed with five divisions; 64-bit integer division is emulated (~100 instructions each)
// and dominated the float4 forward kernel, so launches below 2^31 elements use unsigned 32-bit indices.
// thread -> one output scalar. c_fastest selects the thread->element order so that consecutive lanes touch consecutive addresses of the
// dominant stream (channels for channels-last maps, z for the reference's NCDHW maps).
template <int DIM, typename IdxT>
__global__ void __launch_bounds__(256) roi_fwd_scalar(RoiGeom g, const float *__restrict__ image, const float *__restrict__ boxes,
                                                     const int *__restrict__ box_ind, float *__restrict__ crops, int c_fastest, long long total) {
    for (IdxT t = blockIdx.x * (IdxT)blockDim.x + threadIdx.x; t < (IdxT)total; t += (IdxT)gridDim.x * blockDim.x) {
        IdxT r = t;
        int c, z, x, y;
        if (c_fastest) { c = r % g.C; r /= g.C; z = r % g.cz; r /= g.cz; x = r % g.cw; r /= g.cw; y = r % g.ch; r /= g.ch; }
        else           { z = r % g.cz; r /= g.cz; x = r % g.cw; r /= g.cw; y = r % g.ch; r /= g.ch; c = r % g.C; r /= g.C; }
        const int n = (int)r;
        float *out = crops + n * g.os[0] + c * g.os[1] + y * g.os[2] + x * g.os[3] + z * g.os[4];
        Tap ty, tx, tz; int b;
        if (!bin_taps<DIM>(g, boxes, box_ind, n, y, x, z, ty, tx, tz, b)) { *out = 0.f; continue; }
        const float *p = image + b * g.is[0] + c * g.is[1];
        float v[2];
#pragma unroll
        for (int kz = 0; kz < (DIM == 3 ? 2 : 1); ++kz) {
            const int64_t oz = (DIM == 3) ? (kz ? tz.hi : tz.lo) * g.is[4] : 0;
            const float tl = __ldg(p + ty.lo * g.is[2] + tx.lo * g.is[3] + oz), tr = __ldg(p + ty.lo * g.is[2] + tx.hi * g.is[3] + oz);
            const float bl = __ldg(p + ty.hi * g.is[2] + tx.lo * g.is[3] + oz), br = __ldg(p + ty.hi * g.is[2] + tx.hi * g.is[3] + oz);
            const float top = tl + (tr - tl) * tx.lerp, bot = bl + (br - bl) * tx.lerp;
            v[kz] = top + (bot - top) * ty.lerp;
        }
        *out = (DIM == 3) ? v[0] + (v[1] - v[0]) * tz.lerp : v[0];
    }
}

template <int DIM, typename IdxT>
__global__ void __launch_bounds__(256) roi_bwd_scalar(RoiGeom g, const float *__restrict__ grads, const float *__restrict__ boxes,
                                                     const int *__restrict__ box_ind, float *__restrict__ gimg, int c_fastest, long long total) {
    for (IdxT t = blockIdx.x * (IdxT)blockDim.x + threadIdx.x; t < (IdxT)total; t += (IdxT)gridDim.x * blockDim.x) {
        IdxT r = t;
        int c, z, x, y;
        if (c_fastest) { c = r % g.C; r /= g.C; z = r % g.cz; r /= g.cz; x = r % g.cw; r /= g.cw; y = r % g.ch; r /= g.ch; }
        else           { z = r % g.cz; r /= g.cz; x = r % g.cw; r /= g.cw; y = r % g.ch; r /= g.ch; c = r % g.C; r /= g.C; }
        const int n = (int)r;
        Tap ty, tx, tz; int b;
        if (!bin_taps<DIM>(g, boxes, box_ind, n, y, x, z, ty, tx, tz, b)) continue;
        const float gv = grads[n * g.os[0] + c * g.os[1] + y * g.os[2] + x * g.os[3] + z * g.os[4]];
        float *p = gimg + b * g.is[0] + c * g.is[1];
#pragma unroll
        for (int kz = 0; kz < (DIM == 3 ? 2 : 1); ++kz) {
            const float wz = (DIM == 3) ? (kz ? tz.lerp : 1 - tz.lerp) : 1.f;
            const int64_t oz = (DIM == 3) ? (kz ? tz.hi : tz.lo) * g.is[4] : 0;
#pragma unroll
            for (int ky = 0; ky < 2; ++ky) {
                const float wy = ky ? ty.lerp : 1 - ty.lerp;
                const int64_t oy = (ky ? ty.hi : ty.lo) * g.is[2];
#pragma unroll
                for (int kx = 0; kx < 2; ++kx) {
                    const float wx = kx ? tx.lerp : 1 - tx.lerp;
                    // weight product order of the reference: x * z * y * grad
                    const float w = (DIM == 3) ? wx * wz * wy * gv : wx * wy * gv;
                    if (w != 0.f) atomicAdd(p + oy + (kx ? tx.hi : tx.lo) * g.is[3] + oz, w);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- channels-last float4 kernels ----------------------------------------------------------------
// requires is[1] == 1, os[1] == 1, C % 4 == 0, all other strides % 4 == 0 and 16-byte aligned bases.
template <int DIM, typename IdxT>
__global__ void __launch_bounds__(256) roi_fwd_cl4(RoiGeom g, const float *__restrict__ image, const float *__restrict__ boxes,
                                                  const int *__restrict__ box_ind, float *__restrict__ crops, long long total) {
    const int C4 = g.C >> 2;
    for (IdxT t = blockIdx.x * (IdxT)blockDim.x + threadIdx.x; t < (IdxT)total; t += (IdxT)gridDim.x * blockDim.x) {
        IdxT r = t;
        const int c4 = r % C4; r /= C4;
        const int z = r % g.cz; r /= g.cz;
        const int x = r % g.cw; r /= g.cw;
        const int y = r % g.ch; r /= g.ch;
        const int n = (int)r;
        float4 *out = reinterpret_cast<float4 *>(crops + n * g.os[0] + y * g.os[2] + x * g.os[3] + z * g.os[4]) + c4;
        Tap ty, tx, tz; int b;
        if (!bin_taps<DIM>(g, boxes, box_ind, n, y, x, z, ty, tx, tz, b)) { *out = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
        const float *p = image + b * g.is[0] + 4 * c4;
        float4 v[2];
#pragma unroll
        for (int kz = 0; kz < (DIM == 3 ? 2 : 1); ++kz) {
            const int64_t oz = (DIM == 3) ? (kz ? tz.hi : tz.lo) * g.is[4] : 0;
            const float4 tl = __ldg(reinterpret_cast<const float4 *>(p + ty.lo * g.is[2] + tx.lo * g.is[3] + oz));
            const float4 tr = __ldg(reinterpret_cast<const float4 *>(p + ty.lo * g.is[2] + tx.hi * g.is[3] + oz));
            const float4 bl = __ldg(reinterpret_cast<const float4 *>(p + ty.hi * g.is[2] + tx.lo * g.is[3] + oz));
            const float4 br = __ldg(reinterpret_cast<const float4 *>(p + ty.hi * g.is[2] + tx.hi * g.is[3] + oz));
#define MDT_LERP2(f) { const float top = tl.f + (tr.f - tl.f) * tx.lerp, bot = bl.f + (br.f - bl.f) * tx.lerp; v[kz].f = top + (bot - top) * ty.lerp; }
            MDT_LERP2(x) MDT_LERP2(y) MDT_LERP2(z) MDT_LERP2(w)
#undef MDT_LERP2
        }
        float4 o = v[0];
        if (DIM == 3) {
            o.x = v[0].x + (v[1].x - v[0].x) * tz.lerp; o.y = v[0].y + (v[1].y - v[0].y) * tz.lerp;
            o.z = v[0].z + (v[1].z - v[0].z) * tz.lerp; o.w = v[0].w + (v[1].w - v[0].w) * tz.lerp;
        }
        *out = o;
    }
}

template <int DIM, typename IdxT>
__global__ void __launch_bounds__(256) roi_bwd_cl4(RoiGeom g, const float *__restrict__ grads, const float *__restrict__ boxes,
                                                  const int *__restrict__ box_ind, float *__restrict__ gimg, long long total) {
    const int C4 = g.C >> 2;
    for (IdxT t = blockIdx.x * (IdxT)blockDim.x + threadIdx.x; t < (IdxT)total; t += (IdxT)gridDim.x * blockDim.x) {
        IdxT r = t;
        const int c4 = r % C4; r /= C4;
        const int z = r % g.cz; r /= g.cz;
        const int x = r % g.cw; r /= g.cw;
        const int y = r % g.ch; r /= g.ch;
        const int n = (int)r;
        Tap ty, tx, tz; int b;
        if (!bin_taps<DIM>(g, boxes, box_ind, n, y, x, z, ty, tx, tz, b)) continue;
        const float4 gv = __ldg(reinterpret_cast<const float4 *>(grads + n * g.os[0] + y * g.os[2] + x * g.os[3] + z * g.os[4]) + c4);
        float *p = gimg + b * g.is[0] + 4 * c4;
#pragma unroll
        for (int kz = 0; kz < (DIM == 3 ? 2 : 1); ++kz) {
            const float wz = (DIM == 3) ? (kz ? tz.lerp : 1 - tz.lerp) : 1.f;
            const int64_t oz = (DIM == 3) ? (kz ? tz.hi : tz.lo) * g.is[4] : 0;
#pragma unroll
            for (int ky = 0; ky < 2; ++ky) {
                const float wy = ky ? ty.lerp : 1 - ty.lerp;
                const int64_t oy = (ky ? ty.hi : ty.lo) * g.is[2];
#pragma unroll
                for (int kx = 0; kx < 2; ++kx) {
                    const float wx = kx ? tx.lerp : 1 - tx.lerp;
                    const float w = (DIM == 3) ? wx * wz * wy : wx * wy;
                    if (w != 0.f)  // one 128-bit reduction (RED.E.ADD.F32x4) instead of four scalar atomics
                        atomicAdd(reinterpret_cast<float4 *>(p + oy + (kx ? tx.hi : tx.lo) * g.is[3] + oz),
                                  make_float4(w * gv.x, w * gv.y, w * gv.z, w * gv.w));
                }
            }
        }
    }
}

// ---------------------------------------------------------------- channels-last, one RoI (slice) per CTA ----------------------------------------------------------------
// ncu of roi_fwd_cl4 (P2, 7x7x3, 1024 RoIs): 390 instructions per thread, issue slots 54 % busy, DRAM 10 %: every one of the C/4
// threads of a bin repeats the index decomposition (five integer divisions), three sampling coordinates (a float division each) and
// the 64-bit address arithmetic; staging those per group of bins only moved the cost into a serial chain per group (24 us for the
// 38 MB P2 map and for the 4.7 MB P3 map alike).  But the taps of an RoI are SEPARABLE: ch + cw + cz values (17 for 7x7x3), not
// 3 per bin.  A CTA therefore owns one RoI (or a slice of its bins when the crop is large): ch+cw+cz threads compute one tap each
// (neighbour offsets + lerp weight; box and box_ind loads are independent of each other), one pass fills a per-bin table (packed tap
// indices + output offset), and then all 256 threads stream (bin, 4-channel) items: table lookups, eight 128-bit gathers, lerp, one
// 128-bit store - no divisions and no barrier in the loop.  Expressions are those of roi_fwd_cl4 / roi_bwd_cl4: bit-identical results.
constexpr int kRoiThreads = 256;
constexpr int kMaxCropDim = 64;       // per axis, for the shared-memory tap tables
constexpr int kMaxBinsPerCta = 1024;  // bins per CTA slice (12 KB of tables)

struct RoiTables {
    int lo[3][kMaxCropDim], hi[3][kMaxCropDim];   // element offsets (tap index x axis stride) inside one batch item
    float lerp[3][kMaxCropDim];
    long long out[kMaxBinsPerCta];                // element offset of the bin's channel vector in the crops tensor
    int yxz[kMaxBinsPerCta];                      // y | x << 8 | z << 16
};

template <int DIM, bool BWD>
__device__ __forceinline__ void roi_cl4_per_roi_body(const RoiGeom &g, const float *__restrict__ src, const float *__restrict__ boxes,
                                                     const int *__restrict__ box_ind, float *__restrict__ dst, int bins_per_slice, RoiTables &s) {
    const int n = blockIdx.x, tid = threadIdx.x;
    const int P = g.ch * g.cw * g.cz;
    const int bin0 = blockIdx.y * bins_per_slice, nbins = min(P - bin0, bins_per_slice);
    const int b_in = box_ind[n];
    const bool valid = b_in >= 0 && b_in < g.batch;
    if (valid && tid < g.ch + g.cw + (DIM == 3 ? g.cz : 0)) {   // one tap per thread
        const float *bx = boxes + (size_t)n * (2 * DIM);
        Tap tp; int64_t stride; int axis, o;
        if (tid < g.ch)             { axis = 0; o = tid;               tp = make_tap(sample_coord(bx[0], bx[2], o, g.ch, g.H)); stride = g.is[2]; }
        else if (tid < g.ch + g.cw) { axis = 1; o = tid - g.ch;        tp = make_tap(sample_coord(bx[1], bx[3], o, g.cw, g.W)); stride = g.is[3]; }
        else                        { axis = 2; o = tid - g.ch - g.cw; tp = make_tap(sample_coord(bx[4], bx[5], o, g.cz, g.Z)); stride = g.is[4]; }
        s.lo[axis][o] = (int)(tp.lo * stride);
        s.hi[axis][o] = (int)(tp.hi * stride);
        s.lerp[axis][o] = tp.lerp;
    }
    for (int i = tid; i < nbins; i += kRoiThreads) {
        unsigned r = bin0 + i;
        const int z = r % (unsigned)g.cz; r /= (unsigned)g.cz;
        const int x = r % (unsigned)g.cw; r /= (unsigned)g.cw;
        const int y = (int)r;
        s.yxz[i] = y | (x << 8) | (z << 16);
        s.out[i] = n * g.os[0] + y * g.os[2] + x * g.os[3] + z * g.os[4];
    }
    __syncthreads();
    const int C4 = g.C >> 2;
    const int items = nbins * C4;
    int bl = tid / C4, c4 = tid - bl * C4;                        // item -> (bin, channel group), advanced without divisions
    const int dq = kRoiThreads / C4, dr = kRoiThreads - dq * C4;
    const float *img = BWD ? nullptr : src + (valid ? b_in : 0) * g.is[0];
    float *gim = BWD ? dst + (valid ? b_in : 0) * g.is[0] : nullptr;
    for (int item = tid; item < items; item += kRoiThreads) {
        const long long out_off = s.out[bl];
        if (!valid) {
            if (!BWD) *(reinterpret_cast<float4 *>(dst + out_off) + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const int pk = s.yxz[bl];
            const int y = pk & 255, x = (pk >> 8) & 255, z = pk >> 16;
            const int ylo = s.lo[0][y], yhi = s.hi[0][y], xlo = s.lo[1][x], xhi = s.hi[1][x];
            const float ly = s.lerp[0][y], lx = s.lerp[1][x], lz = (DIM == 3) ? s.lerp[2][z] : 0.f;
            if (!BWD) {
                const float *p = img + 4 * c4;
                float4 v[2];
#pragma unroll
                for (int kz = 0; kz < (DIM == 3 ? 2 : 1); ++kz) {
                    const int oz = (DIM == 3) ? (kz ? s.hi[2][z] : s.lo[2][z]) : 0;
                    const float4 tl = __ldg(reinterpret_cast<const float4 *>(p + ylo + xlo + oz));
                    const float4 tr = __ldg(reinterpret_cast<const float4 *>(p + ylo + xhi + oz));
                    const float4 bl_ = __ldg(reinterpret_cast<const float4 *>(p + yhi + xlo + oz));
                    const float4 br = __ldg(reinterpret_cast<const float4 *>(p + yhi + xhi + oz));
#define MDT_LERP2(f) { const float top = tl.f + (tr.f - tl.f) * lx, bot = bl_.f + (br.f - bl_.f) * lx; v[kz].f = top + (bot - top) * ly; }
                    MDT_LERP2(x) MDT_LERP2(y) MDT_LERP2(z) MDT_LERP2(w)
#undef MDT_LERP2
                }
                float4 o = v[0];
                if (DIM == 3) {
                    o.x = v[0].x + (v[1].x - v[0].x) * lz; o.y = v[0].y + (v[1].y - v[0].y) * lz;
                    o.z = v[0].z + (v[1].z - v[0].z) * lz; o.w = v[0].w + (v[1].w - v[0].w) * lz;
                }
                *(reinterpret_cast<float4 *>(dst + out_off) + c4) = o;
            } else {
                const float4 gv = __ldg(reinterpret_cast<const float4 *>(src + out_off) + c4);
                float *p = gim + 4 * c4;
#pragma unroll
                for (int kz = 0; kz < (DIM == 3 ? 2 : 1); ++kz) {
                    const float wz = (DIM == 3) ? (kz ? lz : 1 - lz) : 1.f;
                    const int oz = (DIM == 3) ? (kz ? s.hi[2][z] : s.lo[2][z]) : 0;
#pragma unroll
                    for (int ky = 0; ky < 2; ++ky) {
                        const float wy = ky ? ly : 1 - ly;
                        const int oy = ky ? yhi : ylo;
#pragma unroll
                        for (int kx = 0; kx < 2; ++kx) {
                            const float wx = kx ? lx : 1 - lx;
                            const float w = (DIM == 3) ? wx * wz * wy : wx * wy;
                            if (w != 0.f)  // one 128-bit reduction (RED.E.ADD.F32x4) instead of four scalar atomics
                                atomicAdd(reinterpret_cast<float4 *>(p + oy + (kx ? xhi : xlo) + oz), make_float4(w * gv.x, w * gv.y, w * gv.z, w * gv.w));
                        }
                    }
                }
            }
        }
        bl += dq; c4 += dr;
        if (c4 >= C4) { c4 -= C4; ++bl; }
    }
}

template <int DIM, bool BWD>
__global__ void __launch_bounds__(kRoiThreads) roi_cl4_per_roi(RoiGeom g, const float *__restrict__ src, const float *__restrict__ boxes,
                                                              const int *__restrict__ box_ind, float *__restrict__ dst, int bins_per_slice) {
    __shared__ RoiTables s;
    roi_cl4_per_roi_body<DIM, BWD>(g, src, boxes, box_ind, dst, bins_per_slice, s);
}

// ---- pyramid variant: ONE launch for all levels of an FPN (replaces the per-level loop of models/mrcnn.py:405-447 — boolean-mask gather,
// <= 4 crop_and_resize calls, concat, un-permute by sort).  Every RoI carries its level; its CTA reads that level's map geometry, so each
// output row is written exactly once and, in the backward pass, each level's gradient map receives only its own RoIs.
constexpr int kMaxPyrLevels = 5;
struct PyrGeom {
    int num_boxes, batch, ch, cw, cz, C, nlevels;
    int64_t os[5];
    int H[kMaxPyrLevels], W[kMaxPyrLevels], Z[kMaxPyrLevels];
    int64_t is[kMaxPyrLevels][5];
    float *map[kMaxPyrLevels];        // forward: source maps (read only); backward: gradient maps
};

template <int DIM, bool BWD>
__global__ void __launch_bounds__(kRoiThreads) roi_cl4_pyramid(const __grid_constant__ PyrGeom pg, const float *__restrict__ crops_or_grads,
                                                              const float *__restrict__ boxes, const int *__restrict__ box_ind,
                                                              const int *__restrict__ roi_level, float *__restrict__ crops_out, int bins_per_slice) {
    __shared__ RoiTables s;
    int lv = roi_level[blockIdx.x];
    lv = lv < 0 ? 0 : (lv >= pg.nlevels ? pg.nlevels - 1 : lv);
    RoiGeom g;
    g.num_boxes = pg.num_boxes; g.batch = pg.batch; g.H = pg.H[lv]; g.W = pg.W[lv]; g.Z = pg.Z[lv];
    g.ch = pg.ch; g.cw = pg.cw; g.cz = pg.cz; g.C = pg.C;
#pragma unroll
    for (int k = 0; k < 5; ++k) { g.is[k] = pg.is[lv][k]; g.os[k] = pg.os[k]; }
    if (!BWD) roi_cl4_per_roi_body<DIM, false>(g, pg.map[lv], boxes, box_ind, crops_out, bins_per_slice, s);
    else      roi_cl4_per_roi_body<DIM, true>(g, crops_or_grads, boxes, box_ind, pg.map[lv], bins_per_slice, s);
}

// the per-RoI kernels keep tap offsets in 32 bits, tap tables of 64 entries per axis and <= 256 channel groups
static bool per_roi_ok(const RoiGeom &g) {
    if ((g.C >> 2) > kRoiThreads || g.ch > kMaxCropDim || g.cw > kMaxCropDim || g.cz > kMaxCropDim) return false;
    const long long span = (long long)g.H * g.is[2] + (long long)g.W * g.is[3] + (long long)g.Z * g.is[4];
    return span < (1LL << 31);
}

// slices per RoI: ~1500 (bin, 4-channel) items per CTA keeps the table pass short and gives the scheduler several CTAs per SM
static void per_roi_grid(const RoiGeom &g, dim3 &grid, int &bins_per_slice) {
    const int P = g.ch * g.cw * g.cz, C4 = g.C >> 2;
    int bins = 1536 / C4;
    if (bins < 1) bins = 1;
    if (bins > kMaxBinsPerCta) bins = kMaxBinsPerCta;
    if (bins > P) bins = P;
    bins_per_slice = bins;
    grid = dim3((unsigned)g.num_boxes, (unsigned)ceil_div(P, bins), 1);
}

static bool cl4_ok(const RoiGeom &g, const void *img, const void *crop) {
    if (g.C % 4 || g.is[1] != 1 || g.os[1] != 1) return false;
    for (int k : {0, 2, 3, 4})
        if (g.is[k] % 4 || g.os[k] % 4) return false;
    return ((uintptr_t)img % 16 == 0) && ((uintptr_t)crop % 16 == 0);
}

// grid-stride loops advance by gridDim*blockDim <= 148*32*256 ~ 1.2 M, so t + stride cannot wrap below 2^31 elements
static bool fits_u32(long long total) { return total < (1LL << 31); }

static int grid_for(long long total, int block) {
    long long need = (total + block - 1) / block;
    long long cap = (long long)num_sms() * 32;  // 8 resident CTAs of 256 threads per SM x 4 waves; grid-stride covers the rest
    return (int)(need < cap ? need : cap);
}

template <int DIM>
static int roi_forward(const float *image, const int64_t *is, const float *boxes, const int *box_ind, int num_boxes, int batch, int H, int W, int Z,
                       int ch, int cw, int cz, int C, float *crops, const int64_t *os, cudaStream_t st) {
    if (num_boxes < 0 || batch <= 0 || H <= 0 || W <= 0 || Z <= 0 || ch <= 0 || cw <= 0 || cz <= 0 || C <= 0 || !is || !os) return MDT_EINVAL;
    if (num_boxes == 0) return MDT_OK;
    if (!image || !boxes || !box_ind || !crops) return MDT_EINVAL;
    RoiGeom g{num_boxes, batch, H, W, Z, ch, cw, cz, C, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
    for (int k = 0; k < DIM + 2; ++k) { g.is[k] = is[k]; g.os[k] = os[k]; }
    const long long total = (long long)num_boxes * C * ch * cw * cz;
    const bool small = fits_u32(total);
    if (cl4_ok(g, image, crops) && per_roi_ok(g)) {
        dim3 grid; int bins;
        per_roi_grid(g, grid, bins);
        roi_cl4_per_roi<DIM, false><<<grid, kRoiThreads, 0, st>>>(g, image, boxes, box_ind, crops, bins);
    } else if (cl4_ok(g, image, crops)) {
        if (small) roi_fwd_cl4<DIM, unsigned><<<grid_for(total / 4, 256), 256, 0, st>>>(g, image, boxes, box_ind, crops, total / 4);
        else       roi_fwd_cl4<DIM, long long><<<grid_for(total / 4, 256), 256, 0, st>>>(g, image, boxes, box_ind, crops, total / 4);
    } else {
        const int c_fastest = (g.is[1] == 1);
        if (small) roi_fwd_scalar<DIM, unsigned><<<grid_for(total, 256), 256, 0, st>>>(g, image, boxes, box_ind, crops, c_fastest, total);
        else       roi_fwd_scalar<DIM, long long><<<grid_for(total, 256), 256, 0, st>>>(g, image, boxes, box_ind, crops, c_fastest, total);
    }
    return launch_status();
}

template <int DIM>
static int roi_backward(const float *grads, const int64_t *gs, const float *boxes, const int *box_ind, int num_boxes, int batch, int H, int W, int Z,
                        int ch, int cw, int cz, int C, float *gimg, const int64_t *is, int zero_init, int64_t image_numel, cudaStream_t st) {
    if (num_boxes < 0 || batch <= 0 || H <= 0 || W <= 0 || Z <= 0 || ch <= 0 || cw <= 0 || cz <= 0 || C <= 0 || !is || !gs || !gimg) return MDT_EINVAL;
    if (zero_init) {
        if (image_numel <= 0) return MDT_EINVAL;
        cudaError_t e = cudaMemsetAsync(gimg, 0, (size_t)image_numel * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    if (num_boxes == 0) return MDT_OK;
    if (!grads || !boxes || !box_ind) return MDT_EINVAL;
    RoiGeom g{num_boxes, batch, H, W, Z, ch, cw, cz, C, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
    for (int k = 0; k < DIM + 2; ++k) { g.is[k] = is[k]; g.os[k] = gs[k]; }
    const long long total = (long long)num_boxes * C * ch * cw * cz;
    const bool small = fits_u32(total);
    if (cl4_ok(g, gimg, grads) && per_roi_ok(g)) {
        dim3 grid; int bins;
        per_roi_grid(g, grid, bins);
        roi_cl4_per_roi<DIM, true><<<grid, kRoiThreads, 0, st>>>(g, grads, boxes, box_ind, gimg, bins);
    } else if (cl4_ok(g, gimg, grads)) {
        if (small) roi_bwd_cl4<DIM, unsigned><<<grid_for(total / 4, 256), 256, 0, st>>>(g, grads, boxes, box_ind, gimg, total / 4);
        else       roi_bwd_cl4<DIM, long long><<<grid_for(total / 4, 256), 256, 0, st>>>(g, grads, boxes, box_ind, gimg, total / 4);
    } else {
        const int c_fastest = (g.is[1] == 1);
        if (small) roi_bwd_scalar<DIM, unsigned><<<grid_for(total, 256), 256, 0, st>>>(g, grads, boxes, box_ind, gimg, c_fastest, total);
        else       roi_bwd_scalar<DIM, long long><<<grid_for(total, 256), 256, 0, st>>>(g, grads, boxes, box_ind, gimg, c_fastest, total);
    }
    return launch_status();
}

template <int DIM>
static int pyramid_run(bool bwd, float *const *maps, const int64_t *is, const int *dims, int nlevels, const float *crops_or_grads, const float *boxes,
                       const int *box_ind, const int *roi_level, int num_boxes, int batch, int ch, int cw, int cz, int C, float *crops_out,
                       const int64_t *os, int zero_init, const int64_t *numel, cudaStream_t st) {
    if (nlevels <= 0 || nlevels > kMaxPyrLevels || num_boxes < 0 || batch <= 0 || ch <= 0 || cw <= 0 || cz <= 0 || C <= 0 || !maps || !is || !dims || !os)
        return MDT_EINVAL;
    PyrGeom pg{};
    pg.num_boxes = num_boxes; pg.batch = batch; pg.ch = ch; pg.cw = cw; pg.cz = cz; pg.C = C; pg.nlevels = nlevels;
    for (int k = 0; k < DIM + 2; ++k) pg.os[k] = os[k];
    const float *crop_ptr = bwd ? crops_or_grads : crops_out;
    for (int l = 0; l < nlevels; ++l) {
        pg.H[l] = dims[3 * l]; pg.W[l] = dims[3 * l + 1]; pg.Z[l] = DIM == 3 ? dims[3 * l + 2] : 1;
        for (int k = 0; k < DIM + 2; ++k) pg.is[l][k] = is[5 * l + k];
        pg.map[l] = maps[l];
        if (!maps[l] || pg.H[l] <= 0 || pg.W[l] <= 0 || pg.Z[l] <= 0) return MDT_EINVAL;
        RoiGeom g{num_boxes, batch, pg.H[l], pg.W[l], pg.Z[l], ch, cw, cz, C, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
        for (int k = 0; k < 5; ++k) { g.is[k] = pg.is[l][k]; g.os[k] = pg.os[k]; }
        if (!cl4_ok(g, maps[l], crop_ptr) || !per_roi_ok(g)) return MDT_EUNSUPPORTED;   // caller falls back to one launch per level
        if (bwd && zero_init) {
            if (!numel || numel[l] <= 0) return MDT_EINVAL;
            cudaError_t e = cudaMemsetAsync(maps[l], 0, (size_t)numel[l] * sizeof(float), st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    if (num_boxes == 0) return MDT_OK;
    if (!boxes || !box_ind || !roi_level || !crop_ptr) return MDT_EINVAL;
    RoiGeom g0{num_boxes, batch, pg.H[0], pg.W[0], pg.Z[0], ch, cw, cz, C, {0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
    dim3 grid; int bins;
    per_roi_grid(g0, grid, bins);
    if (!bwd) roi_cl4_pyramid<DIM, false><<<grid, kRoiThreads, 0, st>>>(pg, nullptr, boxes, box_ind, roi_level, crops_out, bins);
    else      roi_cl4_pyramid<DIM, true><<<grid, kRoiThreads, 0, st>>>(pg, crops_or_grads, boxes, box_ind, roi_level, nullptr, bins);
    return launch_status();
}

}  // namespace mdt

extern "C" {

int mdt_pyramid_roi_align_forward(int dim, const float *const *images, const int64_t *image_strides, const int *image_dims, int nlevels, const float *boxes,
                                  const int *box_ind, const int *roi_level, int num_boxes, int batch, int ch, int cw, int cz, int depth, float *crops,
                                  const int64_t *crop_strides, void *stream) {
    float *const *maps = const_cast<float *const *>(reinterpret_cast<const float *const *>(images));
    if (dim == 3) return mdt::pyramid_run<3>(false, maps, image_strides, image_dims, nlevels, nullptr, boxes, box_ind, roi_level, num_boxes, batch, ch, cw, cz,
                                             depth, crops, crop_strides, 0, nullptr, mdt::as_stream(stream));
    if (dim == 2) return mdt::pyramid_run<2>(false, maps, image_strides, image_dims, nlevels, nullptr, boxes, box_ind, roi_level, num_boxes, batch, ch, cw, 1,
                                             depth, crops, crop_strides, 0, nullptr, mdt::as_stream(stream));
    return MDT_EINVAL;
}
int mdt_pyramid_roi_align_backward(int dim, const float *grads, const int64_t *grad_strides, const float *boxes, const int *box_ind, const int *roi_level,
                                   int num_boxes, int batch, int ch, int cw, int cz, int depth, float *const *grad_images, const int64_t *image_strides,
                                   const int *image_dims, int nlevels, int zero_init, const int64_t *image_numel, void *stream) {
    if (dim == 3) return mdt::pyramid_run<3>(true, grad_images, image_strides, image_dims, nlevels, grads, boxes, box_ind, roi_level, num_boxes, batch, ch, cw,
                                             cz, depth, nullptr, grad_strides, zero_init, image_numel, mdt::as_stream(stream));
    if (dim == 2) return mdt::pyramid_run<2>(true, grad_images, image_strides, image_dims, nlevels, grads, boxes, box_ind, roi_level, num_boxes, batch, ch, cw,
                                             1, depth, nullptr, grad_strides, zero_init, image_numel, mdt::as_stream(stream));
    return MDT_EINVAL;
}

int mdt_crop_and_resize_3d_forward(const float *image, const int64_t *is, const float *boxes, const int *box_ind, int num_boxes, int batch, int H, int W,
                                   int Z, int ch, int cw, int cz, int depth, float /*extrapolation_value*/, float *crops, const int64_t *os, void *stream) {
    return mdt::roi_forward<3>(image, is, boxes, box_ind, num_boxes, batch, H, W, Z, ch, cw, cz, depth, crops, os, mdt::as_stream(stream));
}
int mdt_crop_and_resize_3d_backward(const float *grads, const int64_t *gs, const float *boxes, const int *box_ind, int num_boxes, int batch, int H, int W,
                                    int Z, int ch, int cw, int cz, int depth, float *gimg, const int64_t *is, int zero_init, int64_t image_numel,
                                    void *stream) {
    return mdt::roi_backward<3>(grads, gs, boxes, box_ind, num_boxes, batch, H, W, Z, ch, cw, cz, depth, gimg, is, zero_init, image_numel,
                                mdt::as_stream(stream));
}
static void pad2d(const int64_t *s, int64_t *o) { o[0] = s ? s[0] : 0; o[1] = s ? s[1] : 0; o[2] = s ? s[2] : 0; o[3] = s ? s[3] : 0; o[4] = 0; }
int mdt_crop_and_resize_2d_forward(const float *image, const int64_t *is, const float *boxes, const int *box_ind, int num_boxes, int batch, int H, int W,
                                   int ch, int cw, int depth, float /*extrapolation_value*/, float *crops, const int64_t *os, void *stream) {
    if (!is || !os) return MDT_EINVAL;
    int64_t i5[5], o5[5]; pad2d(is, i5); pad2d(os, o5);
    return mdt::roi_forward<2>(image, i5, boxes, box_ind, num_boxes, batch, H, W, 1, ch, cw, 1, depth, crops, o5, mdt::as_stream(stream));
}
int mdt_crop_and_resize_2d_backward(const float *grads, const int64_t *gs, const float *boxes, const int *box_ind, int num_boxes, int batch, int H, int W,
                                    int ch, int cw, int depth, float *gimg, const int64_t *is, int zero_init, int64_t image_numel, void *stream) {
    if (!is || !gs) return MDT_EINVAL;
    int64_t i5[5], g5[5]; pad2d(is, i5); pad2d(gs, g5);
    return mdt::roi_backward<2>(grads, g5, boxes, box_ind, num_boxes, batch, H, W, 1, ch, cw, 1, depth, gimg, i5, zero_init, image_numel,
                                mdt::as_stream(stream));
}

}  // extern "C"
