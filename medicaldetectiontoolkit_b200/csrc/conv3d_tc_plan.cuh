// Plan + parameter block of the tcgen05 implicit-GEMM conv kernels (conv3d_tc.cu).
#pragma once
#include <stdlib.h>

#include "conv3d_common.cuh"
#include "tc_common.cuh"

namespace mdt {

struct TcConvParams {
    int NB, RD, RH, RW;   // row space (fprop: output voxels; dgrad: input voxels)
    int SD, SH;           // source spatial extents along D, H (W handled by TMA out-of-bounds fill)
    int KD, KH, KW;
    int sd, sh;           // strides along D and H (W stride is 1 on this path)
    int pd, ph, pw;
    int dgrad;
    int Cn, Np, NT;       // true / padded column count; columns per CTA tile (<= 128)
    int nchunks, swz;     // K chunks per tap and swizzle span in bytes (chunk = swz/2 channels)
    int planes;           // 1 = bf16, 2 = split-bf16 x3
    int BW, BH, halo;     // tile = BH lines x BW voxels = 128 rows; halo mode iff BH == 1
    int tiles_w, tiles_h;
    int CPS, TPS, D;      // K chunks per stage, taps per stage (KW in halo mode, 1 otherwise), ring depth
    int a_plane_bytes, b_plane_bytes, a_tx_bytes, stage_bytes, a_region_bytes, a_chunk_bytes, b_chunk_bytes;
    int relu;
    int Q;                // independent accumulator chains per MMA type
    int wreps;            // weight replicas in global memory
    const float *bias, *residual;
    float *out;
    __nv_bfloat16 *out_split;   // optional: the result also as (hi, lo) bf16 planes in the canonical split layout [line][plane][w][out_kg]
    int out_kg;
};

constexpr int kTcThreads = 192;
constexpr int kTcMaxStages = 6;

// ------------------------------------------------------------------------------------------------ host side planning
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct TcPlan {
    bool ok = false;
    int Kc, Kp, Kg, Nc, Np, NT, n_tiles_n, swz, nchunks, BW, BH, halo, a_rows;   // Kp: K padding in shared memory; Kg: channels of the global split layout
    int RD, RH, RW, SD, SH, SW;  // row space / source space
    long long src_rows, dst_rows;
};

static inline TcPlan make_plan(const ConvGeom &g, int pass) {
    TcPlan pl;
    if (pass == 2) return pl;            // wgrad: not on this path yet
    if (g.sw != 1) return pl;            // W stride must be 1 (halo along W)
    const bool dgrad = pass == 1;
    pl.Kc = dgrad ? g.cout : g.cin;
    pl.Nc = dgrad ? g.cin : g.cout;
    pl.RD = dgrad ? g.d : g.od; pl.RH = dgrad ? g.h : g.oh; pl.RW = dgrad ? g.w : g.ow;
    pl.SD = dgrad ? g.od : g.d; pl.SH = dgrad ? g.oh : g.h; pl.SW = dgrad ? g.ow : g.w;
    // K padding: one swizzle-span chunk for <= 64 channels (few, large TMA boxes), 64-channel chunks above
    pl.Kp = pl.Kc <= 16 ? 16 : pl.Kc <= 32 ? 32 : ceil_div(pl.Kc, 64) * 64;   // == conv_tc_kpad_smem(); padding to 16 only was measured 24 % slower
    pl.Kg = ceil_div(pl.Kc, 16) * 16;                                          // == conv_tc_kpad()
    pl.Np = ceil_div(pl.Nc, 16) * 16;
    pl.NT = pl.Np <= 128 ? pl.Np : 128;
    pl.n_tiles_n = ceil_div(pl.Np, pl.NT);
    if (pl.Np % pl.NT) pl.Np = pl.n_tiles_n * pl.NT;   // keep the TMA box inside the packed weight tensor
    pl.swz = (pl.Kp % 64 == 0) ? 128 : (pl.Kp % 32 == 0) ? 64 : 32;
    pl.nchunks = pl.Kp / (pl.swz / 2);
    if (pl.RW >= 128) { pl.BW = 128; pl.BH = 1; }
    else {
        int bw = 16;
        while (bw < pl.RW) bw <<= 1;
        if (bw > 128 || pl.RW < 8) return pl;
        pl.BW = bw; pl.BH = 128 / bw;
        if (g.sh != 1) return pl;        // generic mode loads BH consecutive source lines
    }
    pl.halo = pl.BH == 1;
    pl.a_rows = pl.halo ? 128 + g.kw - 1 : 128;
    if (128 + g.kw - 1 > 256) return pl;
    pl.src_rows = (long long)g.n * pl.SD * pl.SH * pl.SW;
    pl.dst_rows = (long long)g.n * pl.RD * pl.RH * pl.RW;
    pl.ok = true;
    return pl;
}

// pipeline stage sizing: stage = CPS chunks x (A planes + TPS taps x B planes); prefer all kw taps of a halo line in one stage and >= 3
// stages; shrink the chunk group, then the tap group, until at least 2 stages fit in shared memory
static inline int a_rows_loaded(const TcPlan &pl, int kw) { return pl.halo ? 128 + kw - 1 : 128; }
static inline int a_chunk_bytes_of(const TcPlan &pl, int kw, int planes) { return (int)align_up((size_t)planes * a_rows_loaded(pl, kw) * pl.swz, 1024); }

static inline bool plan_stages(const TcPlan &pl, int kw, int planes, int &CPS, int &TPS, int &D, int &stage) {
    const int budget = 196 * 1024;
    const int a_chunk = a_chunk_bytes_of(pl, kw, planes), b_plane = pl.NT * pl.swz;
    auto bytes = [&](int cps, int tps) { return cps * (a_chunk + (int)align_up((size_t)tps * planes * b_plane, 1024)); };
    TPS = pl.halo ? kw : 1;
    CPS = pl.nchunks;
    while (CPS > 1 && 3 * bytes(CPS, TPS) > budget) --CPS;
    while (TPS > 1 && 2 * bytes(CPS, TPS) > budget) --TPS;
    if (2 * bytes(CPS, TPS) > budget) return false;
    // single-stage halo mode keeps A in place across tap groups, so a smaller tap group costs no extra traffic: shrink it until 4 CTAs fit
    if (pl.halo) while (TPS > 1 && bytes(CPS, TPS) > 54 * 1024) --TPS;
    // experiment knobs (tools/conv_layer_bench.py): MDT_TC_TPS caps the taps per stage, MDT_TC_D sets the ring depth
    if (const char *e = getenv("MDT_TC_TPS")) { const int v = atoi(e); if (v >= 1 && v < TPS) TPS = v; }
    stage = bytes(CPS, TPS);
    // Measured on B200 (profiles/r01_mma_rate.txt, r01_stage_sweep.txt): one thread cannot issue tcgen05.mma faster than ~60 cycles each, so for
    // the small N of these layers a single CTA leaves the tensor pipe idle; co-resident CTAs fill it.  Residency beats ring depth: keep the
    // footprint minimal (ONE stage, <= 64 registers/thread) so that 2-5 CTAs share an SM and overlap each other's loads, MMAs and epilogues
    // (sweep over D in {1,2} x taps-per-stage in profiles/r01_stage_sweep.txt: D = 1 with all kw taps per stage wins on every hot layer).
    D = 1;
    if (const char *e = getenv("MDT_TC_D")) { const int v = atoi(e); if (v >= 1 && v <= kTcMaxStages && v * stage <= budget) D = v; }
    return true;
}

}  // namespace mdt
