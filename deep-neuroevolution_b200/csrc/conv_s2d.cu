// conv_s2d.cu -- the member convolutions as shifted-window implicit GEMMs on tcgen05 (TMEM accumulators), A operand
// fed by TMA, persistent warp-specialised CTAs.  Replaces conv_tc_kernel (tc_conv.cu) on the per-tick path.
//
// Same contraction as the reference's per-member conv2d (policies.py:321-327,451-453; models/dqn.py:41-45,
// models/base.py:54-75: extract_image_patches + batched matmul; TF 'SAME', NHWC, HWIO):
//   out[m][n] = act( sum_{ky,kx,ci} in[oy*S-P+ky][ox*S-P+kx][ci] * w[ky][kx][ci][n] + b[n] ),   w = theta + s*noise[idx:]
//
// What changed against conv_tc_kernel (r01: 126 us per 256-slot tick, tensor pipe 13-18 %):
//   * NO im2col copy.  The (zero padded) input is space-to-depth'ed by the stride S, which turns the KSxKS/stride-S
//     convolution into a (KS/S)x(KS/S)/stride-1 convolution over a W x W pixel grid with S*S*CIN channels.  The image
//     is kept in shared memory as channel-octet planes  img[plane][pixel][8 x fp16]  -- which IS the UMMA K-major no-swizzle
//     canonical layout of a matrix whose rows are the pixels (8 pixels = one 128-byte core matrix, SBO = 128, next
//     channel octet LBO = PIXP*16 bytes; a 16-byte row holds 8 fp16 channels).  Output position m = oy*W + ox reads pixel m + ty*W + tx for tap (ty, tx): the
//     SAME image at a start address shifted by (ty*W+tx)*16 bytes.  One smem descriptor per (tap, channel octet, M tile);
//     rows with ox >= HOUT are junk accumulator rows that the epilogue skips (dev/tc_window.cu is the hardware
//     self-test of this addressing; tests/test_gpu_tc.py::test_tcgen05_shifted_window_operand).
//     Staging traffic per member drops from KS*KS/(S*S) x the input (conv3: 9x) to 1x, and for every layer but the first
//   * the image is not staged by threads at all: the PRODUCING layer's epilogue writes the next layer's image (already
//     space-to-depth'ed, zero padded, split into fp16 hi / lo planes) to global memory in exactly the shared-memory
//     layout, and a producer thread brings it in with cp.async.bulk (TMA), one channel-octet group per mbarrier, so the
//     MMAs of member i overlap the loads of member i+1 (a group's buffer is released by tcgen05.commit).
//     The first layer converts the uint8 frame (exact in fp16: one plane, /255 applied to the accumulator).
//   * the member's raw weights are not fetched by the staging threads either (r02 first version: one serialized global
//     round trip per 16 KB chunk, 35-47K cycles per member against 6-16K of MMA work): a second producer thread streams
//     the chunk's theta rows and noise rows (16-byte aligned supersets of the arbitrarily aligned slices) with
//     cp.async.bulk into the ring stage that will hold the operand tile, NSTB chunks deep; converter warps read the raw
//     rows from shared memory, perturb + split, and overwrite the SAME stage with the canonical [B_hi ; B_lo] tile.
//   * persistent CTAs (one per SM), three pipelines: A groups (TMA or staging warps <-> MMA), B ring (staging warps
//     <-> MMA: the member's perturbed weights fl(theta + fl(s*noise)) split h0/h1, [B_h0; B_h1] stacked along N), and a
//     double-buffered TMEM accumulator (MMA <-> epilogue warps), so staging, MMA issue and epilogue of consecutive
//     members overlap.
//   * arithmetic: tcgen05.mma kind::f16 on 2 x fp16 splits (tc05.cuh: x = h0 + h1*2^-11, 22 significand bits; uint8 pixels
//     are exact in fp16 and need one plane) -- twice the MAC rate and half the operand bytes of the 3xTF32 formulation the
//     first r02 version used (the kernels are MMA bound): main accumulator D0 = A_h0*B_h0, correction accumulator
//     D1 = A_h0*B_h1 + A_h1*B_h0 (B_h0 and B_h1 stacked along N: one N = 2*COUT MMA + one N = COUT MMA per K = 16 step),
//     result = D0 + 2^-11 * D1.  fp16 x fp16 products are exact in the fp32 accumulators.
#include "common.cuh"
#include "forward.cuh"
#include "epilogue.cuh"
#include "tc05.cuh"

using namespace tc05;

#ifdef DNE_S2D_TRACE       // dev timeline of CTA 0 (make EXTRA=-DDNE_S2D_TRACE OUT=...; tools/s2d_trace.py); never in the product build
__device__ long long g_s2d_trace[3][512];
extern "C" int dne_debug_s2d_trace(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, g_s2d_trace, sizeof(g_s2d_trace)) == cudaSuccess ? 0 : -3;
}
#define S2D_TR(cond, i) do { if (blockIdx.x == 0 && (cond) && (i) < 512) g_s2d_trace[TRL][(i)] = clock64(); } while (0)
#define S2D_EV(it, g, e) (16 + (((it) * 16 + (g)) * 8 + (e)))
#else
#define S2D_TR(cond, i) do { } while (0)
#define S2D_EV(it, g, e) 0
#endif

namespace {

constexpr int S2D_STAGE_WARPS = 8;                         // B (and uint8 A) staging
constexpr int S2D_EPI_WARPS = 8;                           // two per TMEM lane quarter (they split the column groups)
constexpr int S2D_STAGE_THREADS = S2D_STAGE_WARPS * 32;
constexpr int S2D_EPI_THREADS = S2D_EPI_WARPS * 32;
constexpr int S2D_THREADS = S2D_STAGE_THREADS + S2D_EPI_THREADS + 96;   // + MMA warp + image producer warp + weight producer warp
constexpr int S2D_FRAME_BYTES = 84 * 84 * 4;                            // the uint8 frame stack of the first layer
constexpr int S2D_FRAME_STRIDE = (S2D_FRAME_BYTES + 127) / 128 * 128;
constexpr int S2D_MAX_GROUPS = 8;
constexpr int S2D_MAX_BST = 8;       // ring depth: a stage cycles through TMA latency -> conversion -> MMA, ~4.6K cycles (conv3): 4 stages were the limit
constexpr int S2D_SMEM_BUDGET = 226 * 1024;

constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

// Geometry of one layer's space-to-depth image (shared by the consumer kernel, the producing epilogue and the host).
struct S2dGeom {
    int S, PADB, W, KT, CIN, CP, NG, NPIX, PIXP, PARTS, LBO, GROUP_BYTES, IMG_BYTES, HIN;
};
__host__ __device__ constexpr S2dGeom s2d_geom(int CIN, int KS, int S, int HIN, int HOUT, int PAD, bool in_u8) {
    S2dGeom g{};
    const int HP = (HOUT - 1) * S + KS;
    // plane = 8 channels (one 16-byte fp16 row per pixel); group = 16 channels = one K = 16 MMA step = 2 planes per part
    g.S = S; g.PADB = PAD; g.W = HP / S; g.KT = KS / S; g.CIN = CIN; g.CP = S * S * CIN; g.NG = g.CP / 16;
    g.NPIX = g.W * g.W; g.PIXP = (g.NPIX + 7) / 8 * 8; g.PARTS = in_u8 ? 1 : 2; g.LBO = g.PIXP * 16;
    g.GROUP_BYTES = g.PARTS * 2 * g.LBO; g.IMG_BYTES = g.NG * g.GROUP_BYTES; g.HIN = HIN;
    return g;
}

constexpr int s2d_taps_per_chunk(int ntap, int piece_stride) {       // largest divisor of ntap whose raw landing zone fits ~26 KB
    int best = 1;
    for (int t = 1; t <= ntap; ++t)
        if (ntap % t == 0 && 2 * t * piece_stride <= 26000) best = t;
    return best;
}

template <int CIN, int COUT, int KS, int S, int HIN, int HOUT, int PAD, bool IN_U8>
struct S2dCfg {
    static constexpr S2dGeom G = s2d_geom(CIN, KS, S, HIN, HOUT, PAD, IN_U8);
    static_assert(((HOUT - 1) * S + KS) % S == 0 && KS % S == 0, "space-to-depth needs S | KS and S | padded size");
    static_assert(CIN % 4 == 0 && COUT % 16 == 0 && G.CP % 16 == 0 && G.NG <= S2D_MAX_GROUPS, "tile constraints");
    static constexpr int W = G.W, KT = G.KT, NTAP = KT * KT, NG = G.NG, PIXP = G.PIXP, LBO_A = G.LBO;
    static constexpr int PARTS = G.PARTS, GROUP_BYTES = G.GROUP_BYTES, IMG_BYTES = G.IMG_BYTES;
    static constexpr int MMAX = (HOUT - 1) * (W + 1);            // largest valid accumulator row
    static constexpr int MT = MMAX / 128 + 1;                    // M tiles of 128 rows
    static constexpr int MAXOFF = (KT - 1) * (W + 1);
    static constexpr int REACH = MT * 128 + MAXOFF;              // pixels a descriptor may touch from a plane start
    static constexpr int SLACK = REACH > PIXP ? ((REACH - PIXP) * 16 + 127) / 128 * 128 : 0;
    static constexpr int A_REGION = IMG_BYTES + SLACK;
    static constexpr int LBO_B = 2 * COUT * 16;                  // [B_h0 ; B_h1] stacked along N, 8 fp16 k values per row
    // raw landing zone of a chunk inside its ring stage: per tap one piece of 16 contiguous fp32 weight rows for theta and
    // one for the noise; each piece is the 16-byte aligned superset of its rows (+16 bytes)
    static constexpr int PIECE = 16 * COUT * 4;
    static constexpr int PIECE_STRIDE = PIECE + 16;
    static constexpr int TPC = s2d_taps_per_chunk(NTAP, PIECE_STRIDE);   // taps per chunk
    static constexpr int NCPG = NTAP / TPC;                      // chunks per channel group
    static constexpr int NCH = NG * NCPG;                        // chunks per member
    static constexpr int B_TILE = TPC * 2 * LBO_B;               // TPC taps x 16 k = TPC*2 k-octet planes
    static constexpr int RAW_BYTES = 2 * TPC * PIECE_STRIDE;
    static constexpr int BST_BYTES = (cmax(B_TILE, RAW_BYTES) + 127) / 128 * 128;
    static constexpr int FRAME_REGION = IN_U8 ? 2 * S2D_FRAME_STRIDE : 0;        // double-buffered raw uint8 frame
    static constexpr int NSTB = cmin(S2D_MAX_BST, (S2D_SMEM_BUDGET - A_REGION - FRAME_REGION - 256) / BST_BYTES);
    static_assert(NSTB >= 2, "shared memory: B ring too shallow");
    static constexpr int SMEM_BYTES = A_REGION + NSTB * BST_BYTES + FRAME_REGION + 256;
    static constexpr int ACC_COLS = MT * 2 * COUT;               // one accumulator buffer: per M tile [main | correction]
    static constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128 : 2 * ACC_COLS <= 256 ? 256 : 512;
    static_assert(2 * ACC_COLS <= 512, "TMEM: double-buffered accumulators do not fit");
    static constexpr int B_UNITS = COUT * TPC * 2;               // (column n, k octet) units per chunk
    // converter groups: chunk c is converted by group c % NGRP (independent streams hide the per-chunk hand-off latency)
    static constexpr int NGRP = 2;
    static constexpr int WPG = S2D_STAGE_WARPS / NGRP, TG = 32 * WPG;
    static constexpr int B_UPT = (B_UNITS + TG - 1) / TG;
    static_assert(TG % COUT == 0 && NCH % NGRP == 0, "B unit map: a thread keeps its column; groups alternate chunks");
    static_assert(!IN_U8 || NCPG == 1, "the uint8 frame is staged per channel group by the group's (only) chunk");
};

// Where the epilogue writes: NHWC floats (feeding a dense layer) or the NEXT conv layer's image.
struct S2dOut {
    float* base;
    int64_t slot_stride;        // floats
    int next_img;               // 0: NHWC [HOUT*HOUT][COUT];  1: image of the next layer (geometry below)
    int nS, nPADB, nW, nPIXP, nHP;
    float* xc;                  // nullable (NHWC mode only): the same vector as the A operand of the TMA-fed theta GEMM,
    int xc_ko;                  //   Xc[slot / 128][k octet][h0 | h1][slot % 128][8 x fp16]  (theta_gemm_tma.cu), xc_ko = K / 8
};

// ------------------------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int KS, int S, int HIN, int HOUT, int PAD, bool IN_U8>
__global__ void __launch_bounds__(S2D_THREADS, 1)
conv_s2d_kernel(SlotArgs sa, int64_t off_w, LayerEpi epi, const void* __restrict__ in_base, int64_t in_slot_stride,
                S2dOut so, int n_slots, int vdiv, int in_mod) {
    // vdiv > 1 (virtual-batch-norm reference pass, vbn_kernels.cu): the launch covers n_slots = members * vdiv VIRTUAL
    // slots; virtual slot v is image v % vdiv of member v / vdiv -- the member indexes the slot table (theta row, noise
    // index, scale, active flag, BN statistics), v indexes the input / output buffers.  in_mod > 0: the first layer's frames
    // are shared by all members (frame v % in_mod).
    using Cfg = S2dCfg<CIN, COUT, KS, S, HIN, HOUT, PAD, IN_U8>;
    constexpr int NG = Cfg::NG, NSTB = Cfg::NSTB, MT = Cfg::MT, NTAP = Cfg::NTAP, W = Cfg::W, KT = Cfg::KT;
    constexpr int TPC = Cfg::TPC, NCPG = Cfg::NCPG, NCH = Cfg::NCH;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    __shared__ uint64_t a_full[NG], a_empty[NG], raw_full[NSTB], b_full[NSTB], b_empty[NSTB], acc_full[2], acc_empty[2];
    __shared__ uint64_t frame_full[2], frame_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[COUT], s_mean[COUT], s_inv[COUT], s_gamma[COUT], s_beta[COUT];   // per-channel epilogue

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();                                               // common.cuh: the next kernel of the tick may be scheduled
#ifdef DNE_S2D_TRACE
    constexpr int TRL = IN_U8 ? 0 : (KS == 4 ? 1 : 2);
#endif
    S2D_TR(tid == 0, 0);
    const uint32_t sA = smem_u32(smem), sB = sA + Cfg::A_REGION;
    uint8_t* const gB = smem + Cfg::A_REGION;                                    // generic view of the ring (bulk copies)
    uint8_t* const gFrame = gB + NSTB * Cfg::BST_BYTES;

    if (warp == 0) tmem_alloc(&tmem_base_s, Cfg::TMEM_COLS);
    if (tid == 32) {
        for (int i = 0; i < NG; ++i) {
            mbar_init(&a_full[i], IN_U8 ? Cfg::WPG : 1);          // the converting group (uint8 frame) or the TMA producer
            mbar_init(&a_empty[i], 1);                            // tcgen05.commit
        }
        for (int i = 0; i < NSTB; ++i) {
            mbar_init(&raw_full[i], 1);                           // weight producer (expect_tx)
            mbar_init(&b_full[i], Cfg::WPG);                      // the warps of the converting group
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&frame_full[i], 1);
            mbar_init(&frame_empty[i], S2D_STAGE_WARPS);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], S2D_EPI_WARPS);
        }
        fence_mbar_init();
    }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = tmem_base_s;
    S2D_TR(tid == 0, 1);

    if (warp == S2D_STAGE_WARPS + S2D_EPI_WARPS) {
        // ================= MMA warp: converged loop, one elected lane issues (tc05.cuh: elect_one) =================
        constexpr uint32_t IDESC2 = idesc_f16(128, 2 * COUT), IDESC1 = idesc_f16(128, COUT);
        const uint64_t dA0 = smem_desc(sA, Cfg::LBO_A, 128), dB0 = smem_desc(sB, Cfg::LBO_B, 128);
        uint32_t cb = 0, it = 0;
        for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
            if (!slot_active(sa, slot / vdiv)) continue;
            const uint32_t buf = it & 1;
            mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator buffer
            fence_after_thread_sync();
            const uint32_t d0 = tmem_base + buf * Cfg::ACC_COLS;
            for (int c = 0; c < NCH; ++c, ++cb) {
                const int g = c / NCPG, tc = c - g * NCPG;
                const uint32_t st = cb % NSTB;
                if (tc == 0) mbar_wait(&a_full[g], it & 1);
                S2D_TR(lane == 0 && it < 2 && c < 16, S2D_EV(it, c, 4));
                mbar_wait(&b_full[st], (cb / NSTB) & 1);
                S2D_TR(lane == 0 && it < 2 && c < 16, S2D_EV(it, c, 5));
                fence_after_thread_sync();
                if (elect_one()) {
                    const uint64_t dAg = dA0 + (uint64_t)((g * Cfg::GROUP_BYTES) >> 4);
                    const uint64_t dBs = dB0 + (uint64_t)((st * Cfg::BST_BYTES) >> 4);
#pragma unroll
                    for (int tl = 0; tl < TPC; ++tl) {
                        const int tap = tc * TPC + tl;
                        const int toff = (tap / KT) * W + (tap % KT);
                        const uint64_t dB = dBs + (uint64_t)((tl * 2 * Cfg::LBO_B) >> 4);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint64_t dAh = dAg + (uint64_t)(mt * 128 + toff);         // 16-byte units == pixels
                            const uint32_t d = d0 + mt * 2 * COUT;
                            mma_f16(d, dAh, dB, IDESC2, (c | tl) != 0);                                       // A_h0 * [B_h0 ; B_h1]
                            if (!IN_U8) mma_f16(d + COUT, dAh + (uint64_t)((2 * Cfg::LBO_A) >> 4), dB, IDESC1, 1);   // A_h1 * B_h0 -> correction
                        }
                    }
                    mma_commit(&b_empty[st]);
                    if (tc == NCPG - 1) mma_commit(&a_empty[g]);
                    if (c == NCH - 1) mma_commit(&acc_full[buf]);
                }
                __syncwarp();
                S2D_TR(lane == 0 && it < 2 && c < 16, S2D_EV(it, c, 6));
            }
            ++it;
        }
    } else if (warp == S2D_STAGE_WARPS + S2D_EPI_WARPS + 1) {
        // ================= image producer (TMA): the member's image, one channel-octet group per bulk copy; =================
        // ================= first layer: the member's raw uint8 frame (double buffered)                      =================
        if (lane == 0) {
            pdl_wait();                                          // the image / frame is written by the previous kernel of the tick
            uint32_t it = 0;
            for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
                if (!slot_active(sa, slot / vdiv)) continue;
                if (IN_U8) {
                    const uint32_t fb = it & 1;
                    mbar_wait(&frame_empty[fb], ((it >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&frame_full[fb], S2D_FRAME_BYTES);
                    const int fslot = in_mod > 0 ? slot % in_mod : slot;
                    bulk_g2s(gFrame + fb * S2D_FRAME_STRIDE, (const uint8_t*)in_base + fslot * in_slot_stride, S2D_FRAME_BYTES,
                             &frame_full[fb]);
                } else {
                    const uint8_t* src = (const uint8_t*)((const float*)in_base + slot * in_slot_stride);
                    for (int g = 0; g < NG; ++g) {
                        mbar_wait(&a_empty[g], (it & 1) ^ 1);
                        mbar_arrive_expect_tx(&a_full[g], Cfg::GROUP_BYTES);
                        bulk_g2s(smem + g * Cfg::GROUP_BYTES, src + (size_t)g * Cfg::GROUP_BYTES, Cfg::GROUP_BYTES, &a_full[g]);
                    }
                }
                ++it;
            }
        }
    } else if (warp == S2D_STAGE_WARPS + S2D_EPI_WARPS + 2) {
        // ================= weight producer (TMA): raw theta rows + raw noise rows of every chunk, NSTB chunks deep =================
        if (lane == 0) {
            uint32_t cb = 0;
            for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
                const int ms = slot / vdiv;
                if (!slot_active(sa, ms)) continue;
                const float* th = slot_theta(sa, ms) + off_w;
                const float* nz = sa.noise + sa.noise_idx[ms] + off_w;
                const int a_t = (int)(((uintptr_t)th >> 2) & 3), a_n = (int)(((uintptr_t)nz >> 2) & 3);
                for (int c = 0; c < NCH; ++c, ++cb) {
                    const int g = c / NCPG, tc = c - g * NCPG;
                    const uint32_t st = cb % NSTB;
                    mbar_wait(&b_empty[st], ((cb / NSTB) & 1) ^ 1);
                    S2D_TR(cb < 2 * NCH && c < 16, S2D_EV(cb / NCH, c, 0));
                    mbar_arrive_expect_tx(&raw_full[st], Cfg::RAW_BYTES);
                    uint8_t* dst = gB + st * Cfg::BST_BYTES;
                    const int cp = 16 * g, pp = cp / CIN, ci = cp % CIN, py = pp / S, px = pp % S;
#pragma unroll
                    for (int tl = 0; tl < TPC; ++tl) {
                        const int tap = tc * TPC + tl;
                        const int ty = tap / KT, tx = tap % KT;
                        const int kk = ((ty * S + py) * KS + (tx * S + px)) * CIN + ci;          // first of 16 contiguous weight rows
                        bulk_g2s(dst + tl * Cfg::PIECE_STRIDE, th + (int64_t)kk * COUT - a_t, Cfg::PIECE_STRIDE, &raw_full[st]);
                        bulk_g2s(dst + (TPC + tl) * Cfg::PIECE_STRIDE, nz + (int64_t)kk * COUT - a_n, Cfg::PIECE_STRIDE, &raw_full[st]);
                    }
                    S2D_TR(cb < 2 * NCH && c < 16, S2D_EV(cb / NCH, c, 1));
                }
            }
        }
    } else if (warp < S2D_STAGE_WARPS) {
        // ================= converter warps: raw rows -> perturbed, split [B_hi ; B_lo] tile, in place =================
        // =================                  (first layer: also the uint8 frame -> image planes)       =================
        constexpr int TG = Cfg::TG, NGRP = Cfg::NGRP;
        const int grp = warp / Cfg::WPG, tg = tid - grp * TG;
        // zero fill, once: the uint8 variant relies on the zero padding of its image never being overwritten (the
        // converter warps are its only writers); the TMA-fed variants only need finite values in the slack behind the last
        // plane, which is read into junk accumulator rows.  Only the converter warps wait for it.
        {
            constexpr int Z0 = IN_U8 ? 0 : Cfg::IMG_BYTES, Z1 = Cfg::A_REGION;
            for (int i = Z0 / 16 + tid; i < Z1 / 16; i += S2D_STAGE_THREADS) sts128(sA + i * 16, make_float4(0.f, 0.f, 0.f, 0.f));
            fence_proxy_async_smem();
            named_bar_sync(5, S2D_STAGE_THREADS);
        }
        const int n = tg % COUT;                                 // this thread's output channel in every unit
        constexpr int KQ_STEP = TG / COUT;
        const int kq0 = tg / COUT;
        uint32_t it = 0;
        for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
            const int ms = slot / vdiv;
            if (!slot_active(sa, ms)) continue;
            const float* th = slot_theta(sa, ms) + off_w;
            const float* nz = sa.noise + sa.noise_idx[ms] + off_w;
            const int a_t = (int)(((uintptr_t)th >> 2) & 3), a_n = (int)(((uintptr_t)nz >> 2) & 3);
            const float s = sa.scale[ms];
            if (IN_U8) mbar_wait(&frame_full[it & 1], (it >> 1) & 1);
            for (int c = grp; c < NCH; c += NGRP) {
                const int g = c / NCPG;
                const uint32_t cb = it * NCH + c, st = cb % NSTB;
                if (IN_U8) {
                    // ---- the uint8 frame, channel group g = py (all four px phases): plane h holds the pixel pair
                    // px in {2h, 2h+1}: 8 bytes -> 8 fp16 channels = one 16-byte row ----
                    static_assert(!IN_U8 || (CIN == 4 && S == 4 && HIN == 84), "uint8 staging is written for the 84x84x4 frame stack, stride 4");
                    const uint32_t frame = smem_u32(gFrame + (it & 1) * S2D_FRAME_STRIDE);
                    const int py = g;
                    mbar_wait(&a_empty[g], (it & 1) ^ 1);
                    for (int u = tg; u < 2 * W * W; u += TG) {
                        const int h = u / (W * W), pix = u - h * (W * W);
                        const int yq = pix / W, xq = pix - yq * W;
                        const int y = 4 * yq + py - PAD, x0 = 4 * xq + 2 * h - PAD;
                        if (y >= 0 && y < HIN && x0 >= 0 && x0 < HIN) {
                            uint32_t p0, p1;
                            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(p0), "=r"(p1) : "r"(frame + (y * HIN + x0) * 4));
                            auto h2 = [](uint32_t lo8, uint32_t hi8) {        // two bytes -> two fp16 (exact)
                                return (uint32_t)__half_as_ushort(__ushort2half_rn((unsigned short)lo8)) |
                                       ((uint32_t)__half_as_ushort(__ushort2half_rn((unsigned short)hi8)) << 16);
                            };
                            const uint4 row = make_uint4(h2(p0 & 255u, (p0 >> 8) & 255u), h2((p0 >> 16) & 255u, p0 >> 24),
                                                         h2(p1 & 255u, (p1 >> 8) & 255u), h2((p1 >> 16) & 255u, p1 >> 24));
                            sts128u(sA + g * Cfg::GROUP_BYTES + h * Cfg::LBO_A + pix * 16, row);
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&a_full[g]);
                        if (c + NGRP >= NCH) mbar_arrive(&frame_empty[it & 1]);    // this warp's last read of the raw frame
                    }
                }
                S2D_TR(tg == 0 && it < 2 && c < 16, S2D_EV(it, c, 7));
                mbar_wait(&raw_full[st], (cb / NSTB) & 1);
                S2D_TR(tg == 0 && it < 2 && c < 16, S2D_EV(it, c, 2));
                const uint32_t sBs = sB + st * Cfg::BST_BYTES;
                float w[Cfg::B_UPT][8];
#pragma unroll
                for (int i = 0; i < Cfg::B_UPT; ++i) {
                    const int kp = kq0 + i * KQ_STEP;            // local k octet: tap kp >> 1 of the chunk, rows 8*(kp & 1) .. +7 of its piece
                    if (kp < TPC * 2) {
                        const uint32_t pt = sBs + (kp >> 1) * Cfg::PIECE_STRIDE + (uint32_t)((a_t + (8 * (kp & 1)) * COUT + n) * 4);
                        const uint32_t pn = sBs + (TPC + (kp >> 1)) * Cfg::PIECE_STRIDE + (uint32_t)((a_n + (8 * (kp & 1)) * COUT + n) * 4);
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[i][j] = perturbed(lds32(pt + j * COUT * 4), s, lds32(pn + j * COUT * 4));
                    }
                }
                named_bar_sync(3 + grp, TG);                     // every raw value of the stage is in registers: overwrite it
#pragma unroll
                for (int i = 0; i < Cfg::B_UPT; ++i) {
                    const int kp = kq0 + i * KQ_STEP;
                    if (kp < TPC * 2) {
                        uint4 hi, lo;
                        split_f16x8(w[i], hi, lo);
                        sts128u(sBs + kp * Cfg::LBO_B + n * 16, hi);
                        sts128u(sBs + kp * Cfg::LBO_B + (COUT + n) * 16, lo);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[st]);
                S2D_TR(tg == 0 && it < 2 && c < 16, S2D_EV(it, c, 3));
            }
            ++it;
        }
    } else {
        // ================= epilogue warps: TMEM -> (/255) + bias (+BN) + activation -> global =================
        // warp ew: TMEM lane quarter ew & 3 (== warp % 4, the hardware rule), column-group parity ew >> 2
        const int ew = warp - S2D_STAGE_WARPS, lq = ew & 3, half = ew >> 2;
        const int et = tid - S2D_STAGE_THREADS;
        constexpr float IN_SCALE = IN_U8 ? (1.0f / 255.0f) : 1.0f;
        constexpr int NJ = COUT / 16;
        const int act = epi.act;
        const bool bn = epi.bn != DNE_BN_NONE;
        uint32_t it = 0;
        pdl_wait();                                              // before the first global write (the zero padding below)
        for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
            const int ms = slot / vdiv;
            if (!slot_active(sa, ms)) continue;
            const float* th = slot_theta(sa, ms);
            const int64_t idx = sa.noise_idx[ms];
            const float s = sa.scale[ms];
            for (int c = et; c < COUT; c += S2D_EPI_THREADS) {
                const ChanEpi ce = make_chan_epi(sa, epi, ms, COUT, c, th, idx, s);
                s_bias[c] = ce.bias; s_mean[c] = ce.mean; s_inv[c] = ce.inv; s_gamma[c] = ce.gamma; s_beta[c] = ce.beta;
            }
            float* outp = so.base + slot * so.slot_stride;
            if (so.next_img) {
                // zero padding of the next layer's image: pixels (Y, X) of its padded grid that no output maps to
                const int nHP = so.nHP, no = COUT / 8;
                for (int b = et; b < nHP * nHP; b += S2D_EPI_THREADS) {
                    const int Y = b / nHP, X = b - Y * nHP;
                    if (Y >= so.nPADB && Y < so.nPADB + HOUT && X >= so.nPADB && X < so.nPADB + HOUT) continue;
                    const int pix = (Y / so.nS) * so.nW + (X / so.nS), pp = (Y % so.nS) * so.nS + (X % so.nS);
                    for (int q = 0; q < no; ++q) {
                        const int co = pp * no + q;                                   // channel octet of the next image
                        uint4* p = reinterpret_cast<uint4*>(outp) + (size_t)((co >> 1) * 4 + (co & 1)) * so.nPIXP + pix;
                        p[0] = make_uint4(0u, 0u, 0u, 0u);
                        p[(size_t)2 * so.nPIXP] = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            }
            named_bar_sync(2, S2D_EPI_THREADS);                  // per-channel parameters ready
            const uint32_t buf = it & 1;
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            S2D_TR(et == 0 && it < 2, 2 + it);
            fence_after_thread_sync();
            const uint32_t t0 = tmem_base + buf * Cfg::ACC_COLS + ((uint32_t)(lq * 32) << 16);
#pragma unroll 1
            for (int q = half; q < MT * NJ; q += 2) {             // (M tile, 16-column group) work items of this warp
                const int mt = q / NJ, n0 = (q - mt * NJ) * 16;
                const int m = mt * 128 + lq * 32 + lane;
                const int oy = m / W, ox = m - oy * W;
                const bool valid = (oy < HOUT) && (ox < HOUT);
                float v[16], v2[16];
                __syncwarp();                                    // tcgen05.ld is .sync.aligned: the warp must be converged
                tmem_ld16_async(t0 + (uint32_t)(mt * 2 * COUT + n0), v);
                tmem_ld16_async(t0 + (uint32_t)(mt * 2 * COUT + COUT + n0), v2);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int x = 0; x < 16; x += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[n0 + x]);
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            float r = fmaf(v2[x + y], F16_LO_INV, v[x + y]);        // main + 2^-11 * correction accumulator
                            if (IN_U8) r *= IN_SCALE;
                            r += bb[y];
                            if (bn) r = (r - s_mean[n0 + x + y]) * s_inv[n0 + x + y] * s_gamma[n0 + x + y] + s_beta[n0 + x + y];   // policies.py:322
                            v[x + y] = act == DNE_ACT_RELU ? fmaxf(r, 0.0f) : (act == DNE_ACT_TANH ? tanhf(r) : r);
                        }
                    }
                    if (!so.next_img) {
                        float4* dst = reinterpret_cast<float4*>(outp + (int64_t)(oy * HOUT + ox) * COUT + n0);
#pragma unroll
                        for (int x = 0; x < 16; x += 4) dst[x / 4] = make_float4(v[x], v[x + 1], v[x + 2], v[x + 3]);
                        if (so.xc) {
                            const int ko0 = ((oy * HOUT + ox) * COUT + n0) >> 3;
                            uint4* xp = reinterpret_cast<uint4*>(so.xc) + ((int64_t)(slot >> 7) * so.xc_ko + ko0) * 256 + (slot & 127);
#pragma unroll
                            for (int x = 0; x < 16; x += 8) {
                                const float e[8] = {v[x], v[x + 1], v[x + 2], v[x + 3], v[x + 4], v[x + 5], v[x + 6], v[x + 7]};
                                uint4 hi, lo;
                                split_f16x8(e, hi, lo);
                                xp[(x / 8) * 256] = hi;
                                xp[(x / 8) * 256 + 128] = lo;
                            }
                        }
                    } else {
                        const int Y = oy + so.nPADB, X = ox + so.nPADB;
                        const int pix = (Y / so.nS) * so.nW + (X / so.nS);
                        const int pp = (Y % so.nS) * so.nS + (X % so.nS);
#pragma unroll
                        for (int x = 0; x < 16; x += 8) {
                            const int co = pp * (COUT / 8) + (n0 + x) / 8;            // channel octet of the next image
                            const float e[8] = {v[x], v[x + 1], v[x + 2], v[x + 3], v[x + 4], v[x + 5], v[x + 6], v[x + 7]};
                            uint4 hi, lo;
                            split_f16x8(e, hi, lo);
                            uint4* p = reinterpret_cast<uint4*>(outp) + (size_t)((co >> 1) * 4 + (co & 1)) * so.nPIXP + pix;
                            p[0] = hi;
                            p[(size_t)2 * so.nPIXP] = lo;
                        }
                    }
                }
            }
            fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            S2D_TR(et == 0 && it < 2, 4 + it);
            named_bar_sync(2, S2D_EPI_THREADS);                  // per-channel parameters reusable
            ++it;
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    S2D_TR(tid == 0, 6);
    if (warp == 0) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int CIN, int COUT, int KS, int S, int HIN, int HOUT, int PAD, bool IN_U8>
int launch_s2d(const SlotArgs& sa, const dne_layer_desc& L, const LayerEpi& epi, const void* in, int64_t in_slot_stride,
               const S2dOut& so, int n_slots, int sm_count, cudaStream_t st, int vdiv, int in_mod) {
    using Cfg = S2dCfg<CIN, COUT, KS, S, HIN, HOUT, PAD, IN_U8>;
    auto kern = conv_s2d_kernel<CIN, COUT, KS, S, HIN, HOUT, PAD, IN_U8>;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_done[64] = {};                              // per device (one context per device and process)
    if (dev < 64 && !attr_done[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
            return DNE_ERR_CUDA;
        attr_done[dev] = true;
    }
    const int grid = n_slots < sm_count ? n_slots : sm_count;
    // the first layer follows host copies / the previous tick's graph: launched fully serialized.  The weight producer and
    // the converter warps of the later layers never wait: theta and the noise table are not written inside a tick.
    // (dne_set_option("chain_ticks", 1): the caller guarantees that the stream's previous kernel is the previous tick's head --
    // nothing that writes theta, the noise table or the slot table -- and the first layer joins the chain too.)
    if (dne_launch_chain(kern, dim3(grid), dim3(S2D_THREADS), (size_t)Cfg::SMEM_BYTES, st, !IN_U8 || g_dne_chain_ticks, sa, (int64_t)L.off_w, epi, in,
                         in_slot_stride, so, n_slots, vdiv, in_mod) != cudaSuccess)
        return DNE_ERR_CUDA;
    DNE_LAUNCHED(1);
    return 0;
}

bool shape_is(const dne_layer_desc& L, int cin, int cout, int ks, int stride, int hin, int hout, int pad) {
    return L.kind == DNE_CONV && L.cin == cin && L.cout == cout && L.ksize == ks && L.stride == stride && L.hin == hin &&
           L.hout == hout && L.pad == pad;
}

}  // namespace

// ---- host interface (forward.cuh) ---------------------------------------------------------------------------------
// Shapes compiled in: the Nature-DQN family of the reference (models/dqn.py:25-47, policies.py:321-327,451-453).
bool dne_s2d_supported(const dne_layer_desc& L, bool in_u8) {
    if (in_u8) return shape_is(L, 4, 32, 8, 4, 84, 21, 2) || shape_is(L, 4, 16, 8, 4, 84, 21, 2);
    return shape_is(L, 32, 64, 4, 2, 21, 11, 1) || shape_is(L, 16, 32, 4, 2, 21, 11, 1) || shape_is(L, 64, 64, 3, 1, 11, 11, 1);
}

// Bytes of layer L's INPUT image (what the producing epilogue writes per slot); 0 if L is not an s2d conv layer.
size_t dne_s2d_image_bytes(const dne_layer_desc& L) {
    if (!dne_s2d_supported(L, false)) return 0;
    const S2dGeom g = s2d_geom(L.cin, L.ksize, L.stride, L.hin, L.hout, L.pad, false);
    return (size_t)g.IMG_BYTES;
}

// Geometry of layer L's input image for an external writer (vbn_image_kernel): S, PADB, W, PIXP, HP = W * S.
void dne_s2d_image_geom(const dne_layer_desc& L, int* nS, int* nPADB, int* nW, int* nPIXP, int* nHP) {
    const S2dGeom g = s2d_geom(L.cin, L.ksize, L.stride, L.hin, L.hout, L.pad, false);
    *nS = g.S; *nPADB = g.PADB; *nW = g.W; *nPIXP = g.PIXP; *nHP = g.W * g.S;
}

// next: the following conv layer if it consumes an image (nullptr: write NHWC floats).
int dne_launch_conv_layer_s2d(const SlotArgs& sa, const dne_layer_desc& L, const LayerEpi& epi, bool in_u8, const void* in,
                              int64_t in_slot_stride, float* out, int64_t out_slot_stride, const dne_layer_desc* next,
                              int n_slots, int sm_count, cudaStream_t st, float* xc, int vdiv, int in_mod) {
    S2dOut so;
    so.xc = next ? nullptr : xc;
    so.xc_ko = (L.hout * L.hout * L.cout) / 8;
    so.base = out;
    so.slot_stride = out_slot_stride;
    so.next_img = next ? 1 : 0;
    so.nS = so.nPADB = so.nW = so.nPIXP = so.nHP = 1;
    if (next) {
        const S2dGeom g = s2d_geom(next->cin, next->ksize, next->stride, next->hin, next->hout, next->pad, false);
        so.nS = g.S; so.nPADB = g.PADB; so.nW = g.W; so.nPIXP = g.PIXP; so.nHP = g.W * g.S;
    }
#define ARGS sa, L, epi, in, in_slot_stride, so, n_slots, sm_count, st, (vdiv < 1 ? 1 : vdiv), in_mod
    if (in_u8 && shape_is(L, 4, 32, 8, 4, 84, 21, 2)) return launch_s2d<4, 32, 8, 4, 84, 21, 2, true>(ARGS);
    if (in_u8 && shape_is(L, 4, 16, 8, 4, 84, 21, 2)) return launch_s2d<4, 16, 8, 4, 84, 21, 2, true>(ARGS);
    if (!in_u8 && shape_is(L, 32, 64, 4, 2, 21, 11, 1)) return launch_s2d<32, 64, 4, 2, 21, 11, 1, false>(ARGS);
    if (!in_u8 && shape_is(L, 16, 32, 4, 2, 21, 11, 1)) return launch_s2d<16, 32, 4, 2, 21, 11, 1, false>(ARGS);
    if (!in_u8 && shape_is(L, 64, 64, 3, 1, 11, 11, 1)) return launch_s2d<64, 64, 3, 1, 11, 11, 1, false>(ARGS);
#undef ARGS
    return DNE_ERR_UNSUP;
}
