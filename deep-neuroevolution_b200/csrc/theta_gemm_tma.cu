// theta_gemm_tma.cu -- the shared-theta part of a dense layer, X[slots,K] . theta_w[K,N], as a pure TMA + tcgen05 kernel.
//
//   part[split][m][n] = sum_{k in split} X[m][k] * W[k][n]          (same contract as theta_gemm_tc_kernel / dense_theta_gemm_kernel)
//
// The r01 kernel (tc_conv.cu: theta_gemm_tc_kernel, 31 us per 256-slot LargeModel tick) staged both operands through the
// threads (global -> registers -> TF32 split -> st.shared) with two stages and a block barrier per 16-wide chunk.  Here
// both operands already exist in global memory in the UMMA K-major canonical layout, as 2 x fp16 splits (tc05.cuh:
// x = h0 + h1*2^-11; kind::f16 runs at twice the MAC rate of kind::tf32 and the operands are half the bytes):
//   * W changes once per generation: dne_theta_prepare() relays it out (theta_prep_kernel) into
//       Wc[n tile of 128][k octet][h0 | h1][128 n][8 k]     one k-octet plane = [B_h0 ; B_h1] stacked along N = 4 KB
//   * X is the output of the last convolution: its epilogue (conv_s2d.cu) writes, besides the NHWC vector the noise GEMV
//     streams, the same values as
//       Xc[m tile of 128][k octet][h0 | h1][128 slots][8 k] one k-octet plane = [A_h0 ; A_h1] = 4 KB
// so a K chunk of 32 of either operand is one contiguous 16 KB run and the whole main loop is: one producer thread issuing
// cp.async.bulk into a 4-stage ring, one MMA thread issuing tcgen05.mma (A_h0*[B_h0;B_h1] as one N = 256 MMA into
// [main | correction] accumulator columns plus A_h1*B_h0 into the correction columns; result = main + 2^-11 * correction),
// no staging threads at all.  A CTA owns BOTH 128-row M tiles of a 128-column N tile (the B chunk is read
// once for 256 slots; 2 x 256 accumulator columns = the whole TMEM) and one K split; partials are deterministic.
// The N tiles of one K split form a thread-block cluster: every CTA fetches 1/CL of the shared A chunk and MULTICASTS it
// into all CL CTAs' stages (cp.async.bulk ... .multicast::cluster), so X crosses the L2 -> SM fabric once per split instead
// of once per N tile; a stage is recycled when the MMAs of ALL CL CTAs have read it (tcgen05.commit multicast on the
// empty barriers).
#include "common.cuh"
#include "forward.cuh"
#include "tc05.cuh"

using namespace tc05;

int g_dne_theta_mc = 0;
namespace {
constexpr int TGM_KC = 32, TGM_STAGES = 4;                 // 32 k = four k-octet planes per chunk
constexpr int TGM_PLANE = 256 * 16;                         // bytes of one k-quad plane: 128 hi rows + 128 lo rows
constexpr int TGM_CHUNK = (TGM_KC / 8) * TGM_PLANE;         // 16 KB per operand tile and chunk
constexpr int TGM_EPI_WARPS = 8;
constexpr int TGM_THREADS = 32 * (2 + TGM_EPI_WARPS);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// multicast variant of bulk_g2s: the bytes land at the same shared-memory offset in every CTA of cta_mask, each CTA's
// mbarrier (same offset) receives the complete_tx for its own copy
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

template <int MT, int CL>
__global__ void __launch_bounds__(TGM_THREADS, 1)
theta_gemm_tma_kernel(const float* __restrict__ Xc, const float* __restrict__ Wc, int M, int N, int KQ, int chunks_per_split,
                      int n_chunks, float* __restrict__ part) {
    constexpr int STAGE = (MT + 1) * TGM_CHUNK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    __shared__ uint64_t full_bar[TGM_STAGES], empty_bar[TGM_STAGES], done_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntile = blockIdx.x, split = blockIdx.y, mgroup = blockIdx.z;
    const int c0 = split * chunks_per_split, c1 = min(n_chunks, c0 + chunks_per_split);
    const int nc = max(0, c1 - c0);
    constexpr int TCOLS = MT * 256;

    pdl_trigger();                                          // common.cuh: PDL chain of the tick
    if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
    if (tid == 32) {
        for (int i = 0; i < TGM_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CL); }
        mbar_init(&done_bar, 1);
        fence_mbar_init();
    }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    if (CL > 1) cluster_sync_all();                         // every peer's barriers exist before anyone multicasts into it
    const uint32_t tmem_base = tmem_base_s;
    constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1);
    pdl_wait();                                             // Xc is written by the previous kernel (conv3 epilogue); part is read by the next

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint8_t* xa = (const uint8_t*)Xc + (size_t)(mgroup * MT) * KQ * TGM_PLANE;
            const uint8_t* wb = (const uint8_t*)Wc + (size_t)ntile * KQ * TGM_PLANE;
            for (int i = 0; i < nc; ++i) {
                const int st = i % TGM_STAGES;
                mbar_wait(&empty_bar[st], ((i / TGM_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&full_bar[st], STAGE);
                const size_t koff = (size_t)(c0 + i) * TGM_CHUNK;
                uint8_t* dst = smem + st * STAGE;
                if (CL == 1) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        bulk_g2s(dst + mt * TGM_CHUNK, xa + (size_t)mt * KQ * TGM_PLANE + koff, TGM_CHUNK, &full_bar[st]);
                } else {
                    // this CTA's 1/CL slice of the A region of the stage, delivered to every CTA of the cluster
                    constexpr int SLICE = MT * TGM_CHUNK / CL;
                    const int off = (int)cluster_ctarank() * SLICE, mt = off / TGM_CHUNK, in_chunk = off % TGM_CHUNK;
                    bulk_g2s_mc(dst + off, xa + (size_t)mt * KQ * TGM_PLANE + koff + in_chunk, SLICE, &full_bar[st], CL_MASK);
                }
                bulk_g2s(dst + MT * TGM_CHUNK, wb + koff, TGM_CHUNK, &full_bar[st]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (converged warp, one elected lane) =====
        constexpr uint32_t IDESC2 = idesc_f16(128, 256), IDESC1 = idesc_f16(128, 128);
        const uint64_t d0 = smem_desc(smem_u32(smem), TGM_PLANE, 128);
        for (int i = 0; i < nc; ++i) {
            const int st = i % TGM_STAGES;
            mbar_wait(&full_bar[st], (i / TGM_STAGES) & 1);
            fence_after_thread_sync();
            if (elect_one()) {
                const uint64_t ds = d0 + (uint64_t)((st * STAGE) >> 4);
#pragma unroll
                for (int k16 = 0; k16 < TGM_KC / 16; ++k16) {
                    const uint64_t dB = ds + (uint64_t)((MT * TGM_CHUNK + 2 * k16 * TGM_PLANE) >> 4);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint64_t dAh = ds + (uint64_t)((mt * TGM_CHUNK + 2 * k16 * TGM_PLANE) >> 4);
                        const uint32_t d = tmem_base + mt * 256;
                        mma_f16(d, dAh, dB, IDESC2, (i | k16) != 0);                    // A_h0 * [B_h0 ; B_h1] -> [main | correction]
                        mma_f16(d + 128, dAh + (uint64_t)(2048 >> 4), dB, IDESC1, 1);    // A_h1 * B_h0 -> correction
                    }
                }
                if (CL == 1) mma_commit(&empty_bar[st]);
                else mma_commit_mc(&empty_bar[st], CL_MASK);             // the stage of EVERY peer holds data this CTA multicast
                if (i == nc - 1) mma_commit(&done_bar);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: part[split][m][n] = D[:, n] + 2^-11 * D[:, 128 + n] =====
        const int ew = warp - 2, lq = warp & 3, half = ew >> 2;       // TMEM lane quarter = warp % 4 (hardware rule)
        if (nc > 0) {
            mbar_wait(&done_bar, 0);
            fence_after_thread_sync();
        }
        float* P = part + (int64_t)split * M * N;
#pragma unroll 1
        for (int q = half; q < MT * 8; q += 2) {
            const int mt = q >> 3, j = q & 7;
            float v[16], v2[16];
            if (nc > 0) {
                __syncwarp();
                const uint32_t t = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(mt * 256 + j * 16);
                tmem_ld16_async(t, v);
                tmem_ld16_async(t + 128, v2);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int x = 0; x < 16; ++x) v[x] = v2[x] = 0.0f;
            }
            const int m = (mgroup * MT + mt) * 128 + lq * 32 + lane;
            const int n = ntile * 128 + j * 16;
            if (m < M) {
#pragma unroll
                for (int x = 0; x < 16; x += 4) {
                    if (n + x + 3 < N)
                        *reinterpret_cast<float4*>(P + (int64_t)m * N + n + x) =
                            make_float4(fmaf(v2[x], F16_LO_INV, v[x]), fmaf(v2[x + 1], F16_LO_INV, v[x + 1]),
                                        fmaf(v2[x + 2], F16_LO_INV, v[x + 2]), fmaf(v2[x + 3], F16_LO_INV, v[x + 3]));
                    else
                        for (int y = 0; y < 4; ++y)
                            if (n + x + y < N) P[(int64_t)m * N + n + x + y] = fmaf(v2[x + y], F16_LO_INV, v[x + y]);
                }
            }
        }
    }
    fence_before_thread_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                         // no CTA leaves while a peer may still arrive on its barriers
    if (warp == 0) tmem_dealloc(tmem_base, TCOLS);
}

// W[K][N] (row-major, arbitrary element alignment) -> Wc[n tile][k octet][h0 | h1][128][8 x fp16]; columns >= N are zero
__global__ void theta_prep_kernel(const float* __restrict__ W, int K, int N, int KO, float* __restrict__ Wc) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (ntile, ko, n)
    const int n_tiles = (N + 127) / 128;
    if (u >= (int64_t)n_tiles * KO * 128) return;
    const int nl = (int)(u % 128);
    const int ko = (int)((u / 128) % KO), nt = (int)(u / ((int64_t)128 * KO));
    const int n = nt * 128 + nl;
    float w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = (n < N && 8 * ko + j < K) ? W[(int64_t)(8 * ko + j) * N + n] : 0.0f;
    uint4 hi, lo;
    split_f16x8(w, hi, lo);
    uint4* dst = reinterpret_cast<uint4*>(Wc) + ((int64_t)nt * KO + ko) * 256 + nl;
    dst[0] = hi;
    dst[128] = lo;
}
}  // namespace

// ---- host interface (forward.cuh) ---------------------------------------------------------------------------------
size_t dne_tgm_xc_bytes(int n_slots, int K) { return (size_t)((n_slots + 127) / 128) * (K / 8) * TGM_PLANE; }
size_t dne_tgm_wc_bytes(int K, int N) { return (size_t)((N + 127) / 128) * (K / 8) * TGM_PLANE; }
bool dne_tgm_supported(int K, int N, int k_per_split) { return K % TGM_KC == 0 && N % 4 == 0 && k_per_split % TGM_KC == 0; }

int dne_launch_theta_prep(const float* W, int K, int N, float* Wc, cudaStream_t st) {
    const int KO = K / 8;
    const int64_t units = (int64_t)((N + 127) / 128) * KO * 128;
    theta_prep_kernel<<<(unsigned)((units + 255) / 256), 256, 0, st>>>(W, K, N, KO, Wc);
    DNE_LAUNCHED(1);
    return 0;
}

template <int MT, int CL>
static int launch_tgm(const float* Xc, const float* Wc, int M, int N, int KQ, int cps, int n_chunks, int n_split, int m_groups,
                      int n_tiles, float* part, cudaStream_t st) {
    constexpr int SMEM = TGM_STAGES * (MT + 1) * TGM_CHUNK + 256;
    auto kern = theta_gemm_tma_kernel<MT, CL>;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_done[64] = {};
    if (dev < 64 && !attr_done[dev]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return DNE_ERR_CUDA;
        attr_done[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles, n_split, m_groups);          // blockIdx.x = N tile (= rank in the cluster when CL > 1)
    cfg.blockDim = dim3(TGM_THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_dne_pdl ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, kern, Xc, Wc, M, N, KQ, cps, n_chunks, part) != cudaSuccess) return DNE_ERR_CUDA;
    DNE_LAUNCHED(1);
    return 0;
}

int dne_launch_theta_gemm_tma(const float* Xc, const float* Wc, int M, int K, int N, int k_per_split, int n_split, float* part,
                              cudaStream_t st) {
    if (!dne_tgm_supported(K, N, k_per_split)) return DNE_ERR_UNSUP;
    const int KQ = K / 8, n_chunks = K / TGM_KC, cps = k_per_split / TGM_KC;      // KQ: k-octet planes
    const int m_tiles = (M + 127) / 128, n_tiles = (N + 127) / 128;
    if (m_tiles > 1 && (m_tiles & 1)) return DNE_ERR_UNSUP;          // M tiles come in pairs (one CTA owns two) or alone
#define TGM_ARGS Xc, Wc, M, N, KQ, cps, n_chunks, n_split, (m_tiles >= 2 ? m_tiles / 2 : 1), n_tiles, part, st
    // cluster multicast of the A chunk (g_dne_theta_mc, dne_set_option("theta_mc", 1)) measured SLOWER on B200 (32 us vs
    // 18 us for the LargeModel fc at 256 slots: the 4-CTA cluster runs in lock step and must be co-scheduled): off
    if (g_dne_theta_mc && m_tiles >= 2 && n_tiles == 4) return launch_tgm<2, 4>(TGM_ARGS);
    if (g_dne_theta_mc && m_tiles >= 2 && n_tiles == 2) return launch_tgm<2, 2>(TGM_ARGS);
    if (m_tiles >= 2) return launch_tgm<2, 1>(TGM_ARGS);
    return launch_tgm<1, 1>(TGM_ARGS);
#undef TGM_ARGS
}
