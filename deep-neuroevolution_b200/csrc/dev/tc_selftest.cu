// tc_selftest.cu -- self-test of the tcgen05 plumbing (descriptors, TMEM alloc/ld, mbarrier commit, 3xTF32 split) used by
// tests/test_gpu_tc.py.  Part of the DEV library libdne_dev.so, not of the product ABI (include/dne.h).
#include "dev.cuh"
#include "../tc05.cuh"
using namespace tc05;
constexpr int TG_THREADS = 256;

// =====================================================================================================
// Self-test of the tcgen05 plumbing: C[128,N] = A[128,K] * B[N,K]^T with the same staging / descriptor / 3xTF32 code
// path (single CTA).  Exposed by libdne_dev.so for tests/test_gpu_tc.py.
// =====================================================================================================
template <int N>
__global__ void __launch_bounds__(128) tc_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                           float* __restrict__ C, int K) {
    constexpr int KC = 32;
    constexpr int A_PLANE = 128 * 16, B_PLANE = N * 16;
    constexpr int TCOLS = N < 32 ? 32 : N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint8_t* sA_hi = smem;
    uint8_t* sA_lo = sA_hi + (KC / 4) * A_PLANE;
    uint8_t* sB_hi = sA_lo + (KC / 4) * A_PLANE;
    uint8_t* sB_lo = sB_hi + (KC / 4) * B_PLANE;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
    if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = tmem_base_s;
    constexpr uint32_t IDESC = idesc_tf32(128, N);
    int phase = 0;
    for (int k0 = 0; k0 < K; k0 += KC) {
        for (int u = tid; u < 128 * (KC / 4); u += 128) {
            const int r = u % 128, q = u / 128;
            const float4 v = *reinterpret_cast<const float4*>(A + (int64_t)r * K + k0 + 4 * q);
            float4 hi, lo;
            split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(sA_hi + q * A_PLANE + r * 16) = hi;
            *reinterpret_cast<float4*>(sA_lo + q * A_PLANE + r * 16) = lo;
        }
        for (int u = tid; u < N * (KC / 4); u += 128) {
            const int n = u % N, q = u / N;
            const float4 v = *reinterpret_cast<const float4*>(B + (int64_t)n * K + k0 + 4 * q);
            float4 hi, lo;
            split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(sB_hi + q * B_PLANE + n * 16) = hi;
            *reinterpret_cast<float4*>(sB_lo + q * B_PLANE + n * 16) = lo;
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            fence_after_thread_sync();
            for (int k8 = 0; k8 < KC / 8; ++k8) {
                const uint64_t dAh = smem_desc(smem_u32(sA_hi) + 2 * k8 * A_PLANE, A_PLANE, 128);
                const uint64_t dAl = smem_desc(smem_u32(sA_lo) + 2 * k8 * A_PLANE, A_PLANE, 128);
                const uint64_t dBh = smem_desc(smem_u32(sB_hi) + 2 * k8 * B_PLANE, B_PLANE, 128);
                const uint64_t dBl = smem_desc(smem_u32(sB_lo) + 2 * k8 * B_PLANE, B_PLANE, 128);
                mma_tf32(tmem_base, dAh, dBh, IDESC, (k0 | k8) != 0);
                mma_tf32(tmem_base, dAl, dBh, IDESC, 1);
                mma_tf32(tmem_base, dAh, dBl, IDESC, 1);
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, phase);            // synchronous version: wait before re-staging
        phase ^= 1;
    }
    fence_after_thread_sync();
#pragma unroll
    for (int j = 0; j < N / 16; ++j) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 16, v);
        const int m = warp * 32 + lane;
#pragma unroll
        for (int x = 0; x < 16; ++x) C[(int64_t)m * N + j * 16 + x] = v[x];
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TCOLS);
}

extern "C" int dne_test_tc_gemm(const float* d_A, const float* d_B, float* d_C, int K, int N, void* stream) {
    DNE_CHECK_ARG(d_A && d_B && d_C && K > 0 && K % 32 == 0 && (N == 16 || N == 32 || N == 64), "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = 2 * 8 * 128 * 16 + 2 * 8 * N * 16 + 128;
    if (N == 16) {
        DNE_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tc_gemm_test_kernel<16><<<1, 128, smem, st>>>(d_A, d_B, d_C, K);
    } else if (N == 32) {
        DNE_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tc_gemm_test_kernel<32><<<1, 128, smem, st>>>(d_A, d_B, d_C, K);
    } else {
        DNE_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tc_gemm_test_kernel<64><<<1, 128, smem, st>>>(d_A, d_B, d_C, K);
    }
    DNE_LAUNCH_CHECK1();
    return DNE_OK;
}

