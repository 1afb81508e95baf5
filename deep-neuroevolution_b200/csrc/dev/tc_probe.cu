// tc_probe.cu -- micro-probe of tcgen05.mma kind::tf32 issue/throughput for different shared-memory operand layouts
// (one CTA, operands staged once, `reps` MMA batches accumulated; cycles measured with clock64 around the batch).
// layout 0: K-major no-swizzle (interleave): float4 T[KC/4][rows]           (LBO = rows*16, SBO = 128)
// layout 1: K-major SWIZZLE_32B : 8-row x 32 B atoms, k-block stride rows*32 (SBO = 256)
// layout 2: K-major SWIZZLE_128B: 8-row x 128 B atoms (KC = 32 per row)      (SBO = 1024), K advance = +32 B
#include "dev.cuh"
#include "../tc05.cuh"
using namespace tc05;

__device__ __forceinline__ uint64_t desc_layout(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

template <int N>
__global__ void __launch_bounds__(128) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ C, int layout, int reps,
                                                       long long* __restrict__ cycles) {
    constexpr int KC = 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // 128 x 32 floats = 16 KB
    uint8_t* sB = smem + 16384;               // N x 32 floats
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int TCOLS = N < 32 ? 32 : N;
    if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
    if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
    auto addr = [&](int rows, int r, int q) -> uint32_t {        // byte offset of (row r, k-quad q) in a rows x 32 tile
        if (layout == 0) return q * rows * 16 + r * 16;
        if (layout == 1) return (q >> 1) * rows * 32 + (r >> 3) * 256 + (r & 7) * 32 + (((q & 1) ^ ((r & 7) >> 2)) * 16);
        return (r >> 3) * 1024 + (r & 7) * 128 + ((q ^ (r & 7)) * 16);
    };
    for (int u = tid; u < 128 * 8; u += 128) {
        const int r = u % 128, q = u / 128;
        float4 v = *reinterpret_cast<const float4*>(A + r * KC + 4 * q);
        float4 hi, lo;
        split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(sA + addr(128, r, q)) = hi;
    }
    for (int u = tid; u < N * 8; u += 128) {
        const int n = u % N, q = u / N;
        float4 v = *reinterpret_cast<const float4*>(B + n * KC + 4 * q);
        float4 hi, lo;
        split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(sB + addr(N, n, q)) = hi;
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();
    fence_after_thread_sync();
    const uint32_t tmem_base = tmem_base_s;
    constexpr uint32_t IDESC = idesc_tf32(128, N);
    __shared__ uint64_t full_b, empty_b;
    __shared__ int stop_flag;
    if (tid == 0) { mbar_init(&full_b, 1); mbar_init(&empty_b, 1); fence_mbar_init(); stop_flag = 0; }
    __syncthreads();
    const int mode = layout >> 4;           // 0 chain, 1 commit+wait per 6 MMAs, 2 chain + generic-store traffic,
    layout &= 15;                           // 3 chain + st.shared traffic, 4 full/empty ping-pong with a staging warp
    // converged-warp issue (tc05.cuh: elect_one): descriptors are warp-uniform, one elected lane issues
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    uint64_t dA[4], dB[4];
#pragma unroll
    for (int k8 = 0; k8 < 4; ++k8) {
        if (layout == 0) {
            dA[k8] = desc_layout(a0 + 2 * k8 * 128 * 16, 128 * 16, 128, 0);
            dB[k8] = desc_layout(b0 + 2 * k8 * N * 16, N * 16, 128, 0);
        } else if (layout == 1) {
            dA[k8] = desc_layout(a0 + k8 * 128 * 32, 16, 256, 6);
            dB[k8] = desc_layout(b0 + k8 * N * 32, 16, 256, 6);
        } else {
            dA[k8] = desc_layout(a0 + k8 * 32, 16, 1024, 2);
            dB[k8] = desc_layout(b0 + k8 * 32, 16, 1024, 2);
        }
    }
    auto issue = [&](int n_mma, bool first) {          // n_mma is 4 or 6 (compile-time at the call sites)
        if (elect_one()) {
#pragma unroll
            for (int i = 0; i < 6; ++i)
                if (i < n_mma) mma_tf32(tmem_base, dA[i & 3], dB[i & 3], IDESC, !(first && i == 0));
        }
        __syncwarp();
    };
    uint8_t* scratch = sB + 16384;          // 16 KB scratch for the store-traffic modes
    if (warp == 0) {
        const long long t0 = clock64();
        if (mode == 0 || mode == 2 || mode == 3) {
            for (int rep = 0; rep < reps; ++rep) issue(4, rep == 0);
            if (elect_one()) mma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, 0);
        } else if (mode == 1) {
            for (int rep = 0; rep < reps; ++rep) {
                issue(6, rep == 0);
                if (elect_one()) mma_commit(&bar);
                __syncwarp();
                mbar_wait(&bar, rep & 1);
            }
        } else {
            for (int rep = 0; rep < reps; ++rep) {
                mbar_wait(&full_b, rep & 1);
                fence_after_thread_sync();
                issue(6, rep == 0);
                if (elect_one()) mma_commit(&empty_b);
                __syncwarp();
            }
            if (elect_one()) mma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, 0);
        }
        const long long t1 = clock64();
        if (lane == 0) { cycles[0] = t1 - t0; *(volatile int*)&stop_flag = 1; }
    } else if (warp >= 1 && (mode == 2 || mode == 3)) {
        const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        uint8_t* p = scratch + (tid - 32) * 16;
        const uint32_t pa = smem_u32(p);
        int guard = 0;
        while (*(volatile int*)&stop_flag == 0 && ++guard < (1 << 22)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (mode == 2) asm volatile("st.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + j * 1536), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                else asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(pa + j * 1536), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
        }
    } else if (warp == 1 && mode == 4) {
        for (int rep = 0; rep < reps; ++rep) {
            mbar_wait(&empty_b, (rep & 1) ^ 1);
            asm volatile("st.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(scratch + lane * 16), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_b);
        }
    }
    __syncthreads();
    fence_after_thread_sync();
#pragma unroll
    for (int j = 0; j < N / 16; ++j) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 16, v);
        const int m = warp * 32 + lane;
#pragma unroll
        for (int x = 0; x < 16; ++x) C[m * N + j * 16 + x] = v[x];
    }
    fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TCOLS);
}

// A [128,32], B [N,32] row-major; C[128,N] = reps * A * B^T (tf32-hi only); cycles[0] = clock64 ticks of the MMA batch
extern "C" int dne_probe_mma(const float* d_A, const float* d_B, float* d_C, int N, int layout, int reps,
                             long long* d_cycles, void* stream) {
    DNE_CHECK_ARG(d_A && d_B && d_C && d_cycles && (N == 32 || N == 64 || N == 128) && layout >= 0 && (layout & 15) <= 2 && reps >= 1,
                  "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = 16384 + 16384 + 16384 + 1024;
    if (N == 32) { cudaFuncSetAttribute(tc_probe_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                   tc_probe_kernel<32><<<1, 128, smem, st>>>(d_A, d_B, d_C, layout, reps, d_cycles); }
    else if (N == 64) { cudaFuncSetAttribute(tc_probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                   tc_probe_kernel<64><<<1, 128, smem, st>>>(d_A, d_B, d_C, layout, reps, d_cycles); }
    else { cudaFuncSetAttribute(tc_probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                   tc_probe_kernel<128><<<1, 128, smem, st>>>(d_A, d_B, d_C, layout, reps, d_cycles); }
    DNE_LAUNCH_CHECK1();
    return DNE_OK;
}
