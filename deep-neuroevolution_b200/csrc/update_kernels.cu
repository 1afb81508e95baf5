// update_kernels.cu -- generation update: centred ranks, ES gradient, Adam / SGD step.
//
//   dne_centered_rank  <- es_distributed/es.py:70-85
//   dne_es_grad        <- es_distributed/es.py:115-122, 291-296
//   dne_adam_step      <- es_distributed/optimizers.py:10-17, 35-50  (+ es.py:298)
//   dne_sgd_step       <- es_distributed/optimizers.py:23-32
//
// All three are HBM-bound streaming kernels (no tensor-core shaped work here):
//   es_grad reads n*4P bytes of noise (algorithmic), writes 4P; adam touches 7*4P bytes.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------
// centred ranks.  rank[i] = #{j : x[j] < x[i]} + #{j < i : x[j] == x[i]}  (stable ascending order, the
// canonical tie rule; NaN sorts last like numpy).  O(count^2) compares out of a shared-memory tile:
// count is 2n <= a few 10^4, so this is microseconds and keeps the kernel a single pass.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool key_less(float a, float b) { return (a < b) || ((b != b) && (a == a)); }
__device__ __forceinline__ bool key_eq(float a, float b) { return (a == b) || ((a != a) && (b != b)); }

constexpr int RANK_TILE = 1024;

__global__ void __launch_bounds__(256) centered_rank_kernel(const float* __restrict__ x, int count,
                                                            float* __restrict__ centered,
                                                            int32_t* __restrict__ ranks) {
    __shared__ float tile[RANK_TILE];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float xi = (i < count) ? x[i] : 0.0f;
    int r = 0;
    for (int base = 0; base < count; base += RANK_TILE) {
        const int len = min(RANK_TILE, count - base);
        __syncthreads();
        for (int t = threadIdx.x; t < len; t += blockDim.x) tile[t] = x[base + t];
        __syncthreads();
        if (i < count) {
#pragma unroll 8
            for (int t = 0; t < len; ++t) {
                const float xj = tile[t];
                r += (key_less(xj, xi) || (key_eq(xj, xi) && (base + t) < i)) ? 1 : 0;
            }
        }
    }
    if (i < count) {
        if (ranks) ranks[i] = r;
        // es.py:82-84: y = ranks.astype(f32); y /= (size-1); y -= .5   (two float32 roundings)
        if (centered) centered[i] = __fsub_rn(__fdiv_rn((float)r, (float)(count - 1)), 0.5f);
    }
}

extern "C" int dne_centered_rank(const float* d_returns, int count, float* d_centered, int32_t* d_ranks,
                                 void* stream) {
    DNE_CHECK_ARG(d_returns && count >= 0, "bad arguments");
    if (count == 0) return DNE_OK;
    centered_rank_kernel<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_returns, count, d_centered,
                                                                               d_ranks);
    DNE_LAUNCH_CHECK1();
    return DNE_OK;
}

// ---------------------------------------------------------------------------------------------------
// ES gradient.  g[j] = (1/denom) * sum_i w_i * noise[idx_i + j],  w_i = proc[i,0] - proc[i,1] (float32).
// Each thread owns JPT output elements (strided by blockDim so warp loads are contiguous 128 B runs --
// slices start at arbitrary element offsets, so the loads are scalar, coalesced, L1-bypassing) and keeps
// U slices in flight.  Accumulation is float64 with a single float32 rounding at the end: this is the
// float64 referee of the oracle, and costs nothing on an HBM-bound kernel (1 DFMA per 4 bytes).
// ---------------------------------------------------------------------------------------------------
constexpr int GRAD_THREADS = 256;
constexpr int GRAD_ITILE = 512;   // slices staged in shared memory per outer iteration

template <int JPT, int U>
__global__ void __launch_bounds__(GRAD_THREADS)
es_grad_kernel(const float* __restrict__ noise, const float* __restrict__ proc_n2,
               const int64_t* __restrict__ idx, int n, int64_t P, double inv_denom, float* __restrict__ g,
               int accumulate) {
    __shared__ float s_w[GRAD_ITILE];
    __shared__ int64_t s_idx[GRAD_ITILE];
    const int64_t j0 = (int64_t)blockIdx.x * (GRAD_THREADS * JPT) + threadIdx.x;
    double acc[JPT];
    bool ok[JPT];
#pragma unroll
    for (int c = 0; c < JPT; ++c) {
        acc[c] = 0.0;
        ok[c] = (j0 + (int64_t)c * GRAD_THREADS) < P;
    }
    for (int base = 0; base < n; base += GRAD_ITILE) {
        const int len = min(GRAD_ITILE, n - base);
        __syncthreads();
        for (int t = threadIdx.x; t < len; t += GRAD_THREADS) {
            s_w[t] = __fsub_rn(proc_n2[2 * (base + t)], proc_n2[2 * (base + t) + 1]);   // es.py:292
            s_idx[t] = idx[base + t];
        }
        __syncthreads();
        int t = 0;
        for (; t + U <= len; t += U) {
            float v[U][JPT];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float* p = noise + s_idx[t + u] + j0;
#pragma unroll
                for (int c = 0; c < JPT; ++c) v[u][c] = ok[c] ? ldg_stream_f1(p + c * GRAD_THREADS) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double w = (double)s_w[t + u];
#pragma unroll
                for (int c = 0; c < JPT; ++c) acc[c] = fma(w, (double)v[u][c], acc[c]);
            }
        }
        for (; t < len; ++t) {
            const float* p = noise + s_idx[t] + j0;
            const double w = (double)s_w[t];
#pragma unroll
            for (int c = 0; c < JPT; ++c)
                if (ok[c]) acc[c] = fma(w, (double)ldg_stream_f1(p + c * GRAD_THREADS), acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < JPT; ++c) {
        if (ok[c]) {
            const int64_t j = j0 + (int64_t)c * GRAD_THREADS;
            const float r = (float)(acc[c] * inv_denom);
            g[j] = accumulate ? __fadd_rn(g[j], r) : r;
        }
    }
}

extern "C" int dne_es_grad(dne_ctx* ctx, const float* d_proc_n2, const int64_t* d_noise_idx, int n, int64_t P,
                           double denom, float* d_g, int accumulate, void* stream) {
    DNE_CHECK_ARG(ctx && ctx->noise, "noise table not bound (dne_noise_bind)");
    DNE_CHECK_ARG(d_proc_n2 && d_noise_idx && d_g && n >= 0 && P > 0 && denom != 0.0, "bad arguments");
    DNE_CHECK_ARG(P <= ctx->noise_count, "P larger than the noise table");
    cudaStream_t st = (cudaStream_t)stream;
    const double inv = 1.0 / denom;
    // enough CTAs for >= ~4 waves at 8 CTAs/SM decides the per-thread width
    const int64_t ctas4 = cdiv64(P, (int64_t)GRAD_THREADS * 4);
    if (ctas4 >= (int64_t)ctx->sm_count * 16) {
        es_grad_kernel<4, 4><<<(unsigned)ctas4, GRAD_THREADS, 0, st>>>(ctx->noise, d_proc_n2, d_noise_idx, n, P,
                                                                      inv, d_g, accumulate);
    } else {
        const int64_t ctas1 = cdiv64(P, (int64_t)GRAD_THREADS);
        es_grad_kernel<1, 8><<<(unsigned)ctas1, GRAD_THREADS, 0, st>>>(ctx->noise, d_proc_n2, d_noise_idx, n, P,
                                                                      inv, d_g, accumulate);
    }
    DNE_LAUNCH_CHECK1();
    return DNE_OK;
}

// ---------------------------------------------------------------------------------------------------
// Optimizer steps.  Every float32 operation is an explicit round-to-nearest intrinsic so that nvcc cannot
// contract a*b+c into an FMA: the result is then bit-identical to numpy's float32 elementwise evaluation of
// optimizers.py:45-50 (numpy-1.12 semantics: python scalars are rounded to float32 before meeting the array).
// The two norms for the update ratio are accumulated in float64 per block, then reduced by one block in a
// fixed order (deterministic).
// ---------------------------------------------------------------------------------------------------
constexpr int OPT_THREADS = 256;

__device__ __forceinline__ void block_reduce2_store(double a, double b, double* out2) {
    __shared__ double sa[OPT_THREADS / 32], sb[OPT_THREADS / 32];
    a = warp_sum(a);
    b = warp_sum(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int i = 0; i < OPT_THREADS / 32; ++i) { ta += sa[i]; tb += sb[i]; }
        out2[0] = ta;
        out2[1] = tb;
    }
}

__global__ void __launch_bounds__(OPT_THREADS)
adam_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
            int64_t P, float l2, float neg_a, float b1, float b1c, float b2, float b2c, float eps,
            double* __restrict__ partial) {
    double s_step = 0.0, s_theta = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < P; i += (int64_t)gridDim.x * OPT_THREADS) {
        const float th = theta[i];
        const float gg = __fadd_rn(-g[i], __fmul_rn(l2, th));                                    // es.py:298
        const float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(b1c, gg));                     // optimizers.py:47
        const float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(b2c, __fmul_rn(gg, gg)));      // :48
        const float step = __fdiv_rn(__fmul_rn(neg_a, mi), __fadd_rn(__fsqrt_rn(vi), eps));      // :49
        m[i] = mi;
        v[i] = vi;
        theta[i] = __fadd_rn(th, step);                                                          // :15
        s_step += (double)step * (double)step;
        s_theta += (double)th * (double)th;
    }
    block_reduce2_store(s_step, s_theta, partial + 2 * blockIdx.x);
}

__global__ void __launch_bounds__(OPT_THREADS)
sgd_kernel(float* __restrict__ theta, float* __restrict__ v, const float* __restrict__ g, int64_t P, float l2,
           float mom, float momc, float neg_lr, double* __restrict__ partial) {
    double s_step = 0.0, s_theta = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < P; i += (int64_t)gridDim.x * OPT_THREADS) {
        const float th = theta[i];
        const float gg = __fadd_rn(-g[i], __fmul_rn(l2, th));
        const float vi = __fadd_rn(__fmul_rn(mom, v[i]), __fmul_rn(momc, gg));                   // optimizers.py:30
        const float step = __fmul_rn(neg_lr, vi);                                                // :31
        v[i] = vi;
        theta[i] = __fadd_rn(th, step);
        s_step += (double)step * (double)step;
        s_theta += (double)th * (double)th;
    }
    block_reduce2_store(s_step, s_theta, partial + 2 * blockIdx.x);
}

// one block; thread t sums partials t, t+256, ... then a fixed-order tree: deterministic for a given grid
__global__ void __launch_bounds__(OPT_THREADS)
ratio_finalize_kernel(const double* __restrict__ partial, int nblocks, float* __restrict__ ratio) {
    __shared__ double sa[OPT_THREADS], sb[OPT_THREADS];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += OPT_THREADS) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *ratio = (float)(sqrt(sa[0]) / sqrt(sb[0]));   // optimizers.py:14
}

static int opt_grid(const dne_ctx* ctx, int64_t P) {
    int64_t want = cdiv64(P, OPT_THREADS);
    int64_t cap = (int64_t)ctx->sm_count * 8;
    if (cap > DNE_SCRATCH_DOUBLES / 2) cap = DNE_SCRATCH_DOUBLES / 2;
    return (int)(want < cap ? want : cap);
}

extern "C" int dne_adam_step(dne_ctx* ctx, float* d_theta, float* d_m, float* d_v, const float* d_g, int64_t P,
                             double l2coeff, double stepsize, double beta1, double beta2, double epsilon, int t,
                             float* d_update_ratio, void* stream) {
    if (ctx) dne_prep_invalidate_theta(ctx, d_theta);
    DNE_CHECK_ARG(ctx && d_theta && d_m && d_v && d_g && P > 0 && t >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    // optimizers.py:46 in double (python floats), rounded to float32 where it meets the float32 array
    const double a = stepsize * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t));
    const int grid = opt_grid(ctx, P);
    adam_kernel<<<grid, OPT_THREADS, 0, st>>>(d_theta, d_m, d_v, d_g, P, (float)l2coeff, -(float)a, (float)beta1,
                                             (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                             (float)epsilon, ctx->scratch);
    DNE_LAUNCH_CHECK1();
    if (d_update_ratio) {
        ratio_finalize_kernel<<<1, OPT_THREADS, 0, st>>>(ctx->scratch, grid, d_update_ratio);
        DNE_LAUNCH_CHECK1();
    }
    return DNE_OK;
}

extern "C" int dne_sgd_step(dne_ctx* ctx, float* d_theta, float* d_v, const float* d_g, int64_t P, double l2coeff,
                            double stepsize, double momentum, float* d_update_ratio, void* stream) {
    if (ctx) dne_prep_invalidate_theta(ctx, d_theta);
    DNE_CHECK_ARG(ctx && d_theta && d_v && d_g && P > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = opt_grid(ctx, P);
    sgd_kernel<<<grid, OPT_THREADS, 0, st>>>(d_theta, d_v, d_g, P, (float)l2coeff, (float)momentum,
                                            (float)(1.0 - momentum), (float)(-stepsize), ctx->scratch);
    DNE_LAUNCH_CHECK1();
    if (d_update_ratio) {
        ratio_finalize_kernel<<<1, OPT_THREADS, 0, st>>>(ctx->scratch, grid, d_update_ratio);
        DNE_LAUNCH_CHECK1();
    }
    return DNE_OK;
}
