// preprocess_kernels.cu -- the 210x160 -> 84x84 "warp" of the Atari observation pipeline, both reference flavours.
//
// CPU path  (es_distributed/atari_wrappers.py:105,138-142): max over the last two RGB frames (MaxAndSkipEnv) ->
//           gray = r*0.299 + g*0.587 + b*0.114 in float32 -> PIL Image.resize((84, 84), BILINEAR) on the float image ->
//           uint8 (truncation).  Pillow's BILINEAR is an area-scaled triangle filter when shrinking (support = in/out), two
//           passes (horizontal, then vertical), double accumulation, float32 intermediate: restated here bit for bit
//           (oracle.resize_pillow_bilinear is pinned against Pillow itself).
// GPU path  (gpu_implementation/gym_tensorflow/atari/tf_atari.py:88-92,149): NTSC palette index -> gray float32 LUT,
//           max over the two raw frames, tf.image.resize_bilinear(align_corners=True) -> float32 in [0, 1].
// HBM-bound byte work: one CTA per frame, the 2 x 100 KB (RGB) / 2 x 33 KB (palette) raw frames are read once (the
// horizontal pass's <= 5 overlapping taps hit L1/L2), the 84x84 result written once.
#include "common.cuh"

namespace {
constexpr int RAW_H = 210, RAW_W = 160, RES = 84;
constexpr int KX = 5, KY = 7;                     // ceil(160/84)*2+1, ceil(210/84)*2+1 (Pillow precompute_coeffs ksize)

struct WarpCoeffs {
    int xmin[RES], xcnt[RES], ymin[RES], ycnt[RES];
    double kx[RES * KX], ky[RES * KY];
};
__constant__ WarpCoeffs c_warp;

// Pillow src/libImaging/Resample.c precompute_coeffs for the bilinear (triangle) filter, same double arithmetic.
void pillow_coeffs(int in_size, int out_size, int ksize, int* xmins, int* counts, double* kk) {
    double scale = (double)in_size / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 1.0 * filterscale, ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        double* k = kk + xx * ksize;
        for (int x = 0; x < ksize; ++x) k[x] = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            const double w = a < 1.0 ? 1.0 - a : 0.0;
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        xmins[xx] = xmin;
        counts[xx] = xmax;
    }
}

int upload_coeffs() {
    int dev = 0;
    cudaGetDevice(&dev);
    static bool done[64] = {};
    if (dev < 64 && done[dev]) return 0;
    WarpCoeffs h;
    pillow_coeffs(RAW_W, RES, KX, h.xmin, h.xcnt, h.kx);
    pillow_coeffs(RAW_H, RES, KY, h.ymin, h.ycnt, h.ky);
    if (cudaMemcpyToSymbol(c_warp, &h, sizeof(h)) != cudaSuccess) return -1;
    if (dev < 64) done[dev] = true;
    return 0;
}

// gray of the per-channel max of two RGB frames (atari_wrappers.py:105,139); canonical float32 evaluation order
__device__ __forceinline__ float gray_max_rgb(const uint8_t* a, const uint8_t* b, int p) {
    const float r = (float)max(a[3 * p], b[3 * p]), g = (float)max(a[3 * p + 1], b[3 * p + 1]),
                bl = (float)max(a[3 * p + 2], b[3 * p + 2]);
    return __fadd_rn(__fadd_rn(__fmul_rn(r, 0.299f), __fmul_rn(g, 0.587f)), __fmul_rn(bl, 0.114f));
}

__global__ void __launch_bounds__(256)
warp_rgb_kernel(const uint8_t* __restrict__ raw, uint8_t* __restrict__ out, int n) {
    extern __shared__ float tmp[];                                // [210][84] float32 (Pillow's intermediate image)
    const int f = blockIdx.x;
    if (f >= n) return;
    const uint8_t* fa = raw + (int64_t)f * 2 * RAW_H * RAW_W * 3;
    const uint8_t* fb = fa + RAW_H * RAW_W * 3;
    for (int o = threadIdx.x; o < RAW_H * RES; o += blockDim.x) { // horizontal pass
        const int y = o / RES, xx = o - y * RES;
        const int x0 = c_warp.xmin[xx], cnt = c_warp.xcnt[xx];
        double ss = 0.0;
        for (int x = 0; x < cnt; ++x)
            ss = __dadd_rn(ss, __dmul_rn((double)gray_max_rgb(fa, fb, y * RAW_W + x0 + x), c_warp.kx[xx * KX + x]));
        tmp[o] = (float)ss;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < RES * RES; o += blockDim.x) {   // vertical pass + truncating uint8 cast
        const int yy = o / RES, xx = o - yy * RES;
        const int y0 = c_warp.ymin[yy], cnt = c_warp.ycnt[yy];
        double ss = 0.0;
        for (int y = 0; y < cnt; ++y) ss = __dadd_rn(ss, __dmul_rn((double)tmp[(y0 + y) * RES + xx], c_warp.ky[yy * KY + y]));
        const float v = (float)ss;
        out[(int64_t)f * RES * RES + o] = (uint8_t)(int)v;          // np.array(image, dtype=np.uint8): values lie in [0, 255]
    }
}

__global__ void __launch_bounds__(256)
warp_palette_kernel(const uint8_t* __restrict__ raw, const float* __restrict__ lut, float* __restrict__ out_f,
                    uint8_t* __restrict__ out_u8, int n) {
    __shared__ float s_lut[256];
    const int f = blockIdx.x;
    if (f >= n) return;
    s_lut[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    const uint8_t* fa = raw + (int64_t)f * 2 * RAW_H * RAW_W;
    const uint8_t* fb = fa + RAW_H * RAW_W;
    const float sy = __fdiv_rn((float)(RAW_H - 1), (float)(RES - 1)), sx = __fdiv_rn((float)(RAW_W - 1), (float)(RES - 1));
    auto px = [&](int y, int x) { return fmaxf(s_lut[fa[y * RAW_W + x]], s_lut[fb[y * RAW_W + x]]); };   // tf_atari.py:90-91
    for (int o = threadIdx.x; o < RES * RES; o += blockDim.x) {
        const int yy = o / RES, xx = o - yy * RES;
        const float fy = __fmul_rn((float)yy, sy), fx = __fmul_rn((float)xx, sx);
        const int ylo = (int)fy, xlo = (int)fx;
        const int yhi = min(ylo + 1, RAW_H - 1), xhi = min(xlo + 1, RAW_W - 1);
        const float yl = __fsub_rn(fy, (float)ylo), xl = __fsub_rn(fx, (float)xlo);
        const float tl = px(ylo, xlo), tr = px(ylo, xhi), bl = px(yhi, xlo), br = px(yhi, xhi);
        const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
        const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
        const float v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
        if (out_f) out_f[(int64_t)f * RES * RES + o] = v;
        if (out_u8) out_u8[(int64_t)f * RES * RES + o] = (uint8_t)__float2int_rn(fminf(fmaxf(__fmul_rn(v, 255.0f), 0.0f), 255.0f));
    }
}
}  // namespace

extern "C" int dne_warp_atari_rgb(const uint8_t* d_raw, uint8_t* d_out, int n_frames, void* stream) {
    DNE_CHECK_ARG(d_raw && d_out && n_frames >= 0, "bad arguments");
    if (n_frames == 0) return DNE_OK;
    if (upload_coeffs()) { dne_set_error("dne_warp_atari_rgb: coefficient upload failed"); return DNE_ERR_CUDA; }
    const int smem = RAW_H * RES * (int)sizeof(float);
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_done[64] = {};
    if (dev < 64 && !attr_done[dev]) {
        DNE_CUDA(cudaFuncSetAttribute(warp_rgb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done[dev] = true;
    }
    warp_rgb_kernel<<<n_frames, 256, smem, (cudaStream_t)stream>>>(d_raw, d_out, n_frames);
    DNE_LAUNCH_CHECK1();
    return DNE_OK;
}

extern "C" int dne_warp_atari_palette(const uint8_t* d_raw, const float* d_gray_lut, float* d_out_f32, uint8_t* d_out_u8,
                                      int n_frames, void* stream) {
    DNE_CHECK_ARG(d_raw && d_gray_lut && (d_out_f32 || d_out_u8) && n_frames >= 0, "bad arguments");
    if (n_frames == 0) return DNE_OK;
    warp_palette_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>(d_raw, d_gray_lut, d_out_f32, d_out_u8, n_frames);
    DNE_LAUNCH_CHECK1();
    return DNE_OK;
}
