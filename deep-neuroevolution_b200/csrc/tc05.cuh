// tc05.cuh -- minimal tcgen05 / TMEM / mbarrier PTX wrappers for sm_100a (hand-written; no CUTLASS dependency).
//
// Operand layout used throughout (UMMA "K-major, no swizzle / interleave" canonical layout, 32-bit elements):
// a tile of R rows x KC k-elements is stored as float4 T[KC/4][R]:  element (r, k) lives at byte
//     (k/4) * (R*16)  +  r*16  +  (k%4)*4
// i.e. core matrices of 8 rows x 16 bytes are contiguous (128 B), the next 8-row group follows at SBO = 128 B and the
// next 4 k-elements at LBO = R*16 B.  One tcgen05.mma kind::tf32 consumes K = 8 elements = two such k-quads.
// Consecutive rows are consecutive 16-byte chunks, so staging threads that own consecutive rows write conflict-free.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}

// ---- fences ------------------------------------------------------------------------------------------
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp) -----------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors -------------------------------------------------------------------------------------
// shared-memory matrix descriptor (sm_100 version 1), K-major, SWIZZLE_NONE
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    return d;                        // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor: kind::tf32, fp32 accumulate, A and B K-major, dense
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4)                      // c_format = F32
           | (2u << 7) | (2u << 10)       // a_format = b_format = TF32
           | ((uint32_t)(N >> 3) << 17)   // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// instruction descriptor: kind::f16 with fp16 operands, fp32 accumulate, A and B K-major, dense (K = 16 per instruction)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4)                      // c_format = F32
                                          // a_format = b_format = F16 (0)
           | ((uint32_t)(N >> 3) << 17)   // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// Explicit shared-window accesses with 32-bit addresses.  Stores through a C++ pointer derived from the (re-aligned)
// dynamic shared-memory base compile to GENERIC ST.E / LD.E (the cast hides the address space from ptxas).
__device__ __forceinline__ void sts128(uint32_t saddr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}

// One lane of the (converged) warp is elected.  Issue pattern for tcgen05.mma / tcgen05.commit: the WHOLE warp runs the
// loop and computes the descriptors (warp-uniform values -> uniform registers), and only the elected lane executes the
// instruction.  Issuing from inside `if (lane == 0)` instead makes ptxas wrap every MMA in an ELECT/R2UR waterfall
// loop (~15 instructions, ~100 cycles per MMA) -- measured in tools/probe_mma.py.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns ------------------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait: issue several, then tmem_ld_wait() once before the registers are read
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
          "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- 3xTF32 operand split ---------------------------------------------------------------------------
// hi = rna_tf32(x), lo = rna_tf32(x - hi): both have their low 13 mantissa bits clear, so the tensor core's fp32->tf32
// input truncation is exact.  a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b  (dropped lo*lo term ~2^-24 relative).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    uint32_t l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}

// Cheap variant used by the staging loops (4 integer/float ops instead of two cvt.rna sequences):
// hi = x rounded to TF32 by adding half an ulp to the magnitude bits and masking, lo = (x - hi) truncated to TF32.
// |lo| <= 2^-11 |x| and its truncation loses <= 2^-11 of it, so the dropped part is <= 2^-22 |x| per operand.
__device__ __forceinline__ void split_tf32_fast(float x, float& hi, float& lo) {
    const uint32_t hb = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    hi = __uint_as_float(hb);
    lo = __uint_as_float(__float_as_uint(x - hi) & 0xFFFFE000u);
}

// ---- 2 x fp16 operand split (kind::f16 runs at twice the MAC rate of kind::tf32) ----------------------------------------
// x = h0 + h1 * 2^-11 with h0 = rn_fp16(x), h1 = rn_fp16((x - h0) * 2^11): 22 significand bits, |error| <= 2^-24 |x| (fp16
// subnormals are exact to 2^-25 absolute and h1 picks up the rest).  The lo part is carried SCALED by 2^11 so that it never
// falls into fp16's subnormal range before x itself does; products with it go to a separate accumulator that the epilogue
// folds back with 2^-11:  a*b ~= h0a*h0b + 2^-11 * (h0a*h1b + h1a*h0b)   (dropped h1a*h1b term: 2^-24 relative).
// fp16 products are exact in the fp32 accumulator.  Range: |x| <= 65504 (fp16 max); larger magnitudes saturate.
constexpr float F16_LO_SCALE = 2048.0f, F16_LO_INV = 1.0f / 2048.0f;
__device__ __forceinline__ void split_f16(float x, __half& h0, __half& h1) {
    h0 = __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f));
    h1 = __float2half_rn((x - __half2float(h0)) * F16_LO_SCALE);
}
// two values at a time with the packed conversions (cvt.rn.satfinite.f16x2.f32: saturating, one instruction per pair)
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));      // d = {hi half: a, lo half: b}
    const __half2 h = *reinterpret_cast<const __half2*>(&hi);
    const float2 hf = __half22float2(h);
    const float r0 = (x0 - hf.x) * F16_LO_SCALE, r1 = (x1 - hf.y) * F16_LO_SCALE;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
// eight values -> one 16-byte row of the hi plane and one of the (scaled) lo plane
__device__ __forceinline__ void split_f16x8(const float (&x)[8], uint4& hi, uint4& lo) {
    split_f16x2(x[0], x[1], hi.x, lo.x);
    split_f16x2(x[2], x[3], hi.y, lo.y);
    split_f16x2(x[4], x[5], hi.z, lo.z);
    split_f16x2(x[6], x[7], hi.w, lo.w);
}
__device__ __forceinline__ void sts128u(uint32_t saddr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace tc05

// ---- 1-D bulk async copy (TMA engine, no tensor map): global -> shared, completion on an mbarrier ---------------
namespace tc05 {
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2-only prefetch of a contiguous global range (16-byte aligned, size % 16 == 0): no shared-memory destination
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
}  // namespace tc05
