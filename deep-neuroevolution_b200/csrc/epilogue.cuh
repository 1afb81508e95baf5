// epilogue.cuh -- per-slot helpers shared by the forward kernels: slot lookup, member-weight perturbation, and the
// fused per-output-channel epilogue (bias, virtual batch norm, activation).
#pragma once
#include "common.cuh"
#include "forward.cuh"

__device__ __forceinline__ bool slot_active(const SlotArgs& a, int slot) { return !a.active || a.active[slot]; }
__device__ __forceinline__ const float* slot_theta(const SlotArgs& a, int slot) {
    return a.theta + (a.theta_idx ? (int64_t)a.theta_idx[slot] * a.P : 0);
}
// member weight, exactly as the reference materialises it (es.py:413-419): v = fl(s*n); w = fl(theta + v)
__device__ __forceinline__ float perturbed(float th, float s, float n) { return __fadd_rn(th, __fmul_rn(s, n)); }

// per-output-channel affine (bias, batch-norm) + activation for one slot
struct ChanEpi {
    float bias, mean, inv, gamma, beta;
    int bn, act;
    __device__ __forceinline__ float apply(float acc) const {
        float y = acc + bias;
        if (bn != DNE_BN_NONE) y = (y - mean) * inv * gamma + beta; // policies.py:322 (eps 1e-3, decay 0); batchnorm.py:85-93
        return apply_act(y, act);
    }
};
__device__ __forceinline__ ChanEpi make_chan_epi(const SlotArgs& sa, const LayerEpi& e, int slot, int cout, int n,
                                                 const float* th, int64_t idx, float s) {
    ChanEpi c;
    c.bn = e.bn;
    c.act = e.act;
    c.bias = (e.off_b >= 0) ? perturbed(th[e.off_b + n], s, sa.noise[idx + e.off_b + n]) : 0.0f;
    c.mean = 0.f; c.inv = 1.f; c.gamma = 1.f; c.beta = 0.f;
    if (e.bn != DNE_BN_NONE) {
        const float* st = e.vbn + (int64_t)slot * e.vbn_len + e.bn_off;
        c.mean = st[n];
        c.inv = __fdiv_rn(1.0f, __fsqrt_rn(st[cout + n] + 1e-3f));
        if (e.bn == DNE_BN_TF) {
            c.gamma = perturbed(th[e.off_gamma + n], s, sa.noise[idx + e.off_gamma + n]);
            c.beta = perturbed(th[e.off_beta + n], s, sa.noise[idx + e.off_beta + n]);
        } else {                           // DNE_BN_GPU: the layer has no bias; its 'b' is added after the normalisation
            c.beta = c.bias;
            c.bias = 0.0f;
        }
    }
    return c;
}
