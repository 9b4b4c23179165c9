// loss_common.cuh -- per-ray terms of training/loss.py:34-198, shared by the stand-alone loss kernels (optim.cu) and the
// compositing kernels that carry the loss (composite.cu: vn_composite_loss_fwd / _bwd).
#pragma once
#include "common.cuh"

struct RayLoss { float dc[3]; float e_uss, e_tof, e_rgbd; bool v_uss, v_tof, v_rgbd; };

__device__ __forceinline__ RayLoss ray_loss(const float* __restrict__ rgb, const float* __restrict__ opacity,
                                            const float* __restrict__ depth, const float* __restrict__ gt_rgb,
                                            const float* __restrict__ uss, const float* __restrict__ tof,
                                            const float* __restrict__ rgbd, int64_t n, float bg, float uss_tol) {
    RayLoss r;
    const float op = __ldcg(opacity + n), d = __ldcg(depth + n);   // predecessor's outputs: coherent loads (common.cuh, PDL)
#pragma unroll
    for (int c = 0; c < 3; ++c) r.dc[c] = (__ldcg(rgb + 3 * n + c) + bg * (1.0f - op)) - __ldg(gt_rgb + 3 * n + c);   // rendering.py:225
    r.v_uss = r.v_tof = r.v_rgbd = false;
    r.e_uss = r.e_tof = r.e_rgbd = 0.0f;
    if (uss) { const float m = __ldg(uss + n); r.v_uss = !isnan(m) && (d < m - uss_tol); if (r.v_uss) r.e_uss = d - m; }   // loss.py:186-194
    if (tof) { const float m = __ldg(tof + n); r.v_tof = !isnan(m); if (r.v_tof) r.e_tof = d - m; }                        // loss.py:140-141
    if (rgbd) { const float m = __ldg(rgbd + n); r.v_rgbd = !isnan(m); if (r.v_rgbd) r.e_rgbd = d - m; }                   // loss.py:118-119
    return r;
}


// the same from values held in registers (the compositor's own outputs of this ray)
__device__ __forceinline__ RayLoss ray_loss_vals(float c0, float c1, float c2, float op, float d, const float* __restrict__ gt_rgb,
                                                 const float* __restrict__ uss, const float* __restrict__ tof,
                                                 const float* __restrict__ rgbd, int64_t n, float bg, float uss_tol) {
    RayLoss r;
    const float c[3] = {c0, c1, c2};
#pragma unroll
    for (int k = 0; k < 3; ++k) r.dc[k] = (c[k] + bg * (1.0f - op)) - __ldg(gt_rgb + 3 * n + k);   // rendering.py:225
    r.v_uss = r.v_tof = r.v_rgbd = false;
    r.e_uss = r.e_tof = r.e_rgbd = 0.0f;
    if (uss) { const float m = __ldg(uss + n); r.v_uss = !isnan(m) && (d < m - uss_tol); if (r.v_uss) r.e_uss = d - m; }   // loss.py:186-194
    if (tof) { const float m = __ldg(tof + n); r.v_tof = !isnan(m); if (r.v_tof) r.e_tof = d - m; }                        // loss.py:140-141
    if (rgbd) { const float m = __ldg(rgbd + n); r.v_rgbd = !isnan(m); if (r.v_rgbd) r.e_rgbd = d - m; }                   // loss.py:118-119
    return r;
}

// everything the compositing kernels need to carry the loss
struct LossArgs {
    const float* gt_rgb; const float* uss; const float* tof; const float* rgbd;   // targets (NaN = no measurement; may be NULL)
    float bg, uss_tol;
    float* sums; float* counts;                                                   // [4] each: colour, USS, ToF, RGBD
    float w_color, w_uss, w_tof, w_rgbd;                                          // (backward)
    const float* scale_dev;                                                       // GradScaler scale or NULL (backward)
    float* loss_out;                                                              // [1] or NULL (backward)
};
