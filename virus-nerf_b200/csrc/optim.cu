// optim.cu -- caller side of the path: GradScaler.unscale_ +
// inf check + torch.optim.Adam(eps=1e-15) step over the dense hash table and the MLP
// weights as ONE streaming pass each (training/trainer.py:49-57, 138-141).
#include "common.cuh"
#include "loss_common.cuh"

__global__ void __launch_bounds__(256) grad_check_kernel(const float4* __restrict__ g4, const float* __restrict__ g, int64_t n,
                                                         float* __restrict__ found_inf) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    const int64_t n4 = n / 4;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldcg(g4 + i);
        bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) bad |= !isfinite(g[n4 * 4 + threadIdx.x]);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.0f;
}

VN_API int vn_grad_check(const float* g, int64_t n, float* found_inf, void* stream) {
    VN_REQUIRE(n >= 0 && found_inf && (n == 0 || g), "vn_grad_check: bad arguments");
    VN_REQUIRE(vn_aligned(g, 16), "vn_grad_check: g must be 16-byte aligned");
    if (n == 0) return VN_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)vn_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    vn_launch_pdl(grad_check_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const float4*)g, g, n, found_inf);
    VN_CHECK_LAUNCH("grad_check_kernel");
    return VN_OK;
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, AdamCfg c,
                                                   const float* __restrict__ found_inf,
                                                   const float* __restrict__ scale_dev,
                                                   const float* __restrict__ opt_state) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    if (found_inf && __ldcg(found_inf) != 0.0f) return;   // GradScaler.step skips the optimizer step (flag set by the predecessor: coherent load)
    if (scale_dev) c.inv_scale = 1.0f / *scale_dev;
    if (opt_state) { c.step_size = opt_state[0]; c.bc2_sqrt = opt_state[1]; }   // bias corrections of the device-side step count
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 P = *(float4*)(p + i), M = *(float4*)(m + i), V = *(float4*)(v + i);
        const float4 Gd = __ldcg((const float4*)(g + i));
        adam1(P.x, Gd.x, M.x, V.x, c); adam1(P.y, Gd.y, M.y, V.y, c);
        adam1(P.z, Gd.z, M.z, V.z, c); adam1(P.w, Gd.w, M.w, V.w, c);
        *(float4*)(p + i) = P; *(float4*)(m + i) = M; *(float4*)(v + i) = V;
    } else {
        for (int64_t k = i; k < n; ++k) adam1(p[k], g[k], m[k], v[k], c);
    }
}

VN_API int vn_adam_config(double lr, double beta1, double beta2, double eps, int step, float* h_out6) {
    VN_REQUIRE(h_out6 != nullptr && step >= 1, "vn_adam_config: bad arguments");
    const AdamCfg c = vn_make_adam_cfg(lr, beta1, beta2, eps, step, 1.0f);
    h_out6[0] = c.beta2; h_out6[1] = c.omb1; h_out6[2] = c.omb2; h_out6[3] = c.eps; h_out6[4] = c.step_size; h_out6[5] = c.bc2_sqrt;
    return VN_OK;
}

VN_API int vn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float inv_scale, double lr, double beta1,
                        double beta2, double eps, int step, const float* found_inf, const float* scale_dev,
                        void* stream) {
    VN_REQUIRE(n >= 0 && step >= 1, "vn_adam_step: bad n/step");
    if (n == 0) return VN_OK;
    VN_REQUIRE(p && g && m && v, "vn_adam_step: null pointer");
    VN_REQUIRE(vn_aligned(p, 16) && vn_aligned(g, 16) && vn_aligned(m, 16) && vn_aligned(v, 16),
               "vn_adam_step: buffers must be 16-byte aligned");
    const AdamCfg c = vn_make_adam_cfg(lr, beta1, beta2, eps, step, inv_scale);
    VnProfScope prof(VN_K_ADAM, n, (cudaStream_t)stream);
    vn_launch_pdl(adam_kernel, dim3(vn_blocks((n + 3) / 4, 256)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, c, found_inf, scale_dev,
                  (const float*)nullptr);
    VN_CHECK_LAUNCH("adam_kernel");
    return VN_OK;
}

// ---- optimiser step count on the device ------------------------------------------------------------
// torch's GradScaler.step() skips optimizer.step() when an inf was found, so Adam's state['step'] -- and with it
// the bias corrections -- only advances on steps that are applied.  A host-side counter cannot know that without a
// sync; opt_state = {step_size = lr / (1 - beta1^t), sqrt(1 - beta2^t), bit pattern of the int32 number of APPLIED
// steps, unused} lives on the device: the scaler update advances the count when found_inf == 0 and derives the two
// constants of the NEXT step in double precision (the arithmetic of vn_make_adam_cfg), one thread.
__global__ void opt_state_init_kernel(float* opt_state, int applied, double lr, double beta1, double beta2) {
    vn_opt_state_set(opt_state, applied, lr, beta1, beta2);
    opt_state[3] = 0.0f;
}
VN_API int vn_opt_state_init(float* opt_state, int applied_steps, double lr, double beta1, double beta2, void* stream) {
    VN_REQUIRE(opt_state && applied_steps >= 0, "vn_opt_state_init: bad arguments");
    opt_state_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(opt_state, applied_steps, lr, beta1, beta2);
    VN_CHECK_LAUNCH("opt_state_init_kernel");
    return VN_OK;
}

VN_API int vn_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                            double eps, const float* opt_state, const float* found_inf, const float* scale_dev, void* stream) {
    VN_REQUIRE(n >= 0 && opt_state, "vn_adam_step_dev: bad n / null optimiser state");
    if (n == 0) return VN_OK;
    VN_REQUIRE(p && g && m && v, "vn_adam_step_dev: null pointer");
    VN_REQUIRE(vn_aligned(p, 16) && vn_aligned(g, 16) && vn_aligned(m, 16) && vn_aligned(v, 16),
               "vn_adam_step_dev: buffers must be 16-byte aligned");
    const AdamCfg c = vn_make_adam_cfg(lr, beta1, beta2, eps, 1, 1.0f);          // step-dependent fields come from opt_state
    VnProfScope prof(VN_K_ADAM, n, (cudaStream_t)stream);
    vn_launch_pdl(adam_kernel, dim3(vn_blocks((n + 3) / 4, 256)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, c, found_inf, scale_dev,
                  opt_state);
    VN_CHECK_LAUNCH("adam_kernel");
    return VN_OK;
}

__global__ void scaler_update_kernel(float* scale, int32_t* tracker, float* found_inf, float growth, float backoff, int interval,
                                     float* opt_state, double lr, double beta1, double beta2) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    if (*found_inf != 0.0f) { *scale = *scale * backoff; *tracker = 0; }
    else {
        if (opt_state) vn_opt_state_set(opt_state, __float_as_int(opt_state[2]) + 1, lr, beta1, beta2);   // the step was applied
        const int t = *tracker + 1;
        if (t == interval) { const float grown = *scale * growth; if (isfinite(grown)) *scale = grown; *tracker = 0; }   // torch _amp_update_scale_
        else *tracker = t;
    }
    *found_inf = 0.0f;
}

VN_API int vn_scaler_update(float* scale, int32_t* growth_tracker, float* found_inf, float growth_factor,
                            float backoff_factor, int growth_interval, void* stream) {
    VN_REQUIRE(scale && growth_tracker && found_inf, "vn_scaler_update: null pointer");
    vn_launch_pdl(scaler_update_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, scale, growth_tracker, found_inf, growth_factor, backoff_factor, growth_interval,
                  (float*)nullptr, 0.0, 0.0, 0.0);
    VN_CHECK_LAUNCH("scaler_update_kernel");
    return VN_OK;
}

VN_API int vn_scaler_update_dev(float* scale, int32_t* growth_tracker, float* found_inf, float growth_factor,
                                float backoff_factor, int growth_interval, float* opt_state, double lr, double beta1,
                                double beta2, void* stream) {
    VN_REQUIRE(scale && growth_tracker && found_inf && opt_state, "vn_scaler_update_dev: null pointer");
    vn_launch_pdl(scaler_update_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, scale, growth_tracker, found_inf, growth_factor, backoff_factor, growth_interval,
                  opt_state, lr, beta1, beta2);
    VN_CHECK_LAUNCH("scaler_update_kernel");
    return VN_OK;
}

// ---- f2: training/loss.py as two kernels --------------------------------------------------
__global__ void __launch_bounds__(256) loss_fwd_kernel(const float* __restrict__ rgb, const float* __restrict__ opacity,
                                                       const float* __restrict__ depth, const float* __restrict__ gt_rgb,
                                                       const float* __restrict__ uss, const float* __restrict__ tof,
                                                       const float* __restrict__ rgbd, int64_t N, float bg, float uss_tol,
                                                       float* __restrict__ sums, float* __restrict__ counts) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    __shared__ float sm[8][8];
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // sums[4], counts[4]
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const RayLoss r = ray_loss(rgb, opacity, depth, gt_rgb, uss, tof, rgbd, n, bg, uss_tol);
        v[0] += r.dc[0] * r.dc[0] + r.dc[1] * r.dc[1] + r.dc[2] * r.dc[2]; v[4] += 3.0f;
        v[1] += r.e_uss * r.e_uss; v[5] += r.v_uss ? 1.0f : 0.0f;
        v[2] += r.e_tof * r.e_tof; v[6] += r.v_tof ? 1.0f : 0.0f;
        v[3] += r.e_rgbd * r.e_rgbd; v[7] += r.v_rgbd ? 1.0f : 0.0f;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if (lane == 0) sm[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
        if (t != 0.0f) atomicAdd((threadIdx.x < 4 ? sums : counts) + (threadIdx.x & 3), t);
    }
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ opacity,
                                                       const float* __restrict__ depth, const float* __restrict__ gt_rgb,
                                                       const float* __restrict__ uss, const float* __restrict__ tof,
                                                       const float* __restrict__ rgbd, int64_t N, float bg, float uss_tol,
                                                       const float* __restrict__ sums, const float* __restrict__ counts,
                                                       float w_color, float w_uss, float w_tof, float w_rgbd,
                                                       const float* __restrict__ scale_dev, float* __restrict__ dL_drgb,
                                                       float* __restrict__ dL_ddepth, float* __restrict__ dL_dopacity,
                                                       float* __restrict__ loss_out) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    const float scale = scale_dev ? *scale_dev : 1.0f;
    const float c0 = counts[0], c1 = counts[1], c2 = counts[2], c3 = counts[3];
    // d(mean)/dx = 2 x / count; empty masks contribute nothing (loss.py:140-141, 186-190)
    const float k_color = c0 > 0.f ? scale * w_color * 2.0f / c0 : 0.f;
    const float k_uss = c1 > 0.f ? scale * w_uss * 2.0f / c1 : 0.f;
    const float k_tof = c2 > 0.f ? scale * w_tof * 2.0f / c2 : 0.f;
    const float k_rgbd = c3 > 0.f ? scale * w_rgbd * 2.0f / c3 : 0.f;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0 && loss_out)
        loss_out[0] = (c0 > 0.f ? w_color * sums[0] / c0 : 0.f) + (c1 > 0.f ? w_uss * sums[1] / c1 : 0.f) +
                      (c2 > 0.f ? w_tof * sums[2] / c2 : 0.f) + (c3 > 0.f ? w_rgbd * sums[3] / c3 : 0.f);
    if (n >= N) return;
    const RayLoss r = ray_loss(rgb, opacity, depth, gt_rgb, uss, tof, rgbd, n, bg, uss_tol);
    float go = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float g = k_color * r.dc[c]; dL_drgb[3 * n + c] = g; go -= bg * g; }
    dL_dopacity[n] = go;
    dL_ddepth[n] = k_uss * r.e_uss + k_tof * r.e_tof + k_rgbd * r.e_rgbd;
}

VN_API int vn_loss_fwd(const float* rgb, const float* opacity, const float* depth, const float* gt_rgb, const float* uss,
                       const float* tof, const float* rgbd, int64_t N, float bg, float uss_tol, float* sums, float* counts,
                       void* stream) {
    VN_REQUIRE(N >= 0 && sums && counts, "vn_loss_fwd: bad arguments");
    if (N == 0) return VN_OK;
    VN_REQUIRE(rgb && opacity && depth && gt_rgb, "vn_loss_fwd: null pointer");
    int64_t blocks = (N + 255) / 256;
    if (blocks > 2 * vn_sm_count()) blocks = 2 * vn_sm_count();
    vn_launch_pdl(loss_fwd_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, rgb, opacity, depth, gt_rgb, uss, tof, rgbd, N, bg,
                                                                      uss_tol, sums, counts);
    VN_CHECK_LAUNCH("loss_fwd_kernel");
    return VN_OK;
}

VN_API int vn_loss_bwd(const float* rgb, const float* opacity, const float* depth, const float* gt_rgb, const float* uss,
                       const float* tof, const float* rgbd, int64_t N, float bg, float uss_tol, const float* sums,
                       const float* counts, float w_color, float w_uss, float w_tof, float w_rgbd, const float* scale_dev,
                       float* dL_drgb, float* dL_ddepth, float* dL_dopacity, float* loss_out, void* stream) {
    VN_REQUIRE(N >= 0 && sums && counts, "vn_loss_bwd: bad arguments");
    if (N == 0) return VN_OK;
    VN_REQUIRE(rgb && opacity && depth && gt_rgb && dL_drgb && dL_ddepth && dL_dopacity, "vn_loss_bwd: null pointer");
    vn_launch_pdl(loss_bwd_kernel, dim3(vn_blocks(N, 256)), dim3(256), 0, (cudaStream_t)stream, rgb, opacity, depth, gt_rgb, uss, tof, rgbd, N, bg,
                                                                       uss_tol, sums, counts, w_color, w_uss, w_tof, w_rgbd,
                                                                       scale_dev, dL_drgb, dL_ddepth, dL_dopacity, loss_out);
    VN_CHECK_LAUNCH("loss_bwd_kernel");
    return VN_OK;
}
