// optim.cu -- caller side of the path: GradScaler.unscale_ +
// inf check + torch.optim.Adam(eps=1e-15) step over the dense hash table and the MLP
// weights as ONE streaming pass each (training/trainer.py:49-57, 138-141).
#include "common.cuh"

__global__ void __launch_bounds__(256) grad_check_kernel(const float4* __restrict__ g4, const float* __restrict__ g, int64_t n,
                                                         float* __restrict__ found_inf) {
    const int64_t n4 = n / 4;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
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
    grad_check_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)g, g, n, found_inf);
    VN_CHECK_LAUNCH("grad_check_kernel");
    return VN_OK;
}

struct AdamCfg { float inv_scale, beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt; };

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamCfg& c) {
    g = g * c.inv_scale;                                  // GradScaler.unscale_
    m = m + (g - m) * c.omb1;                             // exp_avg.lerp_(grad, 1 - beta1)
    v = v * c.beta2 + (c.omb2 * g) * g;                   // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;    // (sqrt / bias_correction2_sqrt).add_(eps)
    p = p - c.step_size * (m / denom);                    // param.addcdiv_(exp_avg, denom, -step_size)
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, AdamCfg c,
                                                   const float* __restrict__ found_inf,
                                                   const float* __restrict__ scale_dev) {
    if (found_inf && *found_inf != 0.0f) return;          // GradScaler.step skips the optimizer step
    if (scale_dev) c.inv_scale = 1.0f / *scale_dev;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 P = *(float4*)(p + i), M = *(float4*)(m + i), V = *(float4*)(v + i);
        const float4 Gd = __ldg((const float4*)(g + i));
        adam1(P.x, Gd.x, M.x, V.x, c); adam1(P.y, Gd.y, M.y, V.y, c);
        adam1(P.z, Gd.z, M.z, V.z, c); adam1(P.w, Gd.w, M.w, V.w, c);
        *(float4*)(p + i) = P; *(float4*)(m + i) = M; *(float4*)(v + i) = V;
    } else {
        for (int64_t k = i; k < n; ++k) adam1(p[k], g[k], m[k], v[k], c);
    }
}

VN_API int vn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float inv_scale, float lr, float beta1,
                        float beta2, float eps, int step, const float* found_inf, const float* scale_dev,
                        void* stream) {
    VN_REQUIRE(n >= 0 && step >= 1, "vn_adam_step: bad n/step");
    if (n == 0) return VN_OK;
    VN_REQUIRE(p && g && m && v, "vn_adam_step: null pointer");
    VN_REQUIRE(vn_aligned(p, 16) && vn_aligned(g, 16) && vn_aligned(m, 16) && vn_aligned(v, 16),
               "vn_adam_step: buffers must be 16-byte aligned");
    AdamCfg c;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    c.inv_scale = inv_scale; c.beta1 = beta1; c.beta2 = beta2;
    c.omb1 = 1.0f - beta1; c.omb2 = 1.0f - beta2; c.eps = eps;
    c.step_size = (float)((double)lr / bc1);
    c.bc2_sqrt = (float)sqrt(bc2);
    adam_kernel<<<vn_blocks((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, c, found_inf, scale_dev);
    VN_CHECK_LAUNCH("adam_kernel");
    return VN_OK;
}

__global__ void scaler_update_kernel(float* scale, int32_t* tracker, float* found_inf, float growth, float backoff, int interval) {
    if (*found_inf != 0.0f) { *scale = *scale * backoff; *tracker = 0; }
    else {
        const int t = *tracker + 1;
        if (t == interval) { *scale = *scale * growth; *tracker = 0; }
        else *tracker = t;
    }
    *found_inf = 0.0f;
}

VN_API int vn_scaler_update(float* scale, int32_t* growth_tracker, float* found_inf, float growth_factor,
                            float backoff_factor, int growth_interval, void* stream) {
    VN_REQUIRE(scale && growth_tracker && found_inf, "vn_scaler_update: null pointer");
    scaler_update_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(scale, growth_tracker, found_inf, growth_factor, backoff_factor, growth_interval);
    VN_CHECK_LAUNCH("scaler_update_kernel");
    return VN_OK;
}
