// composite.cu -- per-ray alpha compositing: training forward/backward and test-time.
// Replaces modules/volume_train.py:6-48 (+ Taichi autodiff, :130-175) and
// modules/volume_render_test.py:4-54.
//
// Transmittance lives in a register (the reference round-trips a global T[] scratch and
// read-modify-writes the per-ray outputs every sample); skipped samples get ws = 0.
#include "common.cuh"

// alpha = 1 - exp(-sigma * delta), volume_train.py:37.  The exp is the correctly rounded
// binary32 value (double exp rounded once): 1 - exp(-x) cancels for small x, so a 1-ulp expf
// difference would be a 6e-8 absolute difference in every weight (see oracle.cpp exp_cr).
__device__ __forceinline__ float alpha_of(float sigma, float delta) {
    return vn_sub(1.0f, (float)exp((double)vn_mul(-sigma, delta)));
}

// ---- a8 ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) composite_fwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                            const float* __restrict__ deltas, const float* __restrict__ ts,
                                                            const int32_t* __restrict__ rays_a, int64_t N, int64_t S,
                                                            float T_thr, int32_t* __restrict__ total_samples,
                                                            float* __restrict__ opacity, float* __restrict__ depth,
                                                            float* __restrict__ rgb, float* __restrict__ ws) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int ray = rays_a[3 * n];
    const int64_t start = rays_a[3 * n + 1];
    int ns = rays_a[3 * n + 2];
    if (start + ns > S) ns = (int)max((int64_t)0, S - start);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, dep = 0.f, op = 0.f, T = 1.0f;
    int cnt = 0;
    for (int k = 0; k < ns; ++k) {
        const int64_t s = start + k;
        if (T > T_thr) {                                            // :36
            const float a = alpha_of(__ldg(sigmas + s), __ldg(deltas + s));
            const float w = vn_mul(a, T);                           // :38
            r0 = vn_add(r0, vn_mul(w, __ldg(rgbs + 3 * s)));
            r1 = vn_add(r1, vn_mul(w, __ldg(rgbs + 3 * s + 1)));
            r2 = vn_add(r2, vn_mul(w, __ldg(rgbs + 3 * s + 2)));
            dep = vn_add(dep, vn_mul(w, __ldg(ts + s)));
            op = vn_add(op, w);
            ws[s] = w;
            T = vn_mul(T, vn_sub(1.0f, a));                         // :47
            ++cnt;
        } else {
            ws[s] = 0.0f;
        }
    }
    rgb[3 * ray] = r0; rgb[3 * ray + 1] = r1; rgb[3 * ray + 2] = r2;
    depth[ray] = dep; opacity[ray] = op; total_samples[ray] = cnt;
}

VN_API int vn_composite_train_fwd(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                  const int32_t* rays_a, int64_t N, int64_t S, float T_threshold,
                                  int32_t* total_samples, float* opacity, float* depth, float* rgb, float* ws,
                                  void* stream) {
    VN_REQUIRE(N >= 0 && S >= 0, "vn_composite_train_fwd: negative size");
    if (N == 0) return VN_OK;
    VN_REQUIRE(rays_a && total_samples && opacity && depth && rgb, "vn_composite_train_fwd: null pointer");
    VN_REQUIRE(S == 0 || (sigmas && rgbs && deltas && ts && ws), "vn_composite_train_fwd: null sample pointer");
    composite_fwd_kernel<<<vn_blocks(N, 128), 128, 0, (cudaStream_t)stream>>>(sigmas, rgbs, deltas, ts, rays_a, N, S,
                                                                            T_threshold, total_samples, opacity,
                                                                            depth, rgb, ws);
    VN_CHECK_LAUNCH("composite_fwd_kernel");
    return VN_OK;
}

// ---- a9 ---------------------------------------------------------------------------------
// With G_s = dL/drgb.c_s + dL/ddepth t_s + dL/dopacity + dL/dws_s and R = sum_j G_j w_j:
//   dL/dsigma_s = delta_s * (G_s T_{s+1} - sum_{j>s} G_j w_j),   dL/dc_s = w_s dL/drgb.
// Sweep 1 accumulates R, sweep 2 walks front to back with the running prefix (both in double:
// the suffix R - prefix cancels heavily for the last samples of a ray).
__global__ void __launch_bounds__(128) composite_bwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                            const float* __restrict__ deltas, const float* __restrict__ ts,
                                                            const int32_t* __restrict__ rays_a, int64_t N, int64_t S,
                                                            float T_thr, const float* __restrict__ dL_dopacity,
                                                            const float* __restrict__ dL_ddepth, const float* __restrict__ dL_drgb,
                                                            const float* __restrict__ dL_dws, float* __restrict__ dsigmas,
                                                            float* __restrict__ drgbs) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int ray = rays_a[3 * n];
    const int64_t start = rays_a[3 * n + 1];
    int ns = rays_a[3 * n + 2];
    if (start + ns > S) ns = (int)max((int64_t)0, S - start);
    const float g0 = __ldg(dL_drgb + 3 * ray), g1 = __ldg(dL_drgb + 3 * ray + 1), g2 = __ldg(dL_drgb + 3 * ray + 2);
    const float gd = __ldg(dL_ddepth + ray), go = __ldg(dL_dopacity + ray);
    double R = 0.0;
    float T = 1.0f;
    int last = 0;
    for (int k = 0; k < ns; ++k) {
        const int64_t s = start + k;
        if (!(T > T_thr)) break;
        const float a = alpha_of(__ldg(sigmas + s), __ldg(deltas + s));
        const float w = vn_mul(a, T);
        float G = g0 * __ldg(rgbs + 3 * s) + g1 * __ldg(rgbs + 3 * s + 1) + g2 * __ldg(rgbs + 3 * s + 2) +
                  gd * __ldg(ts + s) + go;
        if (dL_dws) G += __ldg(dL_dws + s);
        R += (double)G * (double)w;
        T = vn_mul(T, vn_sub(1.0f, a));
        last = k + 1;
    }
    T = 1.0f;
    double prefix = 0.0;
    for (int k = 0; k < last; ++k) {
        const int64_t s = start + k;
        const float delta = __ldg(deltas + s);
        const float a = alpha_of(__ldg(sigmas + s), delta);
        const float w = vn_mul(a, T);
        float G = g0 * __ldg(rgbs + 3 * s) + g1 * __ldg(rgbs + 3 * s + 1) + g2 * __ldg(rgbs + 3 * s + 2) +
                  gd * __ldg(ts + s) + go;
        if (dL_dws) G += __ldg(dL_dws + s);
        prefix += (double)G * (double)w;
        const float Tn = vn_mul(T, vn_sub(1.0f, a));
        dsigmas[s] = (float)((double)delta * ((double)G * (double)Tn - (R - prefix)));
        drgbs[3 * s] = w * g0; drgbs[3 * s + 1] = w * g1; drgbs[3 * s + 2] = w * g2;
        T = Tn;
    }
    for (int k = last; k < ns; ++k) {
        const int64_t s = start + k;
        dsigmas[s] = 0.0f;
        drgbs[3 * s] = 0.0f; drgbs[3 * s + 1] = 0.0f; drgbs[3 * s + 2] = 0.0f;
    }
}

VN_API int vn_composite_train_bwd(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                  const int32_t* rays_a, int64_t N, int64_t S, float T_threshold,
                                  const float* dL_dopacity, const float* dL_ddepth, const float* dL_drgb,
                                  const float* dL_dws, float* dsigmas, float* drgbs, void* stream) {
    VN_REQUIRE(N >= 0 && S >= 0, "vn_composite_train_bwd: negative size");
    if (N == 0 || S == 0) return VN_OK;
    VN_REQUIRE(sigmas && rgbs && deltas && ts && rays_a && dL_dopacity && dL_ddepth && dL_drgb && dsigmas && drgbs,
               "vn_composite_train_bwd: null pointer");
    composite_bwd_kernel<<<vn_blocks(N, 128), 128, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, ts, rays_a, N, S, T_threshold, dL_dopacity, dL_ddepth, dL_drgb, dL_dws, dsigmas, drgbs);
    VN_CHECK_LAUNCH("composite_bwd_kernel");
    return VN_OK;
}

// ---- a10 --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) composite_test_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                             const float* __restrict__ deltas, const float* __restrict__ ts,
                                                             const int64_t* __restrict__ pack_info, int64_t* __restrict__ alive,
                                                             int64_t A, float T_thr, float* __restrict__ opacity,
                                                             float* __restrict__ depth, float* __restrict__ rgb) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= A) return;
    const int64_t start = pack_info[2 * n], steps = pack_info[2 * n + 1], ray = alive[n];
    if (steps == 0) { alive[n] = -1; return; }                    // volume_render_test.py:23-24
    float T = vn_sub(1.0f, opacity[ray]);                         // :26
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, dep = 0.f, op = 0.f;
    for (int64_t s = 0; s < steps; ++s) {
        const int64_t k = start + s;
        const float a = alpha_of(__ldg(sigmas + k), __ldg(deltas + k));   // :35
        const float w = vn_mul(a, T);
        c0 = vn_add(c0, vn_mul(w, __ldg(rgbs + 3 * k)));
        c1 = vn_add(c1, vn_mul(w, __ldg(rgbs + 3 * k + 1)));
        c2 = vn_add(c2, vn_mul(w, __ldg(rgbs + 3 * k + 2)));
        dep = vn_add(dep, vn_mul(w, __ldg(ts + k)));
        op = vn_add(op, w);
        T = vn_mul(T, vn_sub(1.0f, a));                           // :44
        if (T <= T_thr) { alive[n] = -1; break; }                 // :46-48
    }
    rgb[3 * ray] = vn_add(rgb[3 * ray], c0);                      // :50-54
    rgb[3 * ray + 1] = vn_add(rgb[3 * ray + 1], c1);
    rgb[3 * ray + 2] = vn_add(rgb[3 * ray + 2], c2);
    depth[ray] = vn_add(depth[ray], dep);
    opacity[ray] = vn_add(opacity[ray], op);
}

VN_API int vn_composite_test(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                             const int64_t* pack_info, int64_t* alive, int64_t A, float T_threshold, float* opacity,
                             float* depth, float* rgb, void* stream) {
    VN_REQUIRE(A >= 0, "vn_composite_test: A < 0");
    if (A == 0) return VN_OK;
    VN_REQUIRE(pack_info && alive && opacity && depth && rgb, "vn_composite_test: null pointer");
    composite_test_kernel<<<vn_blocks(A, 256), 256, 0, (cudaStream_t)stream>>>(sigmas, rgbs, deltas, ts, pack_info, alive,
                                                                             A, T_threshold, opacity, depth, rgb);
    VN_CHECK_LAUNCH("composite_test_kernel");
    return VN_OK;
}
