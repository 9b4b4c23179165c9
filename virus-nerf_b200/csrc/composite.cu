// composite.cu -- per-ray alpha compositing: training forward/backward and test-time.
// Replaces modules/volume_train.py:6-48 (+ Taichi autodiff, :130-175) and
// modules/volume_render_test.py:4-54.
//
// Transmittance lives in registers (the reference round-trips a global T[] scratch and
// read-modify-writes the per-ray outputs every sample); skipped samples get ws = 0.
#include "common.cuh"
#include "loss_common.cuh"

// alpha = 1 - exp(-sigma * delta), volume_train.py:37.  The exp is the correctly rounded
// binary32 value (double exp rounded once): 1 - exp(-x) cancels for small x, so a 1-ulp expf
// difference would be a 6e-8 absolute difference in every weight (see oracle.cpp exp_cr).
__device__ __forceinline__ float alpha_of(float sigma, float delta) {
    return vn_sub(1.0f, (float)exp((double)vn_mul(-sigma, delta)));
}

// ---- a8 ---------------------------------------------------------------------------------
// One WARP per ray.  Lanes take 32 consecutive samples (coalesced loads), alpha is computed
// per lane, the transmittance in front of each sample is an exclusive product scan over the
// warp (shuffles) times the running transmittance of the previous chunks.  T is monotone
// non-increasing, so "T > threshold" (volume_train.py:36) selects a prefix of the ray: the
// chunk loop stops at the first chunk whose first sample fails the test.
#define VN_FULL 0xffffffffu

__device__ __forceinline__ float warp_excl_prod(float v, int lane, float* total) {
    float inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float o = __shfl_up_sync(VN_FULL, inc, off);
        if (lane >= off) inc *= o;
    }
    *total = __shfl_sync(VN_FULL, inc, 31);
    const float ex = __shfl_up_sync(VN_FULL, inc, 1);
    return lane == 0 ? 1.0f : ex;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(VN_FULL, v, off);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(VN_FULL, v, off);
    return v;
}

// LOSS (vn_composite_loss_fwd): the kernel also accumulates the per-term sums / valid counts of training/loss.py from the
// outputs it has in registers (vn_loss_fwd's job; one block-level reduction, 8 atomics per block).
template <bool LOSS>
__global__ void __launch_bounds__(256) composite_fwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                            const float* __restrict__ deltas, const float* __restrict__ ts,
                                                            const int32_t* __restrict__ rays_a, int64_t N, int64_t S,
                                                            float T_thr, int32_t* __restrict__ total_samples,
                                                            float* __restrict__ opacity, float* __restrict__ depth,
                                                            float* __restrict__ rgb, float* __restrict__ ws, const LossArgs la) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (!LOSS && n >= N) return;
    const bool active = n < N;                // LOSS: every warp of the block reaches the block-level reduction
    const int ray = active ? rays_a[3 * n] : 0;
    const int64_t start = active ? rays_a[3 * n + 1] : 0;
    int ns = active ? rays_a[3 * n + 2] : 0;
    if (start + ns > S) ns = (int)max((int64_t)0, S - start);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, dep = 0.f, op = 0.f, T = 1.0f;
    int cnt = 0, k0 = 0;
    for (; k0 < ns; k0 += 32) {
        if (!(T > T_thr)) break;                                    // warp-uniform
        const int k = k0 + lane;
        const bool in = k < ns;
        const int64_t s = start + k;
        float a = 0.0f, c0 = 0.f, c1 = 0.f, c2 = 0.f, tm = 0.f;
        if (in) {
            a = alpha_of(__ldcg(sigmas + s), __ldg(deltas + s));     // :37
            c0 = __ldcg(rgbs + 3 * s); c1 = __ldcg(rgbs + 3 * s + 1); c2 = __ldcg(rgbs + 3 * s + 2);
            tm = __ldg(ts + s);
        }
        float chunk_prod;
        const float Tj = T * warp_excl_prod(1.0f - a, lane, &chunk_prod);   // transmittance in front of sample k
        const bool use = in && (Tj > T_thr);                        // :36
        const float w = use ? a * Tj : 0.0f;                        // :38
        if (in) ws[s] = w;
        r0 += w * c0; r1 += w * c1; r2 += w * c2; dep += w * tm; op += w;
        const unsigned m = __ballot_sync(VN_FULL, use);
        cnt += __popc(m);
        // transmittance after the last USED sample of this chunk (prefix property)
        const int last = 31 - __clz(m | 1u);
        const float Tafter = __shfl_sync(VN_FULL, Tj * (1.0f - a), last);
        T = (m == 0u) ? T : Tafter;
        if (m != __ballot_sync(VN_FULL, in)) { k0 += 32; break; }   // threshold hit inside this chunk
    }
    for (int k = k0 + lane; k < ns; k += 32) ws[start + k] = 0.0f;  // skipped tail
    r0 = warp_sum(r0); r1 = warp_sum(r1); r2 = warp_sum(r2); dep = warp_sum(dep); op = warp_sum(op);
    if (lane == 0 && active) {
        rgb[3 * ray] = r0; rgb[3 * ray + 1] = r1; rgb[3 * ray + 2] = r2;
        depth[ray] = dep; opacity[ray] = op; total_samples[ray] = cnt;
    }
    if (LOSS) {
        __shared__ float sm[8][8];
        const int warp = threadIdx.x >> 5;
        if (lane == 0) {
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // sums[4], counts[4]
            if (active) {
                const RayLoss r = ray_loss_vals(r0, r1, r2, op, dep, la.gt_rgb, la.uss, la.tof, la.rgbd, ray, la.bg, la.uss_tol);
                v[0] = r.dc[0] * r.dc[0] + r.dc[1] * r.dc[1] + r.dc[2] * r.dc[2]; v[4] = 3.0f;
                v[1] = r.e_uss * r.e_uss; v[5] = r.v_uss ? 1.0f : 0.0f;
                v[2] = r.e_tof * r.e_tof; v[6] = r.v_tof ? 1.0f : 0.0f;
                v[3] = r.e_rgbd * r.e_rgbd; v[7] = r.v_rgbd ? 1.0f : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) sm[warp][k] = v[k];
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            float t = 0.0f;
            for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
            if (t != 0.0f) atomicAdd((threadIdx.x < 4 ? la.sums : la.counts) + (threadIdx.x & 3), t);
        }
    }
}

VN_API int vn_composite_train_fwd(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                  const int32_t* rays_a, int64_t N, int64_t S, float T_threshold,
                                  int32_t* total_samples, float* opacity, float* depth, float* rgb, float* ws,
                                  void* stream) {
    VN_REQUIRE(N >= 0 && S >= 0, "vn_composite_train_fwd: negative size");
    if (N == 0) return VN_OK;
    VN_REQUIRE(rays_a && total_samples && opacity && depth && rgb, "vn_composite_train_fwd: null pointer");
    VN_REQUIRE(S == 0 || (sigmas && rgbs && deltas && ts && ws), "vn_composite_train_fwd: null sample pointer");
    VnProfScope prof(VN_K_COMP_FWD, S, (cudaStream_t)stream);
    vn_launch_pdl(composite_fwd_kernel<false>, dim3(vn_blocks(N * 32, 256)), dim3(256), 0, (cudaStream_t)stream, sigmas, rgbs, deltas, ts, rays_a, N, S,
                                                                                 T_threshold, total_samples, opacity,
                                                                                 depth, rgb, ws, LossArgs{});
    VN_CHECK_LAUNCH("composite_fwd_kernel");
    return VN_OK;
}

// a8 + f2: compositing forward that also accumulates the loss terms (vn_composite_train_fwd + vn_loss_fwd as one kernel)
VN_API int vn_composite_loss_fwd(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                 const int32_t* rays_a, int64_t N, int64_t S, float T_threshold, int32_t* total_samples,
                                 float* opacity, float* depth, float* rgb, float* ws, const float* gt_rgb, const float* uss,
                                 const float* tof, const float* rgbd, float bg, float uss_tol, float* sums, float* counts,
                                 void* stream) {
    VN_REQUIRE(N >= 0 && S >= 0, "vn_composite_loss_fwd: negative size");
    if (N == 0) return VN_OK;
    VN_REQUIRE(rays_a && total_samples && opacity && depth && rgb && gt_rgb && sums && counts, "vn_composite_loss_fwd: null pointer");
    VN_REQUIRE(S == 0 || (sigmas && rgbs && deltas && ts && ws), "vn_composite_loss_fwd: null sample pointer");
    LossArgs la{};
    la.gt_rgb = gt_rgb; la.uss = uss; la.tof = tof; la.rgbd = rgbd; la.bg = bg; la.uss_tol = uss_tol; la.sums = sums; la.counts = counts;
    VnProfScope prof(VN_K_COMP_FWD, S, (cudaStream_t)stream);
    vn_launch_pdl(composite_fwd_kernel<true>, dim3(vn_blocks(N * 32, 256)), dim3(256), 0, (cudaStream_t)stream, sigmas, rgbs, deltas, ts,
                  rays_a, N, S, T_threshold, total_samples, opacity, depth, rgb, ws, la);
    VN_CHECK_LAUNCH("composite_fwd_kernel<loss>");
    return VN_OK;
}

// ---- a9 ---------------------------------------------------------------------------------
// With G_s = dL/drgb.c_s + dL/ddepth t_s + dL/dopacity + dL/dws_s and R = sum_j G_j w_j:
//   dL/dsigma_s = delta_s * (G_s T_{s+1} - sum_{j>s} G_j w_j),   dL/dc_s = w_s dL/drgb.
// One warp per ray, two sweeps over the ray: sweep 1 accumulates R, sweep 2 forms the suffix
// as R - (inclusive prefix); R and the prefix are kept in double because the suffix cancels
// heavily for the last samples of a ray.
// LOSS (vn_composite_loss_bwd): the per-ray gradient seeds are not read from dL_d* arrays but formed here from the loss
// terms and the (global) valid counts -- vn_loss_bwd's arithmetic, evaluated by the warp that needs the seeds.
template <bool LOSS>
__global__ void __launch_bounds__(256) composite_bwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                            const float* __restrict__ deltas, const float* __restrict__ ts,
                                                            const int32_t* __restrict__ rays_a, int64_t N, int64_t S,
                                                            float T_thr, const float* __restrict__ dL_dopacity,
                                                            const float* __restrict__ dL_ddepth, const float* __restrict__ dL_drgb,
                                                            const float* __restrict__ dL_dws, float* __restrict__ dsigmas,
                                                            float* __restrict__ drgbs, const float* __restrict__ rgb_out,
                                                            const float* __restrict__ opacity_out, const float* __restrict__ depth_out,
                                                            const LossArgs la) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (LOSS && blockIdx.x == 0 && threadIdx.x == 0 && la.loss_out) {
        const float c0 = la.counts[0], c1 = la.counts[1], c2 = la.counts[2], c3 = la.counts[3];
        la.loss_out[0] = (c0 > 0.f ? la.w_color * la.sums[0] / c0 : 0.f) + (c1 > 0.f ? la.w_uss * la.sums[1] / c1 : 0.f) +
                         (c2 > 0.f ? la.w_tof * la.sums[2] / c2 : 0.f) + (c3 > 0.f ? la.w_rgbd * la.sums[3] / c3 : 0.f);
    }
    if (n >= N) return;
    const int ray = rays_a[3 * n];
    const int64_t start = rays_a[3 * n + 1];
    int ns = rays_a[3 * n + 2];
    if (start + ns > S) ns = (int)max((int64_t)0, S - start);
    float g0, g1, g2, gd, go;
    if (LOSS) {
        // d(mean)/dx = 2 x / count; empty masks contribute nothing (loss.py:140-141, 186-190) -- as loss_bwd_kernel
        const float scale = la.scale_dev ? *la.scale_dev : 1.0f;
        const float c0 = la.counts[0], c1 = la.counts[1], c2 = la.counts[2], c3 = la.counts[3];
        const float k_color = c0 > 0.f ? scale * la.w_color * 2.0f / c0 : 0.f;
        const float k_uss = c1 > 0.f ? scale * la.w_uss * 2.0f / c1 : 0.f;
        const float k_tof = c2 > 0.f ? scale * la.w_tof * 2.0f / c2 : 0.f;
        const float k_rgbd = c3 > 0.f ? scale * la.w_rgbd * 2.0f / c3 : 0.f;
        const RayLoss r = ray_loss(rgb_out, opacity_out, depth_out, la.gt_rgb, la.uss, la.tof, la.rgbd, ray, la.bg, la.uss_tol);
        g0 = k_color * r.dc[0]; g1 = k_color * r.dc[1]; g2 = k_color * r.dc[2];
        go = 0.0f;
        go -= la.bg * g0; go -= la.bg * g1; go -= la.bg * g2;
        gd = k_uss * r.e_uss + k_tof * r.e_tof + k_rgbd * r.e_rgbd;
    } else {
        g0 = __ldcg(dL_drgb + 3 * ray); g1 = __ldcg(dL_drgb + 3 * ray + 1); g2 = __ldcg(dL_drgb + 3 * ray + 2);
        gd = __ldcg(dL_ddepth + ray); go = __ldcg(dL_dopacity + ray);
    }
    // ---- sweep 1: R = sum G_j w_j over the used prefix
    double Rl = 0.0;
    float T = 1.0f;
    for (int k0 = 0; k0 < ns; k0 += 32) {
        if (!(T > T_thr)) break;
        const int k = k0 + lane;
        const bool in = k < ns;
        const int64_t s = start + k;
        float a = 0.0f, G = 0.0f;
        if (in) {
            a = alpha_of(__ldcg(sigmas + s), __ldg(deltas + s));
            G = g0 * __ldcg(rgbs + 3 * s) + g1 * __ldcg(rgbs + 3 * s + 1) + g2 * __ldcg(rgbs + 3 * s + 2) + gd * __ldg(ts + s) + go;
            if (dL_dws) G += __ldcg(dL_dws + s);
        }
        float chunk_prod;
        const float Tj = T * warp_excl_prod(1.0f - a, lane, &chunk_prod);
        const bool use = in && (Tj > T_thr);
        if (use) Rl += (double)G * (double)(a * Tj);
        const unsigned m = __ballot_sync(VN_FULL, use);
        const int last = 31 - __clz(m | 1u);
        const float Tafter = __shfl_sync(VN_FULL, Tj * (1.0f - a), last);
        T = (m == 0u) ? T : Tafter;
        if (m != __ballot_sync(VN_FULL, in)) break;
    }
    const double R = warp_sum_d(Rl);
    // ---- sweep 2
    T = 1.0f;
    double prefix = 0.0;
    int k0 = 0;
    for (; k0 < ns; k0 += 32) {
        if (!(T > T_thr)) break;
        const int k = k0 + lane;
        const bool in = k < ns;
        const int64_t s = start + k;
        float a = 0.0f, G = 0.0f, delta = 0.0f;
        if (in) {
            delta = __ldg(deltas + s);
            a = alpha_of(__ldcg(sigmas + s), delta);
            G = g0 * __ldcg(rgbs + 3 * s) + g1 * __ldcg(rgbs + 3 * s + 1) + g2 * __ldcg(rgbs + 3 * s + 2) + gd * __ldg(ts + s) + go;
            if (dL_dws) G += __ldcg(dL_dws + s);
        }
        float chunk_prod;
        const float Tj = T * warp_excl_prod(1.0f - a, lane, &chunk_prod);
        const bool use = in && (Tj > T_thr);
        const float w = use ? a * Tj : 0.0f;
        // inclusive prefix of G*w inside the chunk (double)
        double inc = (double)G * (double)w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double o = __shfl_up_sync(VN_FULL, inc, off);
            if (lane >= off) inc += o;
        }
        const double pre = prefix + inc;
        prefix += __shfl_sync(VN_FULL, inc, 31);
        if (in) {
            const float Tn = Tj * (1.0f - a);
            dsigmas[s] = use ? (float)((double)delta * ((double)G * (double)Tn - (R - pre))) : 0.0f;
            drgbs[3 * s] = w * g0; drgbs[3 * s + 1] = w * g1; drgbs[3 * s + 2] = w * g2;
        }
        const unsigned m = __ballot_sync(VN_FULL, use);
        const int last = 31 - __clz(m | 1u);
        const float Tafter = __shfl_sync(VN_FULL, Tj * (1.0f - a), last);
        T = (m == 0u) ? T : Tafter;
        if (m != __ballot_sync(VN_FULL, in)) { k0 += 32; break; }
    }
    for (int k = k0 + lane; k < ns; k += 32) {
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
    VnProfScope prof(VN_K_COMP_BWD, S, (cudaStream_t)stream);
    vn_launch_pdl(composite_bwd_kernel<false>, dim3(vn_blocks(N * 32, 256)), dim3(256), 0, (cudaStream_t)stream,
        sigmas, rgbs, deltas, ts, rays_a, N, S, T_threshold, dL_dopacity, dL_ddepth, dL_drgb, dL_dws, dsigmas, drgbs,
        (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, LossArgs{});
    VN_CHECK_LAUNCH("composite_bwd_kernel");
    return VN_OK;
}

// a9 + f2: compositing backward that forms its own gradient seeds from the loss (vn_loss_bwd + vn_composite_train_bwd)
VN_API int vn_composite_loss_bwd(const float* sigmas, const float* rgbs, const float* deltas, const float* ts,
                                 const int32_t* rays_a, int64_t N, int64_t S, float T_threshold, const float* rgb,
                                 const float* opacity, const float* depth, const float* gt_rgb, const float* uss,
                                 const float* tof, const float* rgbd, float bg, float uss_tol, const float* sums,
                                 const float* counts, float w_color, float w_uss, float w_tof, float w_rgbd,
                                 const float* scale_dev, float* dsigmas, float* drgbs, float* loss_out, void* stream) {
    VN_REQUIRE(N >= 0 && S >= 0, "vn_composite_loss_bwd: negative size");
    if (N == 0 || S == 0) return VN_OK;
    VN_REQUIRE(sigmas && rgbs && deltas && ts && rays_a && rgb && opacity && depth && gt_rgb && sums && counts && dsigmas && drgbs,
               "vn_composite_loss_bwd: null pointer");
    LossArgs la{};
    la.gt_rgb = gt_rgb; la.uss = uss; la.tof = tof; la.rgbd = rgbd; la.bg = bg; la.uss_tol = uss_tol;
    la.sums = const_cast<float*>(sums); la.counts = const_cast<float*>(counts);
    la.w_color = w_color; la.w_uss = w_uss; la.w_tof = w_tof; la.w_rgbd = w_rgbd; la.scale_dev = scale_dev; la.loss_out = loss_out;
    VnProfScope prof(VN_K_COMP_BWD, S, (cudaStream_t)stream);
    vn_launch_pdl(composite_bwd_kernel<true>, dim3(vn_blocks(N * 32, 256)), dim3(256), 0, (cudaStream_t)stream,
        sigmas, rgbs, deltas, ts, rays_a, N, S, T_threshold, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr,
        (const float*)nullptr, dsigmas, drgbs, rgb, opacity, depth, la);
    VN_CHECK_LAUNCH("composite_bwd_kernel<loss>");
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
