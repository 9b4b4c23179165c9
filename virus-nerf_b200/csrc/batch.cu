// batch.cu -- batch assembly on the device: the gather half of
// DatasetBase.__call__ (datasets/dataset_base.py:23-76), _calcRayPoses (:194-243) and get_rays
// (datasets/ray_utils.py:51-80) as ONE kernel instead of ~40 torch launches (a boolean-mask
// loop over the cameras, two advanced-indexing gathers per tensor, clones).
#include "common.cuh"

struct BatchArgs {
    const int32_t* img_idxs; const int32_t* pix_idxs; int64_t B;
    const float* poses;          // [N_img, 3, 4] camera-to-world
    const int32_t* cam_slot;     // [N_img] row of `directions` that belongs to the image's camera
    const float* directions;     // [n_cams, HW, 3] camera-frame pixel directions
    int64_t HW; int n_cams; int64_t n_imgs;
    const float* rgbs; int rgb_stride;   // [N_img, HW, rgb_stride], first three channels are used
    const float* depth_maps[4];  // up to four [N_img, HW] sensor depth maps (NULL = absent)
    const int32_t* sensor_ids;   // [N_img] or NULL
    const float* times;          // [N_img] or NULL
    float* rays_o; float* rays_d; float* rgb; float* depth_out[4]; int32_t* ids_out; float* time_out;
    int32_t* err;                // set to 1 when an index is out of range (the ray is filled with NaN)
};

__global__ void __launch_bounds__(256) batch_assemble_kernel(const BatchArgs a) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.B) return;
    const int64_t i = a.img_idxs[n], p = a.pix_idxs[n];
    const bool ok = i >= 0 && i < a.n_imgs && p >= 0 && p < a.HW;
    int cam = ok ? a.cam_slot[i] : -1;
    if (cam < 0 || cam >= a.n_cams) {                       // dataset_base.py:218-219: rays stay NaN
        if (a.err) *a.err = 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) { a.rays_o[3 * n + k] = NAN; a.rays_d[3 * n + k] = NAN; if (a.rgb) a.rgb[3 * n + k] = NAN; }
        for (int s = 0; s < 4; ++s) if (a.depth_out[s]) a.depth_out[s][n] = NAN;
        if (a.ids_out) a.ids_out[n] = -1;
        if (a.time_out) a.time_out[n] = NAN;
        return;
    }
    const float* dir = a.directions + ((int64_t)cam * a.HW + p) * 3;
    const float d0 = __ldg(dir), d1 = __ldg(dir + 1), d2 = __ldg(dir + 2);
    const float* c2w = a.poses + i * 12;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        // rays_d = directions @ c2w[:, :3]^T (ray_utils.py:67-71): row r of the rotation, k = 0, 1, 2
        // (separately rounded products and sums, the oracle's no-contraction convention)
        float acc = vn_mul(d0, __ldg(c2w + 4 * r));
        acc = vn_add(acc, vn_mul(d1, __ldg(c2w + 4 * r + 1)));
        acc = vn_add(acc, vn_mul(d2, __ldg(c2w + 4 * r + 2)));
        a.rays_d[3 * n + r] = acc;
        a.rays_o[3 * n + r] = __ldg(c2w + 4 * r + 3);        // :74
    }
    if (a.rgb) {
        const float* px = a.rgbs + ((int64_t)i * a.HW + p) * a.rgb_stride;
#pragma unroll
        for (int k = 0; k < 3; ++k) a.rgb[3 * n + k] = __ldg(px + k);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (a.depth_out[s]) a.depth_out[s][n] = __ldg(a.depth_maps[s] + i * a.HW + p);
    if (a.ids_out) a.ids_out[n] = a.sensor_ids[i];
    if (a.time_out) a.time_out[n] = a.times[i];
}

VN_API int vn_batch_assemble(const int32_t* img_idxs, const int32_t* pix_idxs, int64_t B, const float* poses,
                             const int32_t* cam_slot, int64_t n_imgs, const float* directions, int n_cams, int64_t HW,
                             const float* rgbs, int rgb_stride, const float* depth0, const float* depth1,
                             const float* depth2, const float* depth3, const int32_t* sensor_ids, const float* times,
                             float* rays_o, float* rays_d, float* rgb, float* out_depth0, float* out_depth1,
                             float* out_depth2, float* out_depth3, int32_t* ids_out, float* time_out, int32_t* err,
                             void* stream) {
    VN_REQUIRE(B >= 0 && n_imgs >= 0 && n_cams >= 1 && HW >= 1, "vn_batch_assemble: bad sizes");
    if (B == 0) return VN_OK;
    VN_REQUIRE(img_idxs && pix_idxs && poses && cam_slot && directions && rays_o && rays_d, "vn_batch_assemble: null pointer");
    VN_REQUIRE(!rgb || (rgbs && rgb_stride >= 3), "vn_batch_assemble: rgb output without an rgb source");
    VN_REQUIRE((!out_depth0 || depth0) && (!out_depth1 || depth1) && (!out_depth2 || depth2) && (!out_depth3 || depth3),
               "vn_batch_assemble: depth output without a depth map");
    VN_REQUIRE((!ids_out || sensor_ids) && (!time_out || times), "vn_batch_assemble: id/time output without a source");
    BatchArgs a;
    a.img_idxs = img_idxs; a.pix_idxs = pix_idxs; a.B = B; a.poses = poses; a.cam_slot = cam_slot; a.directions = directions;
    a.HW = HW; a.n_cams = n_cams; a.n_imgs = n_imgs; a.rgbs = rgbs; a.rgb_stride = rgb_stride;
    a.depth_maps[0] = depth0; a.depth_maps[1] = depth1; a.depth_maps[2] = depth2; a.depth_maps[3] = depth3;
    a.sensor_ids = sensor_ids; a.times = times; a.rays_o = rays_o; a.rays_d = rays_d; a.rgb = rgb;
    a.depth_out[0] = out_depth0; a.depth_out[1] = out_depth1; a.depth_out[2] = out_depth2; a.depth_out[3] = out_depth3;
    a.ids_out = ids_out; a.time_out = time_out; a.err = err;
    batch_assemble_kernel<<<vn_blocks(B, 256), 256, 0, (cudaStream_t)stream>>>(a);
    VN_CHECK_LAUNCH("batch_assemble_kernel");
    return VN_OK;
}

// ---- gather from a pre-assembled RAY POOL (rays_o / rays_d / rgb / up to three depth sensors already per ray) -----------
// The reference's real-time mode keeps the rays of the images seen so far resident on the device; a batch is
// pool[sel[draw[n]]] (sel = the subset of rays that carry a measurement of the sensor being sampled, or NULL for the
// whole pool).  One launch writes all six outputs instead of one advanced-indexing launch per tensor.
__global__ void __launch_bounds__(256) pool_gather_kernel(const int64_t* __restrict__ draw, const int64_t* __restrict__ sel,
                                                          int64_t n_sel, int64_t pool, int64_t B,
                                                          const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                          const float* __restrict__ rgb, const float* __restrict__ d0,
                                                          const float* __restrict__ d1, const float* __restrict__ d2,
                                                          float* __restrict__ o_rays_o, float* __restrict__ o_rays_d,
                                                          float* __restrict__ o_rgb, float* __restrict__ o_d0,
                                                          float* __restrict__ o_d1, float* __restrict__ o_d2, int32_t* err) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B) return;
    int64_t i = draw[n];
    bool ok = i >= 0 && i < (sel ? n_sel : pool);
    if (ok && sel) { i = sel[i]; ok = i >= 0 && i < pool; }
    if (!ok) {
        if (err) *err = 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o_rays_o[3 * n + k] = NAN; o_rays_d[3 * n + k] = NAN; if (o_rgb) o_rgb[3 * n + k] = NAN; }
        if (o_d0) o_d0[n] = NAN; if (o_d1) o_d1[n] = NAN; if (o_d2) o_d2[n] = NAN;
        return;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o_rays_o[3 * n + k] = __ldg(rays_o + 3 * i + k);
        o_rays_d[3 * n + k] = __ldg(rays_d + 3 * i + k);
        if (o_rgb) o_rgb[3 * n + k] = __ldg(rgb + 3 * i + k);
    }
    if (o_d0) o_d0[n] = __ldg(d0 + i);
    if (o_d1) o_d1[n] = __ldg(d1 + i);
    if (o_d2) o_d2[n] = __ldg(d2 + i);
}

VN_API int vn_pool_gather(const int64_t* draw, const int64_t* sel, int64_t n_sel, int64_t pool, int64_t B,
                          const float* rays_o, const float* rays_d, const float* rgb, const float* depth0,
                          const float* depth1, const float* depth2, float* out_rays_o, float* out_rays_d, float* out_rgb,
                          float* out_depth0, float* out_depth1, float* out_depth2, int32_t* err, void* stream) {
    VN_REQUIRE(B >= 0 && pool >= 1 && (!sel || n_sel >= 1), "vn_pool_gather: bad sizes");
    if (B == 0) return VN_OK;
    VN_REQUIRE(draw && rays_o && rays_d && out_rays_o && out_rays_d, "vn_pool_gather: null pointer");
    VN_REQUIRE((!out_rgb || rgb) && (!out_depth0 || depth0) && (!out_depth1 || depth1) && (!out_depth2 || depth2),
               "vn_pool_gather: output without a source");
    pool_gather_kernel<<<vn_blocks(B, 256), 256, 0, (cudaStream_t)stream>>>(draw, sel, n_sel, pool, B, rays_o, rays_d, rgb, depth0,
                                                                           depth1, depth2, out_rays_o, out_rays_d, out_rgb,
                                                                           out_depth0, out_depth1, out_depth2, err);
    VN_CHECK_LAUNCH("pool_gather_kernel");
    return VN_OK;
}
