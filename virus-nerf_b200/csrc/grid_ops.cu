// grid_ops.cu -- spherical harmonics, Morton / bitfield helpers and the VIRUS-NeRF
// occupancy-grid update.  Replaces modules/spherical_harmonics.py:7-42,
// modules/utils.py:120-169, modules/grid.py:165-170,205-211 and the torch op sequences of
// modules/occupancy_grid.py:293-430 (with helpers/geometric_fcts.py:151-171).
#include "common.cuh"

// ---- a11 --------------------------------------------------------------------------------
__device__ __forceinline__ void sh16(float x, float y, float z, float* e) {
    const float xy = vn_mul(x, y), xz = vn_mul(x, z), yz = vn_mul(y, z);
    const float x2 = vn_mul(x, x), y2 = vn_mul(y, y), z2 = vn_mul(z, z);
    e[0] = 0.28209479177387814f;
    e[1] = vn_mul(-0.48860251190291987f, y);
    e[2] = vn_mul(0.48860251190291987f, z);
    e[3] = vn_mul(-0.48860251190291987f, x);
    e[4] = vn_mul(1.0925484305920792f, xy);
    e[5] = vn_mul(-1.0925484305920792f, yz);
    e[6] = vn_sub(vn_mul(0.94617469575755997f, z2), 0.31539156525251999f);
    e[7] = vn_mul(-1.0925484305920792f, xz);
    e[8] = vn_sub(vn_mul(0.54627421529603959f, x2), vn_mul(0.54627421529603959f, y2));
    e[9] = vn_mul(vn_mul(0.59004358992664352f, y), vn_add(vn_mul(-3.0f, x2), y2));
    e[10] = vn_mul(vn_mul(2.8906114426405538f, xy), z);
    e[11] = vn_mul(vn_mul(0.45704579946446572f, y), vn_sub(1.0f, vn_mul(5.0f, z2)));
    e[12] = vn_mul(vn_mul(0.3731763325901154f, z), vn_sub(vn_mul(5.0f, z2), 3.0f));
    e[13] = vn_mul(vn_mul(0.45704579946446572f, x), vn_sub(1.0f, vn_mul(5.0f, z2)));
    e[14] = vn_mul(vn_mul(1.4453057213202769f, z), vn_sub(x2, y2));
    e[15] = vn_mul(vn_mul(0.59004358992664352f, x), vn_add(-x2, vn_mul(3.0f, y2)));
}

__global__ void __launch_bounds__(256) sh_kernel(const float* __restrict__ dirs, int64_t B, float4* __restrict__ emb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float e[16];
    sh16(__ldg(dirs + 3 * i), __ldg(dirs + 3 * i + 1), __ldg(dirs + 3 * i + 2), e);
#pragma unroll
    for (int q = 0; q < 4; ++q) emb[4 * i + q] = make_float4(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
}

VN_API int vn_sh_encode(const float* dirs, int64_t B, float* emb, void* stream) {
    VN_REQUIRE(B >= 0 && (B == 0 || (dirs && emb)), "vn_sh_encode: bad arguments");
    VN_REQUIRE(vn_aligned(emb, 16), "vn_sh_encode: emb must be 16-byte aligned");
    if (B == 0) return VN_OK;
    sh_kernel<<<vn_blocks(B, 256), 256, 0, (cudaStream_t)stream>>>(dirs, B, (float4*)emb);
    VN_CHECK_LAUNCH("sh_kernel");
    return VN_OK;
}

// ---- a15 --------------------------------------------------------------------------------
__global__ void morton3d_kernel(const int32_t* __restrict__ coords, int64_t n, int32_t* __restrict__ indices) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        indices[i] = (int32_t)vn_morton3D((uint32_t)coords[3 * i], (uint32_t)coords[3 * i + 1], (uint32_t)coords[3 * i + 2]);
}
__global__ void morton3d_invert_kernel(const int32_t* __restrict__ indices, int64_t n, int32_t* __restrict__ coords) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t ind = (uint32_t)indices[i];
    coords[3 * i] = (int32_t)vn_morton3D_invert(ind);
    coords[3 * i + 1] = (int32_t)vn_morton3D_invert(ind >> 1);
    coords[3 * i + 2] = (int32_t)vn_morton3D_invert(ind >> 2);
}
__global__ void packbits_kernel(const float4* __restrict__ grid, int64_t n_bytes, float thr, uint8_t* __restrict__ bitfield) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_bytes) return;
    const float4 a = __ldg(grid + 2 * n), b = __ldg(grid + 2 * n + 1);
    uint32_t bits = 0;
    bits |= (a.x > thr) ? 1u : 0u;  bits |= (a.y > thr) ? 2u : 0u;
    bits |= (a.z > thr) ? 4u : 0u;  bits |= (a.w > thr) ? 8u : 0u;
    bits |= (b.x > thr) ? 16u : 0u; bits |= (b.y > thr) ? 32u : 0u;
    bits |= (b.z > thr) ? 64u : 0u; bits |= (b.w > thr) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

VN_API int vn_morton3d(const int32_t* coords, int64_t n, int32_t* indices, void* stream) {
    VN_REQUIRE(n >= 0 && (n == 0 || (coords && indices)), "vn_morton3d: bad arguments");
    if (n == 0) return VN_OK;
    morton3d_kernel<<<vn_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(coords, n, indices);
    VN_CHECK_LAUNCH("morton3d_kernel");
    return VN_OK;
}
VN_API int vn_morton3d_invert(const int32_t* indices, int64_t n, int32_t* coords, void* stream) {
    VN_REQUIRE(n >= 0 && (n == 0 || (coords && indices)), "vn_morton3d_invert: bad arguments");
    if (n == 0) return VN_OK;
    morton3d_invert_kernel<<<vn_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(indices, n, coords);
    VN_CHECK_LAUNCH("morton3d_invert_kernel");
    return VN_OK;
}
VN_API int vn_packbits(const float* grid, int64_t n_bytes, float threshold, uint8_t* bitfield, void* stream) {
    VN_REQUIRE(n_bytes >= 0 && (n_bytes == 0 || (grid && bitfield)), "vn_packbits: bad arguments");
    VN_REQUIRE(vn_aligned(grid, 16), "vn_packbits: grid must be 16-byte aligned");
    if (n_bytes == 0) return VN_OK;
    packbits_kernel<<<vn_blocks(n_bytes, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)grid, n_bytes, threshold, bitfield);
    VN_CHECK_LAUNCH("packbits_kernel");
    return VN_OK;
}

// ---- a14 --------------------------------------------------------------------------------
// torch.linspace(0, 1, steps) float32 (ATen fills symmetrically from both ends)
__device__ __forceinline__ float linspace01(int i, int steps) {
    const float step = vn_div(1.0f, (float)(steps - 1));
    return (i < steps / 2) ? vn_mul(step, (float)i) : vn_sub(1.0f, vn_mul(step, (float)(steps - i - 1)));
}

__device__ __forceinline__ float occ_pdf(float meas, float dist, float std_every_m) {   // occupancy_grid.py:464-465
    const float stds = vn_add(vn_mul(std_every_m, dist), 0.00001f);
    const float diff = vn_sub(meas, dist);
    return expf(vn_div(vn_mul(-0.5f, vn_mul(diff, diff)), vn_mul(stds, stds)));
}

// _rayProb for one (ray, m): occupancy_grid.py:361-385
__device__ __forceinline__ void ray_prob(float me, float dist, int I, float p_false, float std_every_m, float prob_min,
                                         float* po, float* pe, float* terms = nullptr, int64_t term_stride = 0) {
    const float eq_emp = p_false;                                         // :361-363
    const float eq_occ = vn_add(eq_emp, occ_pdf(me, dist, std_every_m));  // :364-367
    float nl_emp = vn_sub(1.0f, vn_mul(eq_emp, dist));                    // :370
    if (nl_emp < prob_min) nl_emp = prob_min;                             // :371
    float integral = 0.0f;
    for (int k = 0; k < I; ++k)                                           // :374-378
        integral = vn_add(integral, occ_pdf(vn_mul(linspace01(k, I), me), dist, std_every_m));
    integral = vn_mul(integral, vn_div(me, (float)I));                    // :379
    float nl_occ = vn_sub(nl_emp, integral);                              // :380
    if (nl_occ < prob_min) nl_occ = prob_min;                             // :381
    *pe = vn_mul(eq_emp, nl_emp);                                         // :384
    *po = vn_mul(eq_occ, nl_occ);                                         // :385
    if (terms) {                                                          // return_probs=True, :387-388
        terms[0] = eq_emp; terms[term_stride] = eq_occ; terms[2 * term_stride] = nl_emp; terms[3 * term_stride] = nl_occ;
    }
}

__global__ void __launch_bounds__(128) occ_calc_pos_prob_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ noise,
    const float* __restrict__ meas, int64_t N, int M, int I, int G, float scale, float noise_every_m, float p_false,
    float std_every_m, float prob_min, float* __restrict__ cell_dists, float* __restrict__ cell_pos,
    int32_t* __restrict__ cell_idxs, float* __restrict__ probs_occ, float* __restrict__ probs_emp,
    float* __restrict__ unit_pos = nullptr, float xyz_min = 0.0f, float xyz_max = 1.0f) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * M) return;
    const int64_t n = t / M;
    const int m = (int)(t % M);
    float o[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = __ldg(rays_o + 3 * n + k); d[k] = __ldg(rays_d + 3 * n + k); }
    const float nrm = sqrtf(vn_add(vn_add(vn_mul(d[0], d[0]), vn_mul(d[1], d[1])), vn_mul(d[2], d[2])));   // :311
    float L = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {                       // helpers/geometric_fcts.py:168-171
        d[k] = vn_div(d[k], nrm);
        float v = INFINITY;
        if (d[k] > 0.0f) v = vn_div(vn_sub(scale, o[k]), d[k]);
        if (d[k] < 0.0f) v = vn_div(vn_sub(-scale, o[k]), d[k]);
        if (v < L) L = v;
    }
    const float dist = vn_mul(linspace01(m, M), L);      // :318-319
    if (cell_dists) cell_dists[t] = dist;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float p = vn_add(o[k], vn_mul(d[k], dist));      // :322
        if (noise) {
            const float nz = vn_sub(vn_mul(2.0f, __ldg(noise + 3 * t + k)), 1.0f);   // :326
            p = vn_add(p, vn_mul(vn_mul(noise_every_m, dist), nz));                  // :327
        }
        if (cell_pos) cell_pos[3 * t + k] = p;
        if (unit_pos) unit_pos[3 * t + k] = vn_div(vn_sub(p, xyz_min), vn_sub(xyz_max, xyz_min));   // networks.py:141
        if (cell_idxs) {
            const float mi = vn_div(vn_mul((float)(G - 1), vn_add(p, scale)), vn_mul(2.0f, scale));   // :479
            int ii = (int)rintf(mi);                     // torch.round = half to even
            cell_idxs[3 * t + k] = min(max(ii, 0), G - 1);                                              // :480
        }
    }
    if (meas && probs_occ && probs_emp)
        ray_prob(__ldg(meas + n), dist, I, p_false, std_every_m, prob_min, probs_occ + t, probs_emp + t);
}

__global__ void __launch_bounds__(128) occ_ray_prob_kernel(const float* __restrict__ meas, const float* __restrict__ dists,
                                                           int64_t N, int M, int I, float p_false, float std_every_m,
                                                           float prob_min, float* __restrict__ probs_occ,
                                                           float* __restrict__ probs_emp, float* __restrict__ terms) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * M) return;
    ray_prob(__ldg(meas + t / M), __ldg(dists + t), I, p_false, std_every_m, prob_min, probs_occ + t, probs_emp + t,
             terms ? terms + t : nullptr, N * M);
}

VN_API int vn_occ_ray_prob(const float* meas, const float* dists, int64_t N, int M, int I, float p_false,
                           float std_every_m, float prob_min, float* probs_occ, float* probs_emp, void* stream) {
    VN_REQUIRE(N >= 0 && M >= 1 && I >= 2, "vn_occ_ray_prob: bad sizes");
    if (N == 0) return VN_OK;
    VN_REQUIRE(meas && dists && probs_occ && probs_emp, "vn_occ_ray_prob: null pointer");
    occ_ray_prob_kernel<<<vn_blocks(N * M, 128), 128, 0, (cudaStream_t)stream>>>(meas, dists, N, M, I, p_false, std_every_m,
                                                                              prob_min, probs_occ, probs_emp, nullptr);
    VN_CHECK_LAUNCH("occ_ray_prob_kernel");
    return VN_OK;
}

VN_API int vn_occ_ray_prob_terms(const float* meas, const float* dists, int64_t N, int M, int I, float p_false,
                                 float std_every_m, float prob_min, float* probs_occ, float* probs_emp, float* terms,
                                 void* stream) {
    VN_REQUIRE(N >= 0 && M >= 1 && I >= 2, "vn_occ_ray_prob_terms: bad sizes");
    if (N == 0) return VN_OK;
    VN_REQUIRE(meas && dists && probs_occ && probs_emp && terms, "vn_occ_ray_prob_terms: null pointer");
    occ_ray_prob_kernel<<<vn_blocks(N * M, 128), 128, 0, (cudaStream_t)stream>>>(meas, dists, N, M, I, p_false, std_every_m,
                                                                              prob_min, probs_occ, probs_emp, terms);
    VN_CHECK_LAUNCH("occ_ray_prob_kernel");
    return VN_OK;
}

VN_API int vn_occ_calc_pos_prob(const float* rays_o, const float* rays_d, const float* noise, const float* meas,
                                int64_t N, int M, int I, int grid_size, float scale, float noise_every_m, float p_false,
                                float std_every_m, float prob_min, float* cell_dists, float* cell_pos,
                                int32_t* cell_idxs, float* probs_occ, float* probs_emp, void* stream) {
    VN_REQUIRE(N >= 0 && M >= 2 && I >= 2 && grid_size >= 2, "vn_occ_calc_pos_prob: bad sizes");
    if (N == 0) return VN_OK;
    VN_REQUIRE(rays_o && rays_d, "vn_occ_calc_pos_prob: null rays");
    occ_calc_pos_prob_kernel<<<vn_blocks(N * M, 128), 128, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, noise, meas, N, M, I, grid_size, scale, noise_every_m, p_false, std_every_m, prob_min,
        cell_dists, cell_pos, cell_idxs, probs_occ, probs_emp);
    VN_CHECK_LAUNCH("occ_calc_pos_prob_kernel");
    return VN_OK;
}

// _nerfProb, occupancy_grid.py:392-408.  Deterministic two-stage double-precision mean.
#define VN_NP_BLOCKS 256
__global__ void __launch_bounds__(256) nerf_mean_stage1(const float* __restrict__ density, int64_t n, double* __restrict__ partials) {
    __shared__ double sm[256];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)VN_NP_BLOCKS * 256) acc += (double)density[i];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(256) nerf_mean_stage2(const double* __restrict__ partials, int64_t n, double thr_max,
                                                        float* __restrict__ out2) {
    __shared__ double sm[256];
    sm[threadIdx.x] = partials[threadIdx.x];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float mean = (float)(sm[0] / (double)n);          // torch.mean(...).item(), :402
        const double thr = fmin(thr_max, (double)mean);
        out2[0] = mean;
        out2[1] = (float)(-log(thr));                           // h_thr, :403
    }
}
__global__ void __launch_bounds__(256) nerf_prob_kernel(const float* __restrict__ density, int64_t n, float slope,
                                                        const float* __restrict__ mean_hthr, float* __restrict__ probs_occ,
                                                        float* __restrict__ probs_emp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float h_thr = mean_hthr[1];
    const float h = logf(__ldg(density + i));                                      // :404
    const float po = vn_div(1.0f, vn_add(1.0f, expf(vn_mul(-slope, vn_sub(h, h_thr)))));   // :405
    probs_occ[i] = po;
    probs_emp[i] = vn_sub(1.0f, po);                                               // :406
}

VN_API int vn_occ_nerf_prob(const float* density, int64_t n, double thr_max, float slope, float* scratch,
                            float* probs_occ, float* probs_emp, void* stream) {
    VN_REQUIRE(n >= 0, "vn_occ_nerf_prob: n < 0");
    if (n == 0) return VN_OK;
    VN_REQUIRE(density && scratch && probs_occ && probs_emp, "vn_occ_nerf_prob: null pointer");
    VN_REQUIRE(vn_aligned(scratch, 8), "vn_occ_nerf_prob: scratch must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    double* partials = (double*)scratch;                 // 256 doubles = 512 floats
    float* mean_hthr = scratch + 2 * VN_NP_BLOCKS;       // + 2 floats
    nerf_mean_stage1<<<VN_NP_BLOCKS, 256, 0, st>>>(density, n, partials);
    VN_CHECK_LAUNCH("nerf_mean_stage1");
    nerf_mean_stage2<<<1, 256, 0, st>>>(partials, n, thr_max, mean_hthr);
    VN_CHECK_LAUNCH("nerf_mean_stage2");
    nerf_prob_kernel<<<vn_blocks(n, 256), 256, 0, st>>>(density, n, slope, mean_hthr, probs_occ, probs_emp);
    VN_CHECK_LAUNCH("nerf_prob_kernel");
    return VN_OK;
}

// _updateGrid, occupancy_grid.py:428-430: gather all, Bayes, scatter.  Duplicate cells: the
// entry with the largest flat index wins (CPU index_put_ order) -- made deterministic with
// an atomicMax "winner" table that is restored to -1 on the way out.
__global__ void __launch_bounds__(256) bayes_gather_kernel(const float* __restrict__ grid, int G, const int32_t* __restrict__ cell_idxs,
                                                           int64_t n, const float* __restrict__ probs_occ,
                                                           const float* __restrict__ probs_emp, int32_t* __restrict__ winner,
                                                           float* __restrict__ new_probs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = ((int64_t)cell_idxs[3 * i] * G + cell_idxs[3 * i + 1]) * G + cell_idxs[3 * i + 2];
    const float p = grid[c], po = __ldg(probs_occ + i), pe = __ldg(probs_emp + i);
    const float num = vn_mul(p, po);
    new_probs[i] = vn_div(num, vn_add(num, vn_mul(vn_sub(1.0f, p), pe)));          // :429
    atomicMax(winner + c, (int32_t)i);
}
__global__ void __launch_bounds__(256) bayes_scatter_kernel(float* __restrict__ grid, int G, const int32_t* __restrict__ cell_idxs,
                                                            int64_t n, int32_t* __restrict__ winner,
                                                            const float* __restrict__ new_probs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = ((int64_t)cell_idxs[3 * i] * G + cell_idxs[3 * i + 1]) * G + cell_idxs[3 * i + 2];
    if (winner[c] == (int32_t)i) {
        grid[c] = new_probs[i];                                                     // :430
        winner[c] = -1;
    }
}

VN_API int vn_occ_bayes_update(float* grid, int grid_size, const int32_t* cell_idxs, int64_t n, const float* probs_occ,
                               const float* probs_emp, int32_t* winner, float* new_probs_tmp, void* stream) {
    VN_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "vn_occ_bayes_update: n out of range");
    if (n == 0) return VN_OK;
    VN_REQUIRE(grid && cell_idxs && probs_occ && probs_emp && winner && new_probs_tmp, "vn_occ_bayes_update: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    bayes_gather_kernel<<<vn_blocks(n, 256), 256, 0, st>>>(grid, grid_size, cell_idxs, n, probs_occ, probs_emp, winner, new_probs_tmp);
    VN_CHECK_LAUNCH("bayes_gather_kernel");
    bayes_scatter_kernel<<<vn_blocks(n, 256), 256, 0, st>>>(grid, grid_size, cell_idxs, n, winner, new_probs_tmp);
    VN_CHECK_LAUNCH("bayes_scatter_kernel");
    return VN_OK;
}

// update() tail, occupancy_grid.py:96-105: decay (in place) + cartesian -> Morton -> bits.
// One thread per output byte = one 2x2x2 block of cells (the low three Morton bits).
__global__ void __launch_bounds__(256) decay_pack_kernel(float* __restrict__ grid, int G, float decay, int apply_decay,
                                                         float thr, int64_t n_bytes, uint8_t* __restrict__ bitfield) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_bytes) return;
    const uint32_t m0 = (uint32_t)(8 * n);
    const uint32_t bx = vn_morton3D_invert(m0), by = vn_morton3D_invert(m0 >> 1), bz = vn_morton3D_invert(m0 >> 2);
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t x = bx + (i & 1), y = by + ((i >> 1) & 1), z = bz + ((i >> 2) & 1);   // utils.py:104-107
        const int64_t c = ((int64_t)x * G + y) * G + z;
        float v = grid[c];
        if (apply_decay) { v = vn_mul(v, decay); grid[c] = v; }                              // :98
        bits |= (v > thr) ? (1u << i) : 0u;                                                  // utils.py:165-167
    }
    bitfield[n] = (uint8_t)bits;
}

VN_API int vn_occ_decay_pack(float* grid, int grid_size, float decay, int apply_decay, float threshold, uint8_t* bitfield,
                             void* stream) {
    VN_REQUIRE(grid && bitfield, "vn_occ_decay_pack: null pointer");
    VN_REQUIRE(grid_size >= 2 && grid_size <= 1024 && (grid_size & (grid_size - 1)) == 0,
               "vn_occ_decay_pack: grid_size must be a power of two in [2,1024] (10-bit Morton)");
    const int64_t G = grid_size, nb = G * G * G / 8;
    decay_pack_kernel<<<vn_blocks(nb, 256), 256, 0, (cudaStream_t)stream>>>(grid, grid_size, decay, apply_decay, threshold, nb, bitfield);
    VN_CHECK_LAUNCH("decay_pack_kernel");
    return VN_OK;
}

// ---- a14, the whole update as ONE host call ------------------------------------------------------------------------
// OccupancyGrid.update (occupancy_grid.py:65-105) after the two batches have been sampled: depth-sensor update
// (_rayUpdate :225-258), NeRF update (_nerfUpdate :261-290 with NGP.density, networks.py:134-148, as hash forward +
// density-only fused MLP), warm-up decay and bitfield repack.  The same kernels in the same order as the module-level
// calls (bit-identical grid), enqueued from C with a caller-owned workspace: no allocation, no Python between launches.
static inline int64_t align4(int64_t n) { return (n + 3) / 4 * 4; }

VN_API int64_t vn_occ_update_ws_floats(int64_t N_ray, int64_t N_nerf, int M) {
    if (N_ray < 0 || N_nerf < 0 || M < 1) return -1;
    const int64_t r = N_ray * M, n = N_nerf * M;
    // ray: idx 3r | po r | pe r | tmp r          nerf: pos 3n | unit 3n | idx 3n | enc 32n | sigma n | po n | pe n | tmp n
    return align4(3 * r) + 3 * align4(r) + 3 * align4(3 * n) + align4(32 * n) + 4 * align4(n) + 1024;
}

VN_API int vn_occ_update(float* grid, int grid_size, uint8_t* bitfield, int32_t* winner, const float* r_rays_o,
                         const float* r_rays_d, const float* r_meas, int64_t N_ray, const float* n_rays_o,
                         const float* n_rays_d, const float* n_noise, int64_t N_nerf, int M, int I, float scale,
                         float noise_every_m, float p_false, float std_every_m, float prob_min, double nerf_thr_max,
                         float nerf_slope, float decay, int apply_decay, float threshold, const float* table,
                         void* table_h, const vn_hash_levels_t* lv, int hash_flags, const float* W1, const float* W2,
                         float xyz_min, float xyz_max, float* ws, int64_t ws_floats, void* stream) {
    VN_REQUIRE(grid && bitfield && winner && ws, "vn_occ_update: null pointer");
    VN_REQUIRE(N_ray >= 0 && N_nerf >= 0 && M >= 2 && I >= 2, "vn_occ_update: bad sizes");
    VN_REQUIRE(N_ray == 0 || (r_rays_o && r_rays_d && r_meas), "vn_occ_update: null ray-update batch");
    VN_REQUIRE(N_nerf == 0 || (n_rays_o && n_rays_d && n_noise && table && lv && W1 && W2), "vn_occ_update: null NeRF-update input");
    VN_REQUIRE(ws_floats >= vn_occ_update_ws_floats(N_ray, N_nerf, M), "vn_occ_update: workspace too small (vn_occ_update_ws_floats)");
    VN_REQUIRE(vn_aligned(ws, 16), "vn_occ_update: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t r = N_ray * M, n = N_nerf * M;
    float* p = ws;
    int32_t* r_idx = (int32_t*)p; p += align4(3 * r);
    float* r_po = p; p += align4(r);
    float* r_pe = p; p += align4(r);
    float* r_tmp = p; p += align4(r);
    float* n_pos = p; p += align4(3 * n);
    float* n_unit = p; p += align4(3 * n);
    int32_t* n_idx = (int32_t*)p; p += align4(3 * n);
    float* n_enc = p; p += align4(32 * n);
    float* n_sig = p; p += align4(n);
    float* n_po = p; p += align4(n);
    float* n_pe = p; p += align4(n);
    float* n_tmp = p; p += align4(n);
    float* scratch = p;                                    // 1024 floats: vn_occ_nerf_prob's partial sums
    if (N_ray > 0) {       // _rayUpdate: _calcPos (no noise) + _rayProb in one kernel, then _updateGrid
        occ_calc_pos_prob_kernel<<<vn_blocks(r, 128), 128, 0, st>>>(r_rays_o, r_rays_d, nullptr, r_meas, N_ray, M, I, grid_size,
                                                                   scale, noise_every_m, p_false, std_every_m, prob_min, nullptr,
                                                                   nullptr, r_idx, r_po, r_pe);
        VN_CHECK_LAUNCH("occ_calc_pos_prob_kernel");
        int rc = vn_occ_bayes_update(grid, grid_size, r_idx, r, r_po, r_pe, winner, r_tmp, stream);
        if (rc) return rc;
    }
    if (N_nerf > 0) {      // _nerfUpdate: _calcPos with position noise -> NGP.density -> _nerfProb -> _updateGrid
        occ_calc_pos_prob_kernel<<<vn_blocks(n, 128), 128, 0, st>>>(n_rays_o, n_rays_d, n_noise, nullptr, N_nerf, M, I, grid_size,
                                                                   scale, noise_every_m, p_false, std_every_m, prob_min, nullptr,
                                                                   n_pos, n_idx, nullptr, nullptr, n_unit, xyz_min, xyz_max);
        VN_CHECK_LAUNCH("occ_calc_pos_prob_kernel");
        int rc;
        if (table_h) {     // half-precision encoder: fp16 copy of the table per call (hash_encoder_half.py:367), fp16 rows
            rc = vn_f32_to_f16(table, table_h, 2 * lv->total_entries, stream); if (rc) return rc;
            rc = vn_hash_encode_fwd_f16(n_unit, table_h, n_enc, n, lv, hash_flags, stream); if (rc) return rc;
            rc = vn_mlp_fwd(n_enc, 1, nullptr, W1, W2, nullptr, nullptr, nullptr, n, 1, n_sig, nullptr, nullptr, stream);
        } else {
            rc = vn_hash_encode_fwd_f32(n_unit, table, n_enc, n, lv, hash_flags, stream); if (rc) return rc;
            rc = vn_mlp_fwd(n_enc, 0, nullptr, W1, W2, nullptr, nullptr, nullptr, n, 1, n_sig, nullptr, nullptr, stream);
        }
        if (rc) return rc;
        rc = vn_occ_nerf_prob(n_sig, n, nerf_thr_max, nerf_slope, scratch, n_po, n_pe, stream); if (rc) return rc;
        rc = vn_occ_bayes_update(grid, grid_size, n_idx, n, n_po, n_pe, winner, n_tmp, stream); if (rc) return rc;
    }
    return vn_occ_decay_pack(grid, grid_size, decay, apply_decay, threshold, bitfield, stream);
}
