// ray_march.cu -- ray/AABB test and occupancy-bitfield ray marching.
// Replaces modules/intersection.py:8-37 and modules/ray_march.py:9-124, 198-269, 328-335.
//
// The marching arithmetic is index-critical (sample counts and (t, dt) sequences must be
// bit-exact against the oracle), so every float expression is written with explicit
// round-to-nearest intrinsics in the reference's source order (no FMA contraction).
#include "common.cuh"
#include "mlp_common.cuh"

struct MarchCfg {
    int cascades;
    int G;
    uint32_t G3;
    float scale, esf, dt_max, G_inv, Gf, Gm1;
};

static MarchCfg make_cfg(int cascades, int grid_size, float scale, float esf) {
    MarchCfg c;
    c.cascades = cascades;
    c.G = grid_size;
    c.G3 = (uint32_t)grid_size * (uint32_t)grid_size * (uint32_t)grid_size;
    c.scale = scale;
    c.esf = esf;
    volatile float a = VN_SQRT3_2 * scale;   // utils.py:56-57: (SQRT3_2 * scale) / grid_size in f32
    volatile float b = a / (float)grid_size;
    c.dt_max = b;
    volatile float gi = 1.0f / (float)grid_size;  // ray_march.py:38
    c.G_inv = gi;
    c.Gf = (float)grid_size;
    c.Gm1 = (float)grid_size - 1.0f;
    return c;
}

// ---- a5 ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ray_aabb_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                       float scale, int64_t N, float2* __restrict__ hits_t) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    float t1 = -INFINITY, t2 = INFINITY;
    const float half_size = vn_div(vn_sub(scale, -scale), 2.0f);     // intersection.py:17
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float o = __ldg(rays_o + 3 * r + k), d = __ldg(rays_d + 3 * r + k);
        const float inv_d = vn_div(1.0f, d);                          // :24
        const float tmin = vn_mul(vn_sub(vn_sub(0.0f, half_size), o), inv_d);   // :26
        const float tmax = vn_mul(vn_sub(vn_add(0.0f, half_size), o), inv_d);   // :27
        t1 = fmaxf(t1, fminf(tmin, tmax));                            // :29-32
        t2 = fminf(t2, fmaxf(tmin, tmax));
    }
    hits_t[r] = (t2 > 0.0f) ? make_float2(fmaxf(t1, VN_NEAR_DISTANCE), t2) : make_float2(-1.0f, -1.0f);
}

VN_API int vn_ray_aabb(const float* rays_o, const float* rays_d, float scale, int64_t N, float* hits_t, void* stream) {
    VN_REQUIRE(N >= 0 && (N == 0 || (rays_o && rays_d && hits_t)), "vn_ray_aabb: bad arguments");
    VN_REQUIRE(vn_aligned(hits_t, 8), "vn_ray_aabb: hits_t must be 8-byte aligned");
    if (N == 0) return VN_OK;
    ray_aabb_kernel<<<vn_blocks(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, scale, N, (float2*)hits_t);
    VN_CHECK_LAUNCH("ray_aabb_kernel");
    return VN_OK;
}

// ---- shared marching step: ray_march.py:44-75 == :232-267 ---------------------------------
struct Ray {
    float o[3], d[3], dinv[3];
};

__device__ __forceinline__ Ray load_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t r) {
    Ray ray;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ray.o[k] = __ldg(rays_o + 3 * r + k);
        ray.d[k] = __ldg(rays_d + 3 * r + k);
        ray.dinv[k] = vn_div(1.0f, ray.d[k]);   // :33
    }
    return ray;
}

// Returns occupancy at t.  xyz/dt are the sample position and step; when the cell is empty
// *t_next is t advanced past the cell exactly as the reference's inner loop does.
__device__ __forceinline__ bool march_probe(const MarchCfg& c, const Ray& ray, const uint8_t* __restrict__ bitfield,
                                            float t, float* xyz, float* dt_out, float* t_next) {
    float mx = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        xyz[k] = vn_add(ray.o[k], vn_mul(t, ray.d[k]));                // :45
        mx = fmaxf(mx, fabsf(xyz[k]));
    }
    const float dt = vn_calc_dt(t, c.esf, c.dt_max);                  // :46
    int mip = 0;
    if (c.cascades > 1) {                                             // utils.py:78-92
        const int m_pos = min(c.cascades - 1, max(0, vn_frexp_bit(mx) + 1));
        const int m_dt = min(c.cascades - 1, max(0, vn_frexp_bit(vn_mul(dt, c.Gf))));
        mip = max(m_pos, m_dt);
    }
    const float mip_bound = fminf(ldexpf(1.0f, mip - 1), c.scale);    // :50
    const float mip_bound_inv = vn_div(1.0f, mip_bound);              // :51
    float nxyz[3];
    uint32_t ci[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        nxyz[k] = vn_clamp(vn_mul(vn_mul(0.5f, vn_add(vn_mul(xyz[k], mip_bound_inv), 1.0f)), c.Gf), 0.0f, c.Gm1);  // :53-57
        ci[k] = vn_f2u(nxyz[k]);
    }
    const uint32_t idx = (uint32_t)mip * c.G3 + vn_morton3D(ci[0], ci[1], ci[2]);   // :59
    const bool occ = (__ldg(bitfield + (idx >> 3)) & (1u << (idx & 7u))) != 0;       // :60
    *dt_out = dt;
    if (!occ) {
        float tmin = INFINITY;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float sgn = (ray.d[k] > 0.0f) ? 1.0f : ((ray.d[k] < 0.0f) ? -1.0f : 0.0f);
            float v = vn_add(vn_add(nxyz[k], 0.5f), vn_mul(0.5f, sgn));
            v = vn_sub(vn_mul(vn_mul(v, c.G_inv), 2.0f), 1.0f);
            v = vn_mul(vn_sub(vn_mul(v, mip_bound), xyz[k]), ray.dinv[k]);          // :67-68
            tmin = fminf(tmin, v);
        }
        const float t_target = vn_add(t, fmaxf(0.0f, tmin));                         // :70
        float tt = vn_add(t, vn_calc_dt(t, c.esf, c.dt_max));                        // :71
        while (tt < t_target) tt = vn_add(tt, vn_calc_dt(tt, c.esf, c.dt_max));      // :72-73
        *t_next = tt;
    }
    return occ;
}

// NGP.density's input normalisation (networks.py:142): (x - xyz_min) / (xyz_max - xyz_min)
// with xyz_min = -scale, xyz_max = +scale, same IEEE operations as the torch expression
__device__ __forceinline__ void write_unit(const MarchCfg& c, float* dst, const float* xyz) {
    const float den = vn_sub(c.scale, -c.scale);
#pragma unroll
    for (int k = 0; k < 3; ++k) dst[k] = vn_div(vn_sub(xyz[k], -c.scale), den);
}

__device__ __forceinline__ float jittered_t1(const MarchCfg& c, float t1, float noise) {
    if (t1 >= 0.0f) t1 = vn_add(t1, vn_mul(vn_calc_dt(t1, c.esf, c.dt_max), noise));   // :39-41
    return t1;
}

// ---- a6: warp-cooperative marcher (both passes) --------------------------------------------
// The sample positions of a ray form a FIXED lattice t_{k+1} = fl(t_k + dt(t_k)) that does not
// depend on the occupancy (both branches of the reference advance t by calc_dt(t)); occupancy
// only decides which lattice points are emitted and which are skipped without being looked
// at (t < t_target).  One warp marches one ray: every lane regenerates the next 32 lattice
// points with the reference's sequential float adds (bit-identical t), probes "its" point in
// parallel (one bitfield latency per 32 points instead of 32), and the emit / skip chain is
// then resolved with ballots.  Points that turn out to be skipped were probed speculatively;
// that costs bandwidth, not correctness.
__device__ __forceinline__ float lattice32(const MarchCfg& c, float t0, int lane, float* t_after) {
    float t = t0, mine = t0;
    if (c.esf == 0.0f) {
        const float dt = vn_calc_dt(0.0f, 0.0f, c.dt_max);      // constant step
#pragma unroll
        for (int k = 0; k < 32; ++k) { if (k == lane) mine = t; t = vn_add(t, dt); }
    } else {
#pragma unroll 4
        for (int k = 0; k < 32; ++k) { if (k == lane) mine = t; t = vn_add(t, vn_calc_dt(t, c.esf, c.dt_max)); }
    }
    *t_after = t;
    return mine;
}

// occupancy at lattice point t and, when empty, the reference's skip target (ray_march.py:67-70)
__device__ __forceinline__ bool probe_point(const MarchCfg& c, const Ray& ray, const uint8_t* __restrict__ bitfield, float t,
                                            float* xyz, float* dt_out, float* t_target) {
    float mx = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) { xyz[k] = vn_add(ray.o[k], vn_mul(t, ray.d[k])); mx = fmaxf(mx, fabsf(xyz[k])); }   // :45
    const float dt = vn_calc_dt(t, c.esf, c.dt_max);
    int mip = 0;
    if (c.cascades > 1) {
        const int m_pos = min(c.cascades - 1, max(0, vn_frexp_bit(mx) + 1));
        const int m_dt = min(c.cascades - 1, max(0, vn_frexp_bit(vn_mul(dt, c.Gf))));
        mip = max(m_pos, m_dt);
    }
    const float mip_bound = fminf(ldexpf(1.0f, mip - 1), c.scale);
    const float mip_bound_inv = vn_div(1.0f, mip_bound);
    float nxyz[3];
    uint32_t ci[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        nxyz[k] = vn_clamp(vn_mul(vn_mul(0.5f, vn_add(vn_mul(xyz[k], mip_bound_inv), 1.0f)), c.Gf), 0.0f, c.Gm1);
        ci[k] = vn_f2u(nxyz[k]);
    }
    const uint32_t idx = (uint32_t)mip * c.G3 + vn_morton3D(ci[0], ci[1], ci[2]);
    const bool occ = (__ldg(bitfield + (idx >> 3)) & (1u << (idx & 7u))) != 0;
    *dt_out = dt;
    float tmin = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float sgn = (ray.d[k] > 0.0f) ? 1.0f : ((ray.d[k] < 0.0f) ? -1.0f : 0.0f);
        float v = vn_add(vn_add(nxyz[k], 0.5f), vn_mul(0.5f, sgn));
        v = vn_sub(vn_mul(vn_mul(v, c.G_inv), 2.0f), 1.0f);
        v = vn_mul(vn_sub(vn_mul(v, mip_bound), xyz[k]), ray.dinv[k]);
        tmin = fminf(tmin, v);
    }
    *t_target = vn_add(t, fmaxf(0.0f, tmin));
    return occ;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) march_warp_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                         const float2* __restrict__ hits_t, const uint8_t* __restrict__ bitfield,
                                                         const float* __restrict__ noise, int64_t N, const MarchCfg c,
                                                         int max_samples, int32_t* __restrict__ counts,
                                                         const int32_t* __restrict__ rays_a, int64_t capacity,
                                                         float* __restrict__ xyzs, float* __restrict__ dirs,
                                                         float* __restrict__ deltas, float* __restrict__ ts,
                                                         float* __restrict__ xyzs_unit) {
    const unsigned full = 0xffffffffu;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    int n_max = max_samples;
    int64_t start = 0;
    if (WRITE) {
        start = rays_a[3 * r + 1];
        n_max = rays_a[3 * r + 2];
        if (n_max == 0) return;
    }
    const Ray ray = load_ray(rays_o, rays_d, r);
    const float2 h = __ldg(hits_t + r);
    float t = jittered_t1(c, h.x, __ldg(noise + r));
    const float t2 = h.y;
    int n = 0;
    bool pending = false;
    float pending_target = 0.0f;
    while (0.0f <= t && t < t2 && n < n_max) {                       // ray_march.py:44 (warp-uniform)
        float t_after;
        const float ti = lattice32(c, t, lane, &t_after);
        const bool inr = ti < t2;
        float xyz[3] = {0.f, 0.f, 0.f}, dt = 0.0f, ttar = 0.0f;
        bool occ = false;
        if (inr) occ = probe_point(c, ray, bitfield, ti, xyz, &dt, &ttar);
        const unsigned inmask = __ballot_sync(full, inr);
        const unsigned occmask = __ballot_sync(full, inr && occ);
        // ---- emit / skip chain, resolved in parallel ------------------------------------------
        // successor of every lattice point: the next point if it is occupied, else the first later
        // point with t >= its skip target (binary search over the increasing lattice; 32 = beyond
        // this chunk); out-of-range points are terminal
        int nxt = lane + 1;
        if (!inr) nxt = 32;
        {
            int lo = lane + 1, hi = 32;
            const bool search = inr && !occ;
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int mid = (lo + hi) >> 1;
                const float tm = __shfl_sync(full, ti, mid & 31);
                if (search && lo < hi) { if (tm >= ttar) hi = mid; else lo = mid + 1; }
            }
            if (search) nxt = lo;                                    // >= lane + 1: at least one step (:71)
        }
        int cur0 = 0;
        if (pending) {                                               // skip started in a previous chunk
            const unsigned ge = __ballot_sync(full, ti >= pending_target);
            if (ge == 0u) cur0 = 32; else { cur0 = __ffs(ge) - 1; pending = false; }
        }
        // points visited from cur0: reachability by pointer doubling (5 rounds cover 32 points)
        unsigned R = (cur0 < 32) ? (1u << cur0) : 0u;
        int jump = nxt;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const unsigned add = (((R >> lane) & 1u) && jump < 32) ? (1u << jump) : 0u;
            R |= __reduce_or_sync(full, add);
            const int j2 = __shfl_sync(full, jump, jump & 31);
            jump = (jump < 32) ? j2 : 32;
        }
        bool done = (R & ~inmask) != 0u;                             // reached a point with t >= t2
        unsigned emit = R & occmask;
        const int room = n_max - n;
        if (__popc(emit) >= room) {                                  // N_samples reaches max_samples (:44)
            if (__popc(emit) > room) emit &= (2u << __fns(emit, 0, room)) - 1u;
            done = true;
        }
        if (!done && R != 0u) {                                      // does the chain leave the chunk in a skip?
            const int last = 31 - __clz(R);
            const int lp = __shfl_sync(full, (int)(inr && !occ), last);
            const float lt = __shfl_sync(full, ttar, last);
            if (lp) { pending = true; pending_target = lt; }
        }
        if (WRITE) {
            if ((emit >> lane) & 1u) {
                const int64_t s = start + n + __popc(emit & ((1u << lane) - 1u));
                if (s < capacity) {
                    xyzs[3 * s] = xyz[0]; xyzs[3 * s + 1] = xyz[1]; xyzs[3 * s + 2] = xyz[2];
                    dirs[3 * s] = ray.d[0]; dirs[3 * s + 1] = ray.d[1]; dirs[3 * s + 2] = ray.d[2];
                    ts[s] = ti; deltas[s] = dt;
                    if (xyzs_unit) write_unit(c, xyzs_unit + 3 * s, xyz);
                }
            }
        }
        if (!WRITE && ts != nullptr && ((emit >> lane) & 1u))      // single-pass mode: row r of ts [N, max_samples]
            ts[r * max_samples + n + __popc(emit & ((1u << lane) - 1u))] = ti;
        n += __popc(emit);
        if (done) break;
        t = t_after;
    }
    if (!WRITE && lane == 0) counts[r] = n;
}

// thread-per-ray variant of both passes: preferable when there are enough rays to fill the
// machine and most lattice points are skipped (carved scenes), because skipped points are
// never probed.  Bit-identical results; the launcher picks by ray count.
template <bool WRITE>
__global__ void __launch_bounds__(128) march_thread_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                           const float2* __restrict__ hits_t, const uint8_t* __restrict__ bitfield,
                                                           const float* __restrict__ noise, int64_t N, const MarchCfg c,
                                                           int max_samples, int32_t* __restrict__ counts,
                                                           const int32_t* __restrict__ rays_a, int64_t capacity,
                                                           float* __restrict__ xyzs, float* __restrict__ dirs,
                                                           float* __restrict__ deltas, float* __restrict__ ts,
                                                           float* __restrict__ xyzs_unit) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    int n_max = max_samples;
    int64_t start = 0;
    if (WRITE) { start = rays_a[3 * r + 1]; n_max = rays_a[3 * r + 2]; if (n_max == 0) return; }
    const Ray ray = load_ray(rays_o, rays_d, r);
    const float2 h = __ldg(hits_t + r);
    float t = jittered_t1(c, h.x, __ldg(noise + r));
    const float t2 = h.y;
    int n = 0;
    while (0.0f <= t && t < t2 && n < n_max) {                       // :44 / :87
        float xyz[3], dt, tn;
        if (march_probe(c, ray, bitfield, t, xyz, &dt, &tn)) {
            if (WRITE) {
                const int64_t s = start + n;
                if (s < capacity) {
                    xyzs[3 * s] = xyz[0]; xyzs[3 * s + 1] = xyz[1]; xyzs[3 * s + 2] = xyz[2];
                    dirs[3 * s] = ray.d[0]; dirs[3 * s + 1] = ray.d[1]; dirs[3 * s + 2] = ray.d[2];
                    ts[s] = t; deltas[s] = dt;
                    if (xyzs_unit) write_unit(c, xyzs_unit + 3 * s, xyz);
                }
            }
            if (!WRITE && ts != nullptr) ts[r * max_samples + n] = t;
            t = vn_add(t, dt); ++n;
        } else t = tn;
    }
    if (!WRITE) counts[r] = n;
}

// rays at or above this count use the thread-per-ray marcher
static const int64_t kWarpMarchMaxRays = 16384;

// ---- exclusive scan (i32), hierarchical: 1024 elements per block ---------------------------
__global__ void __launch_bounds__(256) scan_block_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                         int32_t* __restrict__ block_sums, int64_t n) {
    __shared__ int32_t warp_sums[8];
    const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
    int32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (base + k < n) ? in[base + k] : 0;
    const int32_t tsum = v[0] + v[1] + v[2] + v[3];
    int32_t inc = tsum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int32_t o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    int32_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_sums[w];
    int32_t run = woff + inc - tsum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 255 && block_sums) block_sums[blockIdx.x] = woff + inc;
}

__global__ void __launch_bounds__(256) scan_add_kernel(int32_t* __restrict__ out, const int32_t* __restrict__ block_offsets, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
    const int32_t off = block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i + k < n) out[i + k] += off;
}

static int64_t round_up4(int64_t x) { return (x + 3) / 4 * 4; }

static int64_t scan_hier_ints(int64_t N) {
    int64_t total = 0, n = N;
    while (n > 1024) { n = (n + 1023) / 1024; total += 2 * round_up4(n); }
    return total + 8;
}

// scratch for the count pass: N ints for the scanned starts + the scan hierarchy
VN_API int64_t vn_march_scan_tmp_ints(int64_t N) { return round_up4(N < 0 ? 0 : N) + scan_hier_ints(N); }

// out may alias in
static int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* tmp, cudaStream_t st) {
    if (n <= 0) return VN_OK;
    const int64_t nb = (n + 1023) / 1024;
    if (nb == 1) {
        scan_block_kernel<<<1, 256, 0, st>>>(in, out, nullptr, n);
        VN_CHECK_LAUNCH("scan_block_kernel");
        return VN_OK;
    }
    int32_t* sums = tmp;
    int32_t* sums_scanned = tmp + round_up4(nb);
    scan_block_kernel<<<(unsigned)nb, 256, 0, st>>>(in, out, sums, n);
    VN_CHECK_LAUNCH("scan_block_kernel");
    int rc = exclusive_scan_i32(sums, sums_scanned, nb, tmp + 2 * round_up4(nb), st);
    if (rc) return rc;
    scan_add_kernel<<<(unsigned)nb, 256, 0, st>>>(out, sums_scanned, n);
    VN_CHECK_LAUNCH("scan_add_kernel");
    return VN_OK;
}

// rays_a[r] = (r, start_r, count_r); counter = (total, N): the deterministic equivalent of
// ray_march.py:77-82 (atomic counters) in canonical ray order.
__global__ void __launch_bounds__(256) fill_rays_a_kernel(const int32_t* __restrict__ counts,
                                                          const int32_t* __restrict__ starts, int64_t N,
                                                          int32_t* __restrict__ rays_a, int32_t* __restrict__ counter) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const int32_t cnt = counts[r], st = starts[r];
    rays_a[3 * r] = (int32_t)r;
    rays_a[3 * r + 1] = st;
    rays_a[3 * r + 2] = cnt;
    if (r == N - 1) { counter[0] = st + cnt; counter[1] = (int32_t)N; }
}

static int march_count_impl(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* bitfield,
                            const float* noise, int64_t N, int cascades, int grid_size, float scale,
                            float exp_step_factor, int max_samples, int32_t* counts, int32_t* rays_a,
                            int32_t* counter, int32_t* scan_tmp, float* ts_rows, void* stream) {
    VN_REQUIRE(N >= 0, "vn_march_train_count: N < 0");
    VN_REQUIRE(counter != nullptr, "vn_march_train_count: null counter");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) { VN_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(int32_t), st)); return VN_OK; }
    VN_REQUIRE(rays_o && rays_d && hits_t && bitfield && noise && counts && rays_a && scan_tmp,
               "vn_march_train_count: null pointer");
    VN_REQUIRE(vn_aligned(hits_t, 8), "vn_march_train_count: hits_t must be 8-byte aligned");
    VN_REQUIRE(cascades >= 1 && grid_size >= 1 && grid_size <= 1024 && max_samples >= 0,
               "vn_march_train_count: bad cascades/grid_size/max_samples");
    const MarchCfg c = make_cfg(cascades, grid_size, scale, exp_step_factor);
    VnProfScope prof(VN_K_MARCH_COUNT, N, st);
    if (N < kWarpMarchMaxRays)
        march_warp_kernel<false><<<vn_blocks(N * 32, 256), 256, 0, st>>>(rays_o, rays_d, (const float2*)hits_t, bitfield,
                                                                         noise, N, c, max_samples, counts, nullptr, 0,
                                                                         nullptr, nullptr, nullptr, ts_rows, nullptr);
    else
        march_thread_kernel<false><<<vn_blocks(N, 128), 128, 0, st>>>(rays_o, rays_d, (const float2*)hits_t, bitfield, noise,
                                                                      N, c, max_samples, counts, nullptr, 0, nullptr,
                                                                      nullptr, nullptr, ts_rows, nullptr);
    VN_CHECK_LAUNCH("march kernel <count>");
    int32_t* starts = scan_tmp;
    int rc = exclusive_scan_i32(counts, starts, N, scan_tmp + round_up4(N), st);
    if (rc) return rc;
    fill_rays_a_kernel<<<vn_blocks(N, 256), 256, 0, st>>>(counts, starts, N, rays_a, counter);
    VN_CHECK_LAUNCH("fill_rays_a_kernel");
    return VN_OK;
}

VN_API int vn_march_train_count(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* bitfield,
                                const float* noise, int64_t N, int cascades, int grid_size, float scale,
                                float exp_step_factor, int max_samples, int32_t* counts, int32_t* rays_a,
                                int32_t* counter, int32_t* scan_tmp, void* stream) {
    return march_count_impl(rays_o, rays_d, hits_t, bitfield, noise, N, cascades, grid_size, scale, exp_step_factor,
                            max_samples, counts, rays_a, counter, scan_tmp, nullptr, stream);
}

// ---- a6, single-pass variant: the count pass also records the t of every emitted sample in
// row r of ts_rows [N, max_samples]; pass 2 then needs no second march -- every output of
// ray_march.py:84-124 is a function of (ray, t): xyz = o + t d (:45), dt = calc_dt(t) (:46).
VN_API int vn_march_train_count_rows(const float* rays_o, const float* rays_d, const float* hits_t,
                                     const uint8_t* bitfield, const float* noise, int64_t N, int cascades, int grid_size,
                                     float scale, float exp_step_factor, int max_samples, int32_t* counts,
                                     int32_t* rays_a, int32_t* counter, int32_t* scan_tmp, float* ts_rows, void* stream) {
    VN_REQUIRE(N == 0 || ts_rows != nullptr, "vn_march_train_count_rows: null ts_rows");
    return march_count_impl(rays_o, rays_d, hits_t, bitfield, noise, N, cascades, grid_size, scale, exp_step_factor,
                            max_samples, counts, rays_a, counter, scan_tmp, ts_rows, stream);
}

// warp per ray: sample k of ray r -> row rays_a[r,1] + k of the packed outputs
__global__ void __launch_bounds__(256) march_expand_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                           const int32_t* __restrict__ rays_a, const float* __restrict__ ts_rows,
                                                           int64_t N, int max_samples, const MarchCfg c, int64_t capacity,
                                                           float* __restrict__ xyzs, float* __restrict__ dirs,
                                                           float* __restrict__ deltas, float* __restrict__ ts,
                                                           float* __restrict__ xyzs_unit, uint4* __restrict__ sh_planes,
                                                           int64_t sh_stride) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    const int64_t start = rays_a[3 * r + 1];
    const int n = rays_a[3 * r + 2];
    if (n == 0) return;
    float o[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = __ldg(rays_o + 3 * r + k); d[k] = __ldg(rays_d + 3 * r + k); }
    // the direction encoding is a function of the RAY: evaluate SH16((d/|d| + 1)/2) once (networks.py:160-161,
    // spherical_harmonics.py:16-42 -- the expression sequence of the fused MLP's input stage) and hand every
    // sample its copy as two fp16 operand chunks, so that the MLP kernels bulk-copy instead of recomputing it
    uint4 sh_lo = make_uint4(0u, 0u, 0u, 0u), sh_hi = sh_lo;
    if (sh_planes) {
        const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        float e[16];
        mlp::sh16_half((d[0] / nrm + 1.0f) * 0.5f, (d[1] / nrm + 1.0f) * 0.5f, (d[2] / nrm + 1.0f) * 0.5f, e);
        __half2 h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = __floats2half2_rn(e[2 * j], e[2 * j + 1]);
        sh_lo = *reinterpret_cast<const uint4*>(h); sh_hi = *reinterpret_cast<const uint4*>(h + 4);
    }
    const float* row = ts_rows + r * max_samples;
    for (int k = lane; k < n; k += 32) {
        const int64_t s = start + k;
        if (s >= capacity) break;
        const float t = __ldg(row + k);
        float xyz[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) xyz[j] = vn_add(o[j], vn_mul(t, d[j]));               // ray_march.py:45
        xyzs[3 * s] = xyz[0]; xyzs[3 * s + 1] = xyz[1]; xyzs[3 * s + 2] = xyz[2];
        dirs[3 * s] = d[0]; dirs[3 * s + 1] = d[1]; dirs[3 * s + 2] = d[2];
        ts[s] = t; deltas[s] = vn_calc_dt(t, c.esf, c.dt_max);                            // :46
        if (xyzs_unit) write_unit(c, xyzs_unit + 3 * s, xyz);
        if (sh_planes) { sh_planes[s] = sh_lo; sh_planes[sh_stride + s] = sh_hi; }
    }
}

static int march_expand_launch(const float* rays_o, const float* rays_d, const int32_t* rays_a, const float* ts_rows,
                               int64_t N, int max_samples, int grid_size, float scale, float exp_step_factor,
                               int64_t capacity, float* xyzs, float* dirs, float* deltas, float* ts, float* xyzs_unit,
                               void* sh_planes, int64_t sh_stride, void* stream);

VN_API int vn_march_train_expand(const float* rays_o, const float* rays_d, const int32_t* rays_a, const float* ts_rows,
                                 int64_t N, int max_samples, int grid_size, float scale, float exp_step_factor,
                                 int64_t capacity, float* xyzs, float* dirs, float* deltas, float* ts, float* xyzs_unit,
                                 void* stream) {
    return march_expand_launch(rays_o, rays_d, rays_a, ts_rows, N, max_samples, grid_size, scale, exp_step_factor, capacity,
                               xyzs, dirs, deltas, ts, xyzs_unit, nullptr, 0, stream);
}

VN_API int vn_march_train_expand_sh(const float* rays_o, const float* rays_d, const int32_t* rays_a, const float* ts_rows,
                                    int64_t N, int max_samples, int grid_size, float scale, float exp_step_factor,
                                    int64_t capacity, float* xyzs, float* dirs, float* deltas, float* ts, float* xyzs_unit,
                                    void* sh_planes, int64_t sh_stride, void* stream) {
    VN_REQUIRE(sh_planes && vn_aligned(sh_planes, 16) && sh_stride >= capacity,
               "vn_march_train_expand_sh: sh_planes must be 16-byte aligned with a plane stride >= capacity");
    return march_expand_launch(rays_o, rays_d, rays_a, ts_rows, N, max_samples, grid_size, scale, exp_step_factor, capacity,
                               xyzs, dirs, deltas, ts, xyzs_unit, sh_planes, sh_stride, stream);
}

static int march_expand_launch(const float* rays_o, const float* rays_d, const int32_t* rays_a, const float* ts_rows,
                               int64_t N, int max_samples, int grid_size, float scale, float exp_step_factor,
                               int64_t capacity, float* xyzs, float* dirs, float* deltas, float* ts, float* xyzs_unit,
                               void* sh_planes, int64_t sh_stride, void* stream) {
    VN_REQUIRE(N >= 0 && capacity >= 0 && max_samples >= 0, "vn_march_train_expand: negative size");
    if (N == 0 || capacity == 0) return VN_OK;
    VN_REQUIRE(rays_o && rays_d && rays_a && ts_rows && xyzs && dirs && deltas && ts, "vn_march_train_expand: null pointer");
    VN_REQUIRE(grid_size >= 1 && grid_size <= 1024, "vn_march_train_expand: bad grid_size");
    const MarchCfg c = make_cfg(1, grid_size, scale, exp_step_factor);
    VnProfScope prof(VN_K_MARCH_WRITE, capacity, (cudaStream_t)stream);
    vn_launch_pdl(march_expand_kernel, dim3(vn_blocks(N * 32, 256)), dim3(256), 0, (cudaStream_t)stream, rays_o, rays_d, rays_a, ts_rows, N, max_samples,
                                                                                c, capacity, xyzs, dirs, deltas, ts, xyzs_unit, (uint4*)sh_planes, sh_stride);
    VN_CHECK_LAUNCH("march_expand_kernel");
    return VN_OK;
}

// ---- a6 pass 2: march_warp_kernel<true> -----------------------------------------------------
VN_API int vn_march_train_write(const float* rays_o, const float* rays_d, const float* hits_t, const uint8_t* bitfield,
                                const float* noise, int64_t N, int cascades, int grid_size, float scale,
                                float exp_step_factor, const int32_t* rays_a, int64_t capacity, float* xyzs,
                                float* dirs, float* deltas, float* ts, float* xyzs_unit, void* stream) {
    VN_REQUIRE(N >= 0 && capacity >= 0, "vn_march_train_write: negative size");
    if (N == 0 || capacity == 0) return VN_OK;
    VN_REQUIRE(rays_o && rays_d && hits_t && bitfield && noise && rays_a && xyzs && dirs && deltas && ts,
               "vn_march_train_write: null pointer");
    VN_REQUIRE(vn_aligned(hits_t, 8), "vn_march_train_write: hits_t must be 8-byte aligned");
    const MarchCfg c = make_cfg(cascades, grid_size, scale, exp_step_factor);
    VnProfScope prof(VN_K_MARCH_WRITE, capacity, (cudaStream_t)stream);
    if (N < kWarpMarchMaxRays)
        march_warp_kernel<true><<<vn_blocks(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, (const float2*)hits_t, bitfield, noise, N, c, 0, nullptr, rays_a, capacity, xyzs, dirs, deltas, ts, xyzs_unit);
    else
        march_thread_kernel<true><<<vn_blocks(N, 128), 128, 0, (cudaStream_t)stream>>>(
            rays_o, rays_d, (const float2*)hits_t, bitfield, noise, N, c, 0, nullptr, rays_a, capacity, xyzs, dirs, deltas, ts, xyzs_unit);
    VN_CHECK_LAUNCH("march kernel <write>");
    return VN_OK;
}

// ---- a7 ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) march_test_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                         float2* __restrict__ hits_t, const int64_t* __restrict__ alive,
                                                         int64_t A, const uint8_t* __restrict__ bitfield, const MarchCfg c,
                                                         int max_samples, int64_t* __restrict__ ray_indices,
                                                         uint8_t* __restrict__ valid_mask, float* __restrict__ deltas,
                                                         float* __restrict__ ts, int32_t* __restrict__ counter) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= A) return;
    const int64_t r = alive[n];
    const Ray ray = load_ray(rays_o, rays_d, r);
    const float2 h = hits_t[r];
    float t = h.x;
    const float t2 = h.y;
    int s = 0;
    float t_last = t;
    const int64_t base = n * (int64_t)max_samples;
    while (0.0f < t && t < t2 && s < max_samples) {                   // :231
        float xyz[3], dt, tn;
        if (march_probe(c, ray, bitfield, t, xyz, &dt, &tn)) {
            const int64_t k = base + s;
            ray_indices[k] = r; valid_mask[k] = 1; ts[k] = t; deltas[k] = dt;    // :252-256
            t = vn_add(t, dt); ++s;
            t_last = t;                                                          // :257-258 hits_t[r,0] = t
        } else t = tn;
    }
    if (s > 0) hits_t[r].x = t_last;   // net effect of the per-sample store: t right after the LAST emitted sample
    counter[n] = s;                                                    // :269
}

VN_API int vn_march_test(const float* rays_o, const float* rays_d, float* hits_t, const int64_t* alive, int64_t A,
                         const uint8_t* bitfield, int cascades, int grid_size, float scale, float exp_step_factor,
                         int max_samples, int64_t* ray_indices, uint8_t* valid_mask, float* deltas, float* ts,
                         int32_t* counter, void* stream) {
    VN_REQUIRE(A >= 0 && max_samples >= 0, "vn_march_test: negative size");
    if (A == 0) return VN_OK;
    VN_REQUIRE(rays_o && rays_d && hits_t && alive && bitfield && ray_indices && valid_mask && deltas && ts && counter,
               "vn_march_test: null pointer");
    VN_REQUIRE(vn_aligned(hits_t, 8), "vn_march_test: hits_t must be 8-byte aligned");
    VN_REQUIRE(cascades >= 1 && grid_size >= 1 && grid_size <= 1024, "vn_march_test: bad cascades/grid_size");
    const MarchCfg c = make_cfg(cascades, grid_size, scale, exp_step_factor);
    march_test_kernel<<<vn_blocks(A, 128), 128, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, (float2*)hits_t, alive, A, bitfield, c, max_samples, ray_indices, valid_mask, deltas, ts, counter);
    VN_CHECK_LAUNCH("march_test_kernel");
    return VN_OK;
}

// wrapper compaction, ray_march.py:328-335: packed_info = (cumsum - cnt, cnt); masked arrays
__global__ void __launch_bounds__(256) march_compact_kernel(const int32_t* __restrict__ counter,
                                                            const int32_t* __restrict__ starts, int64_t A, int max_samples,
                                                            const int64_t* __restrict__ ray_indices, const float* __restrict__ deltas,
                                                            const float* __restrict__ ts, int64_t* __restrict__ packed_info,
                                                            int64_t* __restrict__ ray_indices_out, float* __restrict__ deltas_out,
                                                            float* __restrict__ ts_out, int64_t* __restrict__ total) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= A) return;
    const int cnt = counter[n];
    const int64_t st = starts[n];
    packed_info[2 * n] = st;
    packed_info[2 * n + 1] = cnt;
    const int64_t base = n * (int64_t)max_samples;
    for (int s = 0; s < cnt; ++s) {
        ray_indices_out[st + s] = ray_indices[base + s];
        deltas_out[st + s] = deltas[base + s];
        ts_out[st + s] = ts[base + s];
    }
    if (n == A - 1) total[0] = st + cnt;
}

VN_API int vn_march_test_compact(const int32_t* counter, int64_t A, int max_samples, const int64_t* ray_indices,
                                 const float* deltas, const float* ts, int64_t* packed_info, int64_t* ray_indices_out,
                                 float* deltas_out, float* ts_out, int64_t* total, int32_t* scan_tmp, void* stream) {
    VN_REQUIRE(A >= 0 && total != nullptr, "vn_march_test_compact: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (A == 0) { VN_CUDA(cudaMemsetAsync(total, 0, sizeof(int64_t), st)); return VN_OK; }
    VN_REQUIRE(counter && ray_indices && deltas && ts && packed_info && ray_indices_out && deltas_out && ts_out && scan_tmp,
               "vn_march_test_compact: null pointer");
    int32_t* starts = scan_tmp;
    int rc = exclusive_scan_i32(counter, starts, A, scan_tmp + round_up4(A), st);
    if (rc) return rc;
    march_compact_kernel<<<vn_blocks(A, 256), 256, 0, st>>>(counter, starts, A, max_samples, ray_indices, deltas, ts,
                                                            packed_info, ray_indices_out, deltas_out, ts_out, total);
    VN_CHECK_LAUNCH("march_compact_kernel");
    return VN_OK;
}
