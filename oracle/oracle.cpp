// oracle.cpp -- CPU restatement of the VIRUS-NeRF Instant-NGP hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library, and only as the checker / reported CPU baseline.
//
// PARITY STATUS: the reference's kernels are Taichi (taichi_nightly==1.7.0.post20230921,
// requirements.txt:19) which is not installable here, and the reference ships no golden
// vectors.  This restatement is pinned two ways (tests/golden/make_golden.py):
//   * the reference's own @ti.kernel SOURCE is executed under a pure-Python f32 shim
//     (tests/golden/ti_shim) and its outputs are committed as fixtures;
//   * the pure-torch parts of the reference (occupancy_grid.py, geometric_fcts.py,
//     loss.py) are imported and run directly.
// What stays unpinned is Taichi's compiled arithmetic itself (fast-math exp, FMA
// contraction, atomic order) -- fixed here as strict IEEE-754 binary32, round to nearest,
// NO fma contraction (build with -ffp-contract=off), glibc expf/logf, wrapping u32.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fopenmp).
//
// Threading: `threads` <= 1 runs the loops sequentially in source order (the parity
// mode); `threads` > 1 uses OpenMP over the outer (ray / point) loop and is used only
// for the timed CPU baseline.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VO_API extern "C" __attribute__((visibility("default")))

typedef _Float16 half_t;

// ------------------------------------------------------------------------------------
// constants: modules/utils.py:12-16
// ------------------------------------------------------------------------------------
static const int   MAX_SAMPLES = 1024;
static const float NEAR_DISTANCE = 0.01f;
static const float SQRT3_MAX_SAMPLES = (float)(1.7320508075688772 / 1024.0);
static const float SQRT3_2 = (float)(1.7320508075688772 * 2.0);

// exp used by the compositing kernels: correctly rounded binary32 exp (double exp, rounded
// once).  alpha = 1 - exp(-sigma*delta) cancels catastrophically for small sigma*delta, so a
// 1-ulp difference in expf shows up as a 6e-8 ABSOLUTE difference in every weight; fixing the
// exp to the correctly rounded value makes oracle and kernels agree on alpha bit for bit.
static inline float exp_cr(float x) { return (float)std::exp((double)x); }

static inline int vo_threads(int threads) {
#ifdef _OPENMP
    return threads > 1 ? threads : 1;
#else
    (void)threads; return 1;
#endif
}

// ------------------------------------------------------------------------------------
// a1. hash-grid level geometry.  Host part: modules/utils.py:19-42 (res_in_level_np,
// scale_in_level_np, align_to), modules/hash_encoder.py:183-208 (offset loop).
// Kernel part: modules/hash_encoder.py:73-80 (grid_scale / grid_resolution in f32).
// Returns total number of entries (sum of per-level sizes).  `res_host` is the float64
// host resolution, `res_kernel`/`scale_kernel` what the f32 kernel computes; the caller
// (tests) checks they agree.
// ------------------------------------------------------------------------------------
VO_API int64_t vo_hash_levels(double base_res, double max_res, int levels, int64_t max_params,
                              int32_t* offsets, int32_t* sizes, float* scale_kernel,
                              uint32_t* res_kernel, double* res_host, int32_t* begin_fast_hash_level,
                              double* log_b_out) {
    double log_b = std::log(max_res / base_res) / (double)(levels - 1);  // utils.py:31-40
    if (log_b_out) *log_b_out = log_b;
    int64_t offset = 0;
    int begin_fast = levels;
    for (int i = 0; i < levels; ++i) {
        double r = std::ceil(base_res * std::exp((double)i * log_b) - 1.0) + 1.0;  // utils.py:19-29
        double full = r * r * r;
        int64_t full_aligned = (int64_t)((full + 8 - 1) / 8) * 8;  // align_to, utils.py:42
        int64_t size_i = std::min<int64_t>(max_params, full_aligned);
        offsets[i] = (int32_t)offset;
        sizes[i] = (int32_t)size_i;
        if (full > (double)size_i && begin_fast == levels) begin_fast = i;  // hash_encoder.py:200-203
        offset += size_i;
        // kernel-side f32 geometry, hash_encoder.py:73-80
        float ls = (float)log_b;
        float e = expf((float)i * ls);
        float sc = (float)base_res * e - 1.0f;
        scale_kernel[i] = sc;
        res_kernel[i] = (uint32_t)ceilf(sc) + 1u;
        if (res_host) res_host[i] = r;
    }
    *begin_fast_hash_level = begin_fast;
    return offset;
}

// hash_encoder.py:43-71 (fast_hash, under_hash, grid_pos2hash_index)
static inline uint32_t hash_index(bool dense, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t res,
                                  uint32_t map_size) {
    uint32_t h;
    if (dense) {
        uint32_t stride = 1u;
        h = 0u;
        h += c0 * stride; stride *= res;
        h += c1 * stride; stride *= res;
        h += c2 * stride;
    } else {
        h = (c0 * 1u) ^ (c1 * 2654435761u) ^ (c2 * 805459861u);
    }
    return h % map_size;
}

// CUDA-style saturating float -> u32 cast (negative / NaN -> 0); the reference leaves
// this to LLVM fptoui (undefined for negatives) -- see DESIGN.md "oracle decisions".
static inline uint32_t f2u_sat(float x) {
    if (!(x > 0.0f)) return 0u;
    if (x >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)x;
}

struct Corner { uint32_t idx[8]; float w[8]; };

// one (point, level) of hash_encoder.py:97-133
static inline void hash_corners(const float* xyz, float scale, uint32_t res, uint32_t map_size,
                                bool dense, Corner& c) {
    float pos[3]; uint32_t g[3];
    for (int d = 0; d < 3; ++d) {
        float p = xyz[d] * scale + 0.5f;          // :106 (no contraction: -ffp-contract=off)
        float fl = floorf(p);
        g[d] = f2u_sat(fl);                        // :107
        pos[d] = p - (float)g[d];                  // :108
    }
    for (int idx = 0; idx < 8; ++idx) {            // :114-133
        float w = 1.0f;
        uint32_t cl[3];
        for (int d = 0; d < 3; ++d) {
            if ((idx & (1 << d)) == 0) { cl[d] = g[d]; w *= 1.0f - pos[d]; }
            else                       { cl[d] = g[d] + 1u; w *= pos[d]; }
        }
        c.idx[idx] = hash_index(dense, cl[0], cl[1], cl[2], res, map_size);
        c.w[idx] = w;
    }
}

// debug/KAT: the 8 corner indices (level-local, before adding offsets) and weights
VO_API void vo_hash_indices(const float* xyz, int64_t S, int levels, const int32_t* sizes,
                            const float* scales, const uint32_t* res, int begin_fast,
                            int32_t* idx_out /*[S,L,8]*/, float* w_out /*[S,L,8] or null*/) {
    for (int64_t i = 0; i < S; ++i)
        for (int l = 0; l < levels; ++l) {
            Corner c;
            hash_corners(xyz + 3 * i, scales[l], res[l], (uint32_t)sizes[l], l < begin_fast, c);
            for (int k = 0; k < 8; ++k) {
                idx_out[(i * levels + l) * 8 + k] = (int32_t)c.idx[k];
                if (w_out) w_out[(i * levels + l) * 8 + k] = c.w[k];
            }
        }
}

// a2. hash_encoder_kernel, modules/hash_encoder.py:89-143 (F = 2)
VO_API void vo_hash_fwd_f32(const float* xyz, const float* table, float* out, int64_t S, int levels,
                            const int32_t* offsets, const int32_t* sizes, const float* scales,
                            const uint32_t* res, int begin_fast, int threads) {
    int nt = vo_threads(threads);
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
    for (int64_t i = 0; i < S; ++i) {
        for (int l = 0; l < levels; ++l) {
            Corner c;
            hash_corners(xyz + 3 * i, scales[l], res[l], (uint32_t)sizes[l], l < begin_fast, c);
            const float* tb = table + 2 * (int64_t)offsets[l];
            float a0 = 0.0f, a1 = 0.0f;
            for (int k = 0; k < 8; ++k) {           // :137-138
                a0 += c.w[k] * tb[2 * (int64_t)c.idx[k] + 0];
                a1 += c.w[k] * tb[2 * (int64_t)c.idx[k] + 1];
            }
            out[i * (2 * levels) + 2 * l + 0] = a0;  // :140-142
            out[i * (2 * levels) + 2 * l + 1] = a1;
        }
    }
}

// a3. reverse mode of the above (hash_encoder.py:264-277): table.grad += w * dout
VO_API void vo_hash_bwd_f32(const float* xyz, const float* dout, float* grad, int64_t S, int levels,
                            const int32_t* offsets, const int32_t* sizes, const float* scales,
                            const uint32_t* res, int begin_fast, int threads) {
    int nt = vo_threads(threads);
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
    for (int64_t i = 0; i < S; ++i) {
        for (int l = 0; l < levels; ++l) {
            Corner c;
            hash_corners(xyz + 3 * i, scales[l], res[l], (uint32_t)sizes[l], l < begin_fast, c);
            float* gb = grad + 2 * (int64_t)offsets[l];
            float d0 = dout[i * (2 * levels) + 2 * l + 0];
            float d1 = dout[i * (2 * levels) + 2 * l + 1];
            for (int k = 0; k < 8; ++k) {
                float v0 = c.w[k] * d0, v1 = c.w[k] * d1;
                if (nt > 1) {
#pragma omp atomic
                    gb[2 * (int64_t)c.idx[k] + 0] += v0;
#pragma omp atomic
                    gb[2 * (int64_t)c.idx[k] + 1] += v1;
                } else {
                    gb[2 * (int64_t)c.idx[k] + 0] += v0;
                    gb[2 * (int64_t)c.idx[k] + 1] += v1;
                }
            }
        }
    }
}

// a4. half encoder forward, modules/hash_encoder_half.py:112-161.  table is the fp16 copy
// (hash_table.to(float16), :367); weights f32; local_features (f16) += f16(w * table) (:159)
VO_API void vo_hash_fwd_f16(const float* xyz, const uint16_t* table_h, uint16_t* out_h, int64_t S,
                            int levels, const int32_t* offsets, const int32_t* sizes,
                            const float* scales, const uint32_t* res, int begin_fast, int threads) {
    const half_t* table = (const half_t*)table_h;
    half_t* out = (half_t*)out_h;
    int nt = vo_threads(threads);
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
    for (int64_t i = 0; i < S; ++i) {
        for (int l = 0; l < levels; ++l) {
            Corner c;
            hash_corners(xyz + 3 * i, scales[l], res[l], (uint32_t)sizes[l], l < begin_fast, c);
            const half_t* tb = table + 2 * (int64_t)offsets[l];
            half_t a0 = (half_t)0.0f, a1 = (half_t)0.0f;
            for (int k = 0; k < 8; ++k) {
                half_t p0 = (half_t)(c.w[k] * (float)tb[2 * (int64_t)c.idx[k] + 0]);
                half_t p1 = (half_t)(c.w[k] * (float)tb[2 * (int64_t)c.idx[k] + 1]);
                a0 = (half_t)((float)a0 + (float)p0);
                a1 = (half_t)((float)a1 + (float)p1);
            }
            out[i * (2 * levels) + 2 * l + 0] = a0;
            out[i * (2 * levels) + 2 * l + 1] = a1;
        }
    }
}

// a4. half encoder backward, hash_encoder_half.py:164-213.  dout is f16; the product
// w * dout is formed in f32 (:211); skipped when dout is all-zero or w*dout is all-zero
// (:210-213); accumulated into the f32 `hash_grad` buffer (:300-306).
VO_API void vo_hash_bwd_f16(const float* xyz, const uint16_t* dout_h, float* grad, int64_t S,
                            int levels, const int32_t* offsets, const int32_t* sizes,
                            const float* scales, const uint32_t* res, int begin_fast, int threads) {
    const half_t* dout = (const half_t*)dout_h;
    int nt = vo_threads(threads);
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
    for (int64_t i = 0; i < S; ++i) {
        for (int l = 0; l < levels; ++l) {
            float d0 = (float)dout[i * (2 * levels) + 2 * l + 0];
            float d1 = (float)dout[i * (2 * levels) + 2 * l + 1];
            if (d0 == 0.0f && d1 == 0.0f) continue;
            Corner c;
            hash_corners(xyz + 3 * i, scales[l], res[l], (uint32_t)sizes[l], l < begin_fast, c);
            float* gb = grad + 2 * (int64_t)offsets[l];
            for (int k = 0; k < 8; ++k) {
                float v0 = c.w[k] * d0, v1 = c.w[k] * d1;
                if (v0 == 0.0f && v1 == 0.0f) continue;
                if (nt > 1) {
#pragma omp atomic
                    gb[2 * (int64_t)c.idx[k] + 0] += v0;
#pragma omp atomic
                    gb[2 * (int64_t)c.idx[k] + 1] += v1;
                } else {
                    gb[2 * (int64_t)c.idx[k] + 0] += v0;
                    gb[2 * (int64_t)c.idx[k] + 1] += v1;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// a5. ray_aabb_intersect, modules/intersection.py:8-37
// ------------------------------------------------------------------------------------
VO_API void vo_ray_aabb(const float* rays_o, const float* rays_d, float scale, int64_t N,
                        float* hits_t) {
    for (int64_t r = 0; r < N; ++r) {
        float t1 = -INFINITY, t2 = INFINITY;
        for (int d = 0; d < 3; ++d) {
            float o = rays_o[3 * r + d], dir = rays_d[3 * r + d];
            float inv_d = 1.0f / dir;                       // :24
            float half_size = (scale - (-scale)) / 2.0f;    // :17
            float tmin = (0.0f - half_size - o) * inv_d;    // :26
            float tmax = (0.0f + half_size - o) * inv_d;    // :27
            float lo = fminf(tmin, tmax), hi = fmaxf(tmin, tmax);  // :29-30
            t1 = fmaxf(t1, lo);                             // :31
            t2 = fminf(t2, hi);                             // :32
        }
        if (t2 > 0.0f) { hits_t[2 * r] = fmaxf(t1, NEAR_DISTANCE); hits_t[2 * r + 1] = t2; }
        else           { hits_t[2 * r] = -1.0f; hits_t[2 * r + 1] = -1.0f; }
    }
}

// ------------------------------------------------------------------------------------
// helpers: modules/utils.py:54-117
// ------------------------------------------------------------------------------------
static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

static inline float calc_dt(float t, float esf, int grid_size, float scale) {   // utils.py:54-57
    return clampf(t * esf, SQRT3_MAX_SAMPLES, SQRT3_2 * scale / (float)grid_size);
}

static inline int frexp_bit(float x) {                                           // utils.py:60-75
    int exponent = 0;
    if (x != 0.0f) {
        uint32_t bits; std::memcpy(&bits, &x, 4);
        exponent = (int)((bits & 0x7f800000u) >> 23) - 127;
        bits &= 0x7fffffu; bits |= 0x3f800000u;
        float frac; std::memcpy(&frac, &bits, 4);
        if (frac < 0.5f) exponent -= 1;
        else if (frac > 1.0f) exponent += 1;
    }
    return exponent;
}

static inline int mip_from_pos(const float* xyz, int cascades) {                 // utils.py:78-84
    float mx = fmaxf(fmaxf(fabsf(xyz[0]), fabsf(xyz[1])), fabsf(xyz[2]));
    int exponent = frexp_bit(mx) + 1;
    return std::min(cascades - 1, std::max(0, exponent));
}

static inline int mip_from_dt(float dt, int grid_size, int cascades) {           // utils.py:87-92
    int exponent = frexp_bit(dt * (float)grid_size);
    return std::min(cascades - 1, std::max(0, exponent));
}

static inline uint32_t expand_bits(uint32_t v) {                                 // utils.py:95-101
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {            // utils.py:104-107
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert1(uint32_t x) {                            // utils.py:110-117
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

VO_API void vo_morton3d(const int32_t* coords, int64_t n, int32_t* indices) {     // utils.py:137-142
    for (int64_t i = 0; i < n; ++i)
        indices[i] = (int32_t)morton3D((uint32_t)coords[3 * i], (uint32_t)coords[3 * i + 1],
                                       (uint32_t)coords[3 * i + 2]);
}
VO_API void vo_morton3d_invert(const int32_t* indices, int64_t n, int32_t* coords) {  // :120-127
    for (int64_t i = 0; i < n; ++i) {
        uint32_t ind = (uint32_t)indices[i];
        coords[3 * i + 0] = (int32_t)morton3D_invert1(ind >> 0);
        coords[3 * i + 1] = (int32_t)morton3D_invert1(ind >> 1);
        coords[3 * i + 2] = (int32_t)morton3D_invert1(ind >> 2);
    }
}
VO_API void vo_packbits(const float* grid, int64_t n_bytes, float thr, uint8_t* bitfield) {  // :157-169
    for (int64_t n = 0; n < n_bytes; ++n) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; ++i)
            if (grid[8 * n + i] > thr) bits |= (uint8_t)(1u << i);
        bitfield[n] = bits;
    }
}

// ------------------------------------------------------------------------------------
// a6/a7. the shared marching step, modules/ray_march.py:44-75 (train) == :232-267 (test)
// ------------------------------------------------------------------------------------
struct MarchCfg { int cascades, grid_size; float scale, esf; };

// returns occupancy at t; on empty advances t past the cell as the reference does
static inline bool march_probe(const MarchCfg& c, const float* o, const float* d, const float* d_inv,
                               const uint8_t* bitfield, float t, float* xyz, float* dt_out,
                               float* t_next_if_empty) {
    const int G = c.grid_size;
    const uint32_t G3 = (uint32_t)G * (uint32_t)G * (uint32_t)G;
    const float grid_size_inv = 1.0f / (float)G;
    for (int k = 0; k < 3; ++k) xyz[k] = o[k] + t * d[k];                       // :45
    float dt = calc_dt(t, c.esf, G, c.scale);                                   // :46
    int mip = std::max(mip_from_pos(xyz, c.cascades), mip_from_dt(dt, G, c.cascades));  // :47-48
    float mip_bound = fminf(ldexpf(1.0f, mip - 1), c.scale);                    // :50
    float mip_bound_inv = 1.0f / mip_bound;                                     // :51
    float nxyz[3]; uint32_t ci[3];
    for (int k = 0; k < 3; ++k) {
        nxyz[k] = clampf(0.5f * (xyz[k] * mip_bound_inv + 1.0f) * (float)G, 0.0f, (float)G - 1.0f);  // :53-57
        ci[k] = f2u_sat(nxyz[k]);
    }
    uint32_t idx = (uint32_t)mip * G3 + morton3D(ci[0], ci[1], ci[2]);          // :59
    bool occ = (bitfield[idx / 8] & (1u << (idx % 8))) != 0;                     // :60
    *dt_out = dt;
    if (!occ) {
        float tmin = INFINITY;
        for (int k = 0; k < 3; ++k) {
            float sgn = (d[k] > 0.0f) ? 1.0f : ((d[k] < 0.0f) ? -1.0f : 0.0f);
            float tx = (((nxyz[k] + 0.5f + 0.5f * sgn) * grid_size_inv * 2.0f - 1.0f) * mip_bound - xyz[k]) * d_inv[k];  // :67-68
            tmin = fminf(tmin, tx);
        }
        float t_target = t + fmaxf(0.0f, tmin);                                  // :70
        t += calc_dt(t, c.esf, G, c.scale);                                      // :71
        while (t < t_target) t += calc_dt(t, c.esf, G, c.scale);                 // :72-73
        *t_next_if_empty = t;
    }
    return occ;
}

// a6. raymarching_train_kernel pass 1 (ray_march.py:29-75): per-ray sample counts.
// noise is passed in (torch.rand_like on the host side, :139).
VO_API void vo_march_train_count(const float* rays_o, const float* rays_d, const float* hits_t,
                                 const uint8_t* bitfield, const float* noise, int64_t N, int cascades,
                                 int grid_size, float scale, float esf, float max_samples,
                                 int32_t* counts, int threads) {
    MarchCfg c{cascades, grid_size, scale, esf};
    int nt = vo_threads(threads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16) if (nt > 1)
    for (int64_t r = 0; r < N; ++r) {
        const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
        float d_inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};               // :33
        float t1 = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        if (t1 >= 0.0f) { float dt = calc_dt(t1, esf, grid_size, scale); t1 += dt * noise[r]; }  // :39-41
        float t = t1; int n = 0;
        while (0.0f <= t && t < t2 && (float)n < max_samples) {                  // :44
            float xyz[3], dt, tn;
            if (march_probe(c, o, d, d_inv, bitfield, t, xyz, &dt, &tn)) { t += dt; n += 1; }
            else t = tn;
        }
        counts[r] = n;
    }
}

// pass 2 (ray_march.py:84-124) with the canonical ray order: rays_a[r] = (r, start_r, N_r),
// start = exclusive scan of counts in ray order (= the reference run single-threaded).
VO_API int64_t vo_march_train_write(const float* rays_o, const float* rays_d, const float* hits_t,
                                    const uint8_t* bitfield, const float* noise, int64_t N,
                                    int cascades, int grid_size, float scale, float esf,
                                    const int32_t* counts, int32_t* rays_a, float* xyzs, float* dirs,
                                    float* deltas, float* ts, int threads) {
    MarchCfg c{cascades, grid_size, scale, esf};
    int64_t total = 0;
    for (int64_t r = 0; r < N; ++r) {
        rays_a[3 * r] = (int32_t)r; rays_a[3 * r + 1] = (int32_t)total; rays_a[3 * r + 2] = counts[r];
        total += counts[r];
    }
    int nt = vo_threads(threads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16) if (nt > 1)
    for (int64_t r = 0; r < N; ++r) {
        const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
        float d_inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        float t1 = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        if (t1 >= 0.0f) { float dt = calc_dt(t1, esf, grid_size, scale); t1 += dt * noise[r]; }
        float t = t1; int samples = 0; const int n = counts[r];
        const int64_t start = rays_a[3 * r + 1];
        while (t < t2 && samples < n) {                                          // :87
            float xyz[3], dt, tn;
            if (march_probe(c, o, d, d_inv, bitfield, t, xyz, &dt, &tn)) {
                int64_t s = start + samples;
                for (int k = 0; k < 3; ++k) { xyzs[3 * s + k] = xyz[k]; dirs[3 * s + k] = d[k]; }
                ts[s] = t; deltas[s] = dt;                                       // :113-114
                t += dt; samples += 1;
            } else t = tn;
        }
    }
    return total;
}

// a7. raymarching_test_kernel, ray_march.py:198-269.  Writes slot n*max_samples+s, mutates
// hits_t[r,0], and returns per-alive-ray counts; the wrapper's cumsum / mask compaction
// (:328-335) is restated in oracle/__init__.py.
VO_API void vo_march_test(const float* rays_o, const float* rays_d, float* hits_t,
                          const int64_t* alive, int64_t A, const uint8_t* bitfield, int cascades,
                          int grid_size, float scale, float esf, int max_samples,
                          int64_t* ray_indices, uint8_t* valid_mask, float* deltas, float* ts,
                          int32_t* counter) {
    MarchCfg c{cascades, grid_size, scale, esf};
    for (int64_t n = 0; n < A; ++n) {
        int64_t r = alive[n];
        const float* o = rays_o + 3 * r; const float* d = rays_d + 3 * r;
        float d_inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        float t = hits_t[2 * r], t2 = hits_t[2 * r + 1];
        int s = 0; int64_t base = n * (int64_t)max_samples;
        while (0.0f < t && t < t2 && s < max_samples) {                          // :231
            float xyz[3], dt, tn;
            if (march_probe(c, o, d, d_inv, bitfield, t, xyz, &dt, &tn)) {
                int64_t k = base + s;
                ray_indices[k] = r; valid_mask[k] = 1; ts[k] = t; deltas[k] = dt;  // :252-256
                t += dt; hits_t[2 * r] = t; s += 1;                              // :257-259
            } else t = tn;
        }
        counter[n] = s;                                                          // :269
    }
}

// ------------------------------------------------------------------------------------
// a8. volume_rendering_kernel, modules/volume_train.py:22-48.  T is kept in a local (the
// reference's T[] scratch has the same values for the ray's own samples); ws of skipped
// samples is written as 0 (the reference leaves torch.empty garbage there).
// ------------------------------------------------------------------------------------
VO_API void vo_composite_train_fwd(const float* sigmas, const float* rgbs, const float* deltas,
                                   const float* ts, const int32_t* rays_a, int64_t N, float T_thr,
                                   int32_t* total_samples, float* opacity, float* depth, float* rgb,
                                   float* ws, int threads) {
    int nt = vo_threads(threads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16) if (nt > 1)
    for (int64_t n = 0; n < N; ++n) {
        int ray = rays_a[3 * n], start = rays_a[3 * n + 1], ns = rays_a[3 * n + 2];
        float r0 = 0, r1 = 0, r2 = 0, dep = 0, op = 0; int cnt = 0;
        float T = 1.0f;
        for (int k = 0; k < ns; ++k) {
            int64_t s = (int64_t)start + k;
            if (T > T_thr) {                                                     // :36
                float a = 1.0f - exp_cr(-sigmas[s] * deltas[s]);                 // :37
                float w = a * T;                                                 // :38
                r0 += w * rgbs[3 * s]; r1 += w * rgbs[3 * s + 1]; r2 += w * rgbs[3 * s + 2];
                dep += w * ts[s]; op += w; ws[s] = w;
                T = T * (1.0f - a);                                              // :47
                cnt += 1;
            } else ws[s] = 0.0f;
        }
        rgb[3 * ray] = r0; rgb[3 * ray + 1] = r1; rgb[3 * ray + 2] = r2;
        depth[ray] = dep; opacity[ray] = op; total_samples[ray] = cnt;
    }
}

// a9. reverse mode of a8 (volume_train.py:130-175), hand-derived back-to-front recurrence:
//   G_s = dL/drgb . c_s + dL/ddepth t_s + dL/dopacity + dL/dws_s
//   dL/da_s = T_s (G_s - dT_{s+1});  dT_s = G_s a_s + dT_{s+1} (1 - a_s);  dT_{last+1} = 0
//   dsigma_s = dL/da_s * delta_s * exp(-sigma_s delta_s);  dc_s = w_s dL/drgb
// Accumulated in double so the oracle is the accurate side of the rtol-1e-4 comparison.
VO_API void vo_composite_train_bwd(const float* sigmas, const float* rgbs, const float* deltas,
                                   const float* ts, const int32_t* rays_a, int64_t N, float T_thr,
                                   const float* dL_dopacity, const float* dL_ddepth,
                                   const float* dL_drgb, const float* dL_dws, float* dsigmas,
                                   float* drgbs, int threads) {
    int nt = vo_threads(threads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16) if (nt > 1)
    for (int64_t n = 0; n < N; ++n) {
        int ray = rays_a[3 * n], start = rays_a[3 * n + 1], ns = rays_a[3 * n + 2];
        std::vector<float> Ts(ns + 1), as(ns);
        int last = 0; float T = 1.0f;
        for (int k = 0; k < ns; ++k) {
            int64_t s = (int64_t)start + k;
            Ts[k] = T;
            if (T > T_thr) {
                float a = 1.0f - exp_cr(-sigmas[s] * deltas[s]);
                as[k] = a; T = T * (1.0f - a); last = k + 1;
            } else as[k] = 0.0f;
        }
        for (int k = last; k < ns; ++k) {
            int64_t s = (int64_t)start + k;
            dsigmas[s] = 0.0f; drgbs[3 * s] = drgbs[3 * s + 1] = drgbs[3 * s + 2] = 0.0f;
        }
        double dT_next = 0.0;
        const float g0 = dL_drgb[3 * ray], g1 = dL_drgb[3 * ray + 1], g2 = dL_drgb[3 * ray + 2];
        const float gd = dL_ddepth[ray], go = dL_dopacity[ray];
        for (int k = last - 1; k >= 0; --k) {
            int64_t s = (int64_t)start + k;
            double a = as[k], Tk = Ts[k];
            double G = (double)g0 * rgbs[3 * s] + (double)g1 * rgbs[3 * s + 1] + (double)g2 * rgbs[3 * s + 2]
                     + (double)gd * ts[s] + (double)go + (dL_dws ? (double)dL_dws[s] : 0.0);
            double da = Tk * (G - dT_next);
            double w = a * Tk;
            dsigmas[s] = (float)(da * (double)deltas[s] * (1.0 - a));
            drgbs[3 * s] = (float)(w * g0); drgbs[3 * s + 1] = (float)(w * g1); drgbs[3 * s + 2] = (float)(w * g2);
            dT_next = G * a + dT_next * (1.0 - a);
        }
    }
}

// a10. composite_test, modules/volume_render_test.py:4-54 (in place; alive[n] = -1 when done)
VO_API void vo_composite_test(const float* sigmas, const float* rgbs, const float* deltas,
                              const float* ts, const int64_t* pack_info, int64_t* alive, int64_t A,
                              float T_thr, float* opacity, float* depth, float* rgb) {
    for (int64_t n = 0; n < A; ++n) {
        int64_t start = pack_info[2 * n], steps = pack_info[2 * n + 1], ray = alive[n];
        if (steps == 0) { alive[n] = -1; continue; }                             // :23-24
        float T = 1.0f - opacity[ray];                                           // :26
        float c0 = 0, c1 = 0, c2 = 0, dep = 0, op = 0;
        for (int64_t s = 0; s < steps; ++s) {
            int64_t k = start + s;
            float delta = deltas[k];
            float a = 1.0f - exp_cr(-sigmas[k] * delta);                         // :35
            float w = a * T;
            c0 += w * rgbs[3 * k]; c1 += w * rgbs[3 * k + 1]; c2 += w * rgbs[3 * k + 2];
            dep += w * ts[k]; op += w;
            T *= 1.0f - a;                                                       // :44
            if (T <= T_thr) { alive[n] = -1; break; }                            // :46-48
        }
        rgb[3 * ray] += c0; rgb[3 * ray + 1] += c1; rgb[3 * ray + 2] += c2;       // :50-54
        depth[ray] += dep; opacity[ray] += op;
    }
}

// ------------------------------------------------------------------------------------
// a11. dir_encoder, modules/spherical_harmonics.py:16-42 (input used as given)
// ------------------------------------------------------------------------------------
VO_API void vo_sh_encode(const float* dirs, int64_t B, float* emb) {
    for (int64_t i = 0; i < B; ++i) {
        float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
        float* e = emb + 16 * i;
        e[0] = 0.28209479177387814f;
        e[1] = -0.48860251190291987f * y;
        e[2] = 0.48860251190291987f * z;
        e[3] = -0.48860251190291987f * x;
        e[4] = 1.0925484305920792f * xy;
        e[5] = -1.0925484305920792f * yz;
        e[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
        e[7] = -1.0925484305920792f * xz;
        e[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
        e[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
        e[10] = 2.8906114426405538f * xy * z;
        e[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
        e[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
        e[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
        e[14] = 1.4453057213202769f * z * (x2 - y2);
        e[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
    }
}

// ------------------------------------------------------------------------------------
// a12. NGP.density / NGP.forward dense part, modules/networks.py:134-164, 271-282 in fp32
// (what the reference computes on a CUDA-less host).  Weights are torch nn.Linear layout
// [out, in], bias-free.  W1 [64,32], W2 [16,64], W3 [64,32], W4 [64,64], W5 [3,64].
// enc is the hash encoding [S,32]; dirs are the raw ray directions [S,3].
// Outputs sigmas [S], rgbs [S,3]; optionally h [S,16].
// ------------------------------------------------------------------------------------
static inline void matvec(const float* W, const float* x, float* y, int out, int in) {
    for (int o = 0; o < out; ++o) {
        float acc = 0.0f;
        for (int i = 0; i < in; ++i) acc += W[o * in + i] * x[i];
        y[o] = acc;
    }
}

VO_API void vo_mlp_fwd(const float* enc, const float* dirs, int64_t S, const float* W1, const float* W2,
                       const float* W3, const float* W4, const float* W5, float* sigmas, float* rgbs,
                       float* h_out, int threads) {
    int nt = vo_threads(threads);
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
    for (int64_t s = 0; s < S; ++s) {
        float h1[64], h[16], in2[32], h3[64], h4[64], o[3];
        matvec(W1, enc + 32 * s, h1, 64, 32);
        for (int i = 0; i < 64; ++i) h1[i] = fmaxf(h1[i], 0.0f);
        matvec(W2, h1, h, 16, 64);
        sigmas[s] = expf(h[0]);                                                  // networks.py:145 (TruncExp fwd :23)
        if (h_out) for (int i = 0; i < 16; ++i) h_out[16 * s + i] = h[i];
        const float* d = dirs + 3 * s;
        float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);              // :160
        float dn[3] = {(d[0] / nrm + 1.0f) / 2.0f, (d[1] / nrm + 1.0f) / 2.0f, (d[2] / nrm + 1.0f) / 2.0f};  // :161
        vo_sh_encode(dn, 1, in2);
        for (int i = 0; i < 16; ++i) in2[16 + i] = h[i];                         // :162 cat([d, h])
        matvec(W3, in2, h3, 64, 32);
        for (int i = 0; i < 64; ++i) h3[i] = fmaxf(h3[i], 0.0f);
        matvec(W4, h3, h4, 64, 64);
        for (int i = 0; i < 64; ++i) h4[i] = fmaxf(h4[i], 0.0f);
        matvec(W5, h4, o, 3, 64);
        for (int i = 0; i < 3; ++i) rgbs[3 * s + i] = 1.0f / (1.0f + expf(-o[i]));
    }
}

// ------------------------------------------------------------------------------------
// a14. OccupancyGrid pieces, modules/occupancy_grid.py
// ------------------------------------------------------------------------------------
// helpers/geometric_fcts.py:151-171
VO_API void vo_dist_to_cube_border(const float* rays_o, const float* rays_d, int64_t N, float cube_min,
                                   float cube_max, float* dists) {
    for (int64_t n = 0; n < N; ++n) {
        float m = INFINITY;
        for (int k = 0; k < 3; ++k) {
            float o = rays_o[3 * n + k], d = rays_d[3 * n + k], v = INFINITY;
            if (d > 0.0f) v = (cube_max - o) / d;
            if (d < 0.0f) v = (cube_min - o) / d;
            if (v < m) m = v;   // torch.min: first minimum; NaN cannot occur for finite inputs
        }
        dists[n] = m;
    }
}

// torch.linspace(0, 1, steps) in float32 (ATen RangeFactories: symmetric fill)
static inline float linspace01(int i, int steps) {
    float step = (1.0f - 0.0f) / (float)(steps - 1);
    int halfway = steps / 2;
    if (i < halfway) return 0.0f + step * (float)i;
    return 1.0f - step * (float)(steps - i - 1);
}

static inline float round_half_even(float x) { return nearbyintf(x); }

// _calcPos, occupancy_grid.py:293-335 (+ _c2idx :479-480).  noise [N,M,3] in [0,1) or null.
VO_API void vo_occ_calc_pos(const float* rays_o, const float* rays_d, const float* noise, int64_t N,
                            int M, int grid_size, float scale, float noise_every_m, float* cell_dists,
                            float* cell_pos, int32_t* cell_idxs) {
    for (int64_t n = 0; n < N; ++n) {
        const float* o = rays_o + 3 * n;
        float d[3] = {rays_d[3 * n], rays_d[3 * n + 1], rays_d[3 * n + 2]};
        float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);              // :311
        for (int k = 0; k < 3; ++k) d[k] = d[k] / nrm;
        float L; vo_dist_to_cube_border(o, d, 1, -scale, scale, &L);            // :312-317
        for (int m = 0; m < M; ++m) {
            float dist = linspace01(m, M) * L;                                   // :318-319
            cell_dists[n * M + m] = dist;
            for (int k = 0; k < 3; ++k) {
                float p = o[k] + d[k] * dist;                                    // :322
                if (noise) {
                    float nz = 2.0f * noise[(n * M + m) * 3 + k] - 1.0f;         // :326
                    p = p + noise_every_m * dist * nz;                           // :327
                }
                cell_pos[(n * M + m) * 3 + k] = p;
                float mi = (float)(grid_size - 1) * (p + scale) / (2.0f * scale);  // :479
                float r = round_half_even(mi);
                int32_t ii = (int32_t)r;
                ii = std::min(std::max(ii, 0), grid_size - 1);                   // :480
                cell_idxs[(n * M + m) * 3 + k] = ii;
            }
        }
    }
}

static inline float occ_pdf(float meas, float dist, float std_every_m) {         // :464-465
    float stds = std_every_m * dist + 0.00001f;
    float diff = meas - dist;
    return expf((-0.5f * (diff * diff)) / (stds * stds));
}

// _rayProb, occupancy_grid.py:338-389
VO_API void vo_occ_ray_prob(const float* meas, const float* dists, int64_t N, int M, int I,
                            float p_false, float std_every_m, float prob_min, float* probs_occ,
                            float* probs_emp) {
    for (int64_t n = 0; n < N; ++n) {
        float me = meas[n];
        for (int m = 0; m < M; ++m) {
            float dist = dists[n * M + m];
            float eq_emp = p_false;                                              // :361-363
            float eq_occ = eq_emp + occ_pdf(me, dist, std_every_m);              // :364-367
            float nl_emp = 1.0f - eq_emp * dist;                                 // :370
            if (nl_emp < prob_min) nl_emp = prob_min;                            // :371
            float integral = 0.0f;
            for (int k = 0; k < I; ++k) {
                float y = linspace01(k, I) * me;                                 // :374
                integral += occ_pdf(y, dist, std_every_m);                       // :375-378
            }
            integral = integral * (me / (float)I);                               // :379
            float nl_occ = nl_emp - integral;                                    // :380
            if (nl_occ < prob_min) nl_occ = prob_min;                            // :381
            probs_emp[n * M + m] = eq_emp * nl_emp;                              // :384
            probs_occ[n * M + m] = eq_occ * nl_occ;                              // :385
        }
    }
}

// _nerfProb, occupancy_grid.py:392-408 given densities
VO_API void vo_occ_nerf_prob(const float* density, int64_t n, double thr_max, float slope,
                             float* probs_occ, float* probs_emp) {
    // torch.mean in f32: accumulate in double then round (torch CPU uses pairwise/vectorised f32
    // sums; double is the accurate side)
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) acc += density[i];
    float mean = (float)(acc / (double)n);
    double thr = std::min(thr_max, (double)mean);                         // :402
    float h_thr = (float)(-std::log(thr));                                        // :403
    for (int64_t i = 0; i < n; ++i) {
        float h = logf(density[i]);                                               // :404
        float po = 1.0f / (1.0f + expf(-slope * (h - h_thr)));                    // :405
        probs_occ[i] = po; probs_emp[i] = 1.0f - po;                              // :406
    }
}

// _updateGrid, occupancy_grid.py:411-430: gather all, Bayes, scatter; duplicates: the last
// flat index wins (= CPU index_put_ order, the canonical order fixed in DESIGN.md)
VO_API void vo_occ_update_grid(float* grid, int grid_size, const int32_t* cell_idxs, int64_t n,
                               const float* probs_occ, const float* probs_emp) {
    std::vector<float> newp(n);
    const int64_t G = grid_size;
    for (int64_t i = 0; i < n; ++i) {
        int64_t c = ((int64_t)cell_idxs[3 * i] * G + cell_idxs[3 * i + 1]) * G + cell_idxs[3 * i + 2];
        float p = grid[c];
        newp[i] = (p * probs_occ[i]) / (p * probs_occ[i] + (1.0f - p) * probs_emp[i]);  // :429
    }
    for (int64_t i = 0; i < n; ++i) {
        int64_t c = ((int64_t)cell_idxs[3 * i] * G + cell_idxs[3 * i + 1]) * G + cell_idxs[3 * i + 2];
        grid[c] = newp[i];                                                        // :430
    }
}

// update() tail, occupancy_grid.py:96-105: optional decay, then a15 = grid.py:165-170
// (cartesian2morton: morton_grid[morton(x,y,z)] = grid[x,y,z]) + grid.py:205-211 (packbits)
VO_API void vo_occ_decay_pack(float* grid, int grid_size, float decay /*1.0 = none*/, int apply_decay,
                              float thr, uint8_t* bitfield) {
    const int64_t G = grid_size, G3 = G * G * G;
    if (apply_decay) for (int64_t i = 0; i < G3; ++i) grid[i] *= decay;            // :98
    std::vector<float> mg(G3, 0.0f);
    for (int64_t x = 0; x < G; ++x) for (int64_t y = 0; y < G; ++y) for (int64_t z = 0; z < G; ++z)
        mg[morton3D((uint32_t)x, (uint32_t)y, (uint32_t)z)] = grid[(x * G + y) * G + z];
    vo_packbits(mg.data(), G3 / 8, thr, bitfield);
}

// ------------------------------------------------------------------------------------
// caller side (f1): GradScaler unscale + torch.optim.Adam(eps=1e-15) step, training/trainer.py:49-57,
// 138-141 (torch Adam, no weight decay, no amsgrad).  Returns 1 if a non-finite grad was found
// (the step is then skipped, as GradScaler.step does).
// ------------------------------------------------------------------------------------
VO_API int vo_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float inv_scale,
                        double lr, double beta1, double beta2, double eps, int step) {
    for (int64_t i = 0; i < n; ++i) if (!std::isfinite(g[i] * inv_scale)) return 1;
    // torch (_single_tensor_adam): python-double arithmetic on the hyper-parameters, one rounding to f32
    // where the scalar meets the f32 tensor
    double bc1 = 1.0 - std::pow(beta1, step), bc2 = 1.0 - std::pow(beta2, step);
    const float step_size = (float)(lr / bc1);
    const float bc2_sqrt = (float)std::sqrt(bc2);
    const float b2 = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2), epsf = (float)eps;
    for (int64_t i = 0; i < n; ++i) {
        float gi = g[i] * inv_scale;
        m[i] = std::fmaf(gi - m[i], omb1, m[i]);             // torch: exp_avg.lerp_(grad, 1-beta1)
        v[i] = std::fmaf(omb2 * gi, gi, v[i] * b2);          // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
        float denom = sqrtf(v[i]) / bc2_sqrt + epsf;
        p[i] = p[i] - step_size * (m[i] / denom);
    }
    return 0;
}

VO_API int vo_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
