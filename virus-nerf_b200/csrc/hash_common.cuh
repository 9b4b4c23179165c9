// hash_common.cuh -- per-(point, level) geometry and the warp-aggregated gradient scatter of the hash-grid encoder,
// shared by hash_encoder.cu (stand-alone forward / backward kernels) and mlp_bwd_pipe.cu (the fused MLP backward that
// scatters d(enc) straight from tensor memory into the table gradient).
// modules/hash_encoder.py:43-60, 89-143 (+ autodiff :264-277), modules/hash_encoder_half.py:164-213.
#pragma once
#include "common.cuh"

struct HashParams {
    int levels;        // launches cover levels [level_begin, levels)
    int level_begin;
    int level_end;
    int begin_fast;
    int offsets[VN_MAX_LEVELS];
    uint32_t sizes[VN_MAX_LEVELS];
    uint32_t pow2mask[VN_MAX_LEVELS];  // size-1 when size is a power of two, else 0
    float scales[VN_MAX_LEVELS];
    uint32_t res[VN_MAX_LEVELS];
};

static inline int make_params(const vn_hash_levels_t* lv, HashParams& P) {
    VN_REQUIRE(lv != nullptr, "hash levels: null");
    VN_REQUIRE(lv->levels >= 1 && lv->levels <= VN_MAX_LEVELS, "hash levels: levels=%d out of [1,%d]",
               lv->levels, VN_MAX_LEVELS);
    P.levels = lv->levels;
    P.level_begin = 0;
    P.level_end = lv->levels;
    P.begin_fast = lv->begin_fast_hash_level;
    for (int l = 0; l < lv->levels; ++l) {
        VN_REQUIRE(lv->sizes[l] > 0, "hash levels: size[%d] <= 0", l);
        P.offsets[l] = lv->offsets[l];
        P.sizes[l] = (uint32_t)lv->sizes[l];
        uint32_t s = (uint32_t)lv->sizes[l];
        P.pow2mask[l] = ((s & (s - 1)) == 0) ? (s - 1) : 0u;
        P.scales[l] = lv->scales[l];
        P.res[l] = lv->res[l];
    }
    return VN_OK;
}

// ---- per-(point, level) geometry shared by fwd / bwd / indices --------------------------
struct Cell {
    uint32_t g[3];
    float f[3];
};

__device__ __forceinline__ Cell cell_of(float x, float y, float z, float scale) {
    Cell c;
    float p0 = vn_add(vn_mul(x, scale), 0.5f);  // hash_encoder.py:106, not contracted
    float p1 = vn_add(vn_mul(y, scale), 0.5f);
    float p2 = vn_add(vn_mul(z, scale), 0.5f);
    float f0 = floorf(p0), f1 = floorf(p1), f2 = floorf(p2);
    c.g[0] = vn_f2u(f0); c.g[1] = vn_f2u(f1); c.g[2] = vn_f2u(f2);   // :107
    c.f[0] = vn_sub(p0, (float)c.g[0]);                                // :108
    c.f[1] = vn_sub(p1, (float)c.g[1]);
    c.f[2] = vn_sub(p2, (float)c.g[2]);
    return c;
}

template <bool DENSE>
__device__ __forceinline__ uint32_t corner_index(const Cell& c, int k, uint32_t res, uint32_t size, uint32_t mask) {
    uint32_t c0 = c.g[0] + (k & 1), c1 = c.g[1] + ((k >> 1) & 1), c2 = c.g[2] + ((k >> 2) & 1);
    uint32_t h;
    if (DENSE) {
        h = c0 + c1 * res + c2 * (res * res);      // under_hash, :53-60 (wrapping u32)
        if (h >= size) h %= size;                  // only reachable on the x/y/z == 1 faces
    } else {
        h = c0 ^ (c1 * 2654435761u) ^ (c2 * 805459861u);  // fast_hash, :43-51
        h = mask ? (h & mask) : (h % size);
    }
    return h;
}

__device__ __forceinline__ float corner_weight(const Cell& c, int k) {
    // w = ((1 * wx) * wy) * wz in source order, :118-125
    float w = (k & 1) ? c.f[0] : vn_sub(1.0f, c.f[0]);
    w = vn_mul(w, (k & 2) ? c.f[1] : vn_sub(1.0f, c.f[1]));
    w = vn_mul(w, (k & 4) ? c.f[2] : vn_sub(1.0f, c.f[2]));
    return w;
}

// ---- backward ---------------------------------------------------------------------------
// Warp pre-reduction: lanes are consecutive samples of a ray, so lanes in the same grid cell
// form contiguous runs.  Head flags come from cell equality with the previous lane; a
// segmented shuffle scan leaves each run's sum in its last lane, which issues the
// red.global.add.v2.f32.  Runs are detected per level; when (nearly) every lane is its own
// run the scan is skipped.
// All 32 lanes contribute to the same 8 corners: reduce 8 float2 values across the warp by
// recursive halving (each step a lane keeps half of its values and adds its partner's copy of
// that half): 4 + 2 + 1 float2 exchanges, then two plain butterfly steps.  18 shuffles instead
// of 80, and the 8 corner sums end up in 8 different lanes (lane & 3 == 0), which then issue
// the 8 atomics in parallel.
__device__ __forceinline__ void warp_reduce8_distribute(float* v0, float* v1, int lane, float* r0, float* r1) {
    const unsigned full = 0xffffffffu;
    float a0[4], a1[4];
    const bool hi16 = lane & 16;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float s0 = hi16 ? v0[j] : v0[j + 4], s1 = hi16 ? v1[j] : v1[j + 4];
        const float k0 = hi16 ? v0[j + 4] : v0[j], k1 = hi16 ? v1[j + 4] : v1[j];
        a0[j] = k0 + __shfl_xor_sync(full, s0, 16);
        a1[j] = k1 + __shfl_xor_sync(full, s1, 16);
    }
    float b0[2], b1[2];
    const bool hi8 = lane & 8;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float s0 = hi8 ? a0[j] : a0[j + 2], s1 = hi8 ? a1[j] : a1[j + 2];
        const float k0 = hi8 ? a0[j + 2] : a0[j], k1 = hi8 ? a1[j + 2] : a1[j];
        b0[j] = k0 + __shfl_xor_sync(full, s0, 8);
        b1[j] = k1 + __shfl_xor_sync(full, s1, 8);
    }
    const bool hi4 = lane & 4;
    float c0 = (hi4 ? b0[1] : b0[0]) + __shfl_xor_sync(full, hi4 ? b0[0] : b0[1], 4);
    float c1 = (hi4 ? b1[1] : b1[0]) + __shfl_xor_sync(full, hi4 ? b1[0] : b1[1], 4);
    c0 += __shfl_xor_sync(full, c0, 2); c1 += __shfl_xor_sync(full, c1, 2);
    c0 += __shfl_xor_sync(full, c0, 1); c1 += __shfl_xor_sync(full, c1, 1);
    *r0 = c0; *r1 = c1;   // corner index held by this lane: ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1)
}

constexpr int kScanMaxHeads = 24;
// (Measured dead end, profiles/r2_kbench.md: NOT looking for runs at the finest levels -- where nearly every lane is its
// own run -- is slower, 0.526 -> 0.531 / 0.535 ms for the fused kernel with the cut at level 14 / 13: the warps of rays
// that run along a grid axis still merge there, and a red lane costs more than the 12 instructions of the test.)

// run detection shared by the scatter variants: bit l of the result = lane l starts a new run (its cell differs from the
// previous lane's, or one of the two lanes carries no gradient).  Two shuffles: (x | y << 16) and (z | invalid << 31).
__device__ __forceinline__ unsigned run_heads(const Cell& c, bool valid, int lane) {
    const unsigned full = 0xffffffffu;
    const uint32_t k1 = c.g[0] | (c.g[1] << 16), k2 = c.g[2] | (valid ? 0u : 0x80000000u);
    const uint32_t p1 = __shfl_up_sync(full, k1, 1), p2 = __shfl_up_sync(full, k2, 1);
    const bool head = (lane == 0) || !valid || p1 != k1 || p2 != k2;
    return __ballot_sync(full, head);
}

// segmented inclusive scan of the 8 float2 corner values over the runs given by `heads`; returns true in the last lane
// of every run (which then holds the run's sums)
__device__ __forceinline__ bool run_scan(float* v0, float* v1, unsigned heads, int lane) {
    const unsigned full = 0xffffffffu;
    const int seg_start = 31 - __clz(heads & (full >> (31 - lane)));
    const int max_run = __reduce_max_sync(full, lane - seg_start + 1);     // bounds the number of useful scan steps
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        if (off >= max_run) break;
        const bool take = (lane - off) >= seg_start;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float a = __shfl_up_sync(full, v0[k], off), b = __shfl_up_sync(full, v1[k], off);
            if (take) { v0[k] += a; v1[k] += b; }
        }
    }
    return (lane == 31) || ((heads >> (lane + 1)) & 1u);
}

// Hashed level whose table size is a power of two (every hashed level of the shipped configurations).
// index(x, y, z) = (x ^ y P1 ^ z P2) & mask with odd primes: for an EVEN cell x the two x-corners of a (y, z) pair are
// the entries i and i ^ 1 -- one aligned 16-byte red.v4 -- and WHICH of them sits at the even address is the parity
// q ^ ky ^ kz with q = (y ^ z) & 1.  Instead of ordering the four values of a pair with selects when the atomic is
// issued, the pair is STORED in address order from the start: slot 2j of pair j = (ky, kz) holds x-offset xs_j (0 or 1),
// slot 2j + 1 the other one; that costs two selects on the x-weights.  All lanes of a run share the cell, hence the
// order, so the run sums need no reordering either.  The same products in the same order as corner_weight().
template <bool AGG>
__device__ __forceinline__ void level_scatter_hashed_pow2(float* __restrict__ grad_level, const Cell& c, uint32_t mask, float d0,
                                                          float d1, bool valid) {
    const int lane = threadIdx.x & 31;
    d0 = valid ? d0 : 0.0f;
    d1 = valid ? d1 : 0.0f;
    const bool pair = (c.g[0] & 1u) == 0u;                      // all four pairs of the cell alike
    const bool q = ((c.g[1] ^ c.g[2]) & 1u) != 0u;
    const bool swap_e = pair && q, swap_o = pair && !q;         // pairs (0,0), (1,1)  /  pairs (1,0), (0,1)
    const float wx0 = vn_sub(1.0f, c.f[0]), wx1 = c.f[0];
    const float e_first = swap_e ? wx1 : wx0, e_second = swap_e ? wx0 : wx1;
    const float o_first = swap_o ? wx1 : wx0, o_second = swap_o ? wx0 : wx1;
    const float wy0 = vn_sub(1.0f, c.f[1]), wy1 = c.f[1], wz0 = vn_sub(1.0f, c.f[2]), wz1 = c.f[2];
    float v0[8], v1[8];
    {
        // pair j = ky + 2 kz: j = 0 (0,0) E, 1 (1,0) O, 2 (0,1) O, 3 (1,1) E
        const float w[8] = {vn_mul(vn_mul(e_first, wy0), wz0), vn_mul(vn_mul(e_second, wy0), wz0),
                            vn_mul(vn_mul(o_first, wy1), wz0), vn_mul(vn_mul(o_second, wy1), wz0),
                            vn_mul(vn_mul(o_first, wy0), wz1), vn_mul(vn_mul(o_second, wy0), wz1),
                            vn_mul(vn_mul(e_first, wy1), wz1), vn_mul(vn_mul(e_second, wy1), wz1)};
#pragma unroll
        for (int k = 0; k < 8; ++k) { v0[k] = vn_mul(w[k], d0); v1[k] = vn_mul(w[k], d1); }
    }
    bool leader = valid;
    if (AGG) {
        const unsigned heads = run_heads(c, valid, lane);
        // (a whole warp inside one cell of a hashed level does not occur with marched samples: no special case)
        if (__popc(heads) <= kScanMaxHeads) leader = run_scan(v0, v1, heads, lane) && valid;
    }
    if (!leader) return;
    const uint32_t y0 = c.g[1] * 2654435761u, y1 = y0 + 2654435761u, z0 = c.g[2] * 805459861u, z1 = z0 + 805459861u;
    const uint32_t A[4] = {y0 ^ z0, y1 ^ z0, y0 ^ z1, y1 ^ z1};
    const uint32_t xe = c.g[0] + (swap_e ? 1u : 0u), xo = c.g[0] + (swap_o ? 1u : 0u), x1 = c.g[0] + 1u;
    if (pair) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t i_first = (((j == 0 || j == 3) ? xe : xo) ^ A[j]) & mask;     // even by construction
            vn_red_add_v4(grad_level + 2 * (size_t)i_first, v0[2 * j], v1[2 * j], v0[2 * j + 1], v1[2 * j + 1]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            vn_red_add_v2(grad_level + 2 * (size_t)((c.g[0] ^ A[j]) & mask), v0[2 * j], v1[2 * j]);
            vn_red_add_v2(grad_level + 2 * (size_t)((x1 ^ A[j]) & mask), v0[2 * j + 1], v1[2 * j + 1]);
        }
    }
}

template <typename DT, bool DENSE, bool AGG, bool ZERO_SKIP>
__device__ __forceinline__ void level_scatter(float* __restrict__ grad_level, const Cell& c, uint32_t res,
                                              uint32_t size, uint32_t mask, float d0, float d1, bool valid) {
    if (!DENSE && mask > 2u) {       // warp-uniform (a property of the level)
        level_scatter_hashed_pow2<AGG>(grad_level, c, mask, d0, d1, valid);
        return;
    }
    const int lane = threadIdx.x & 31;
    d0 = valid ? d0 : 0.0f;
    d1 = valid ? d1 : 0.0f;
    float v0[8], v1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float w = corner_weight(c, k);
        v0[k] = vn_mul(w, d0);
        v1[k] = vn_mul(w, d1);
    }
    bool leader = valid;
    if (AGG) {
        const unsigned heads = run_heads(c, valid, lane);
        if (heads == 1u) {          // the whole warp sits in one cell (coarse levels)
            float r0, r1;
            warp_reduce8_distribute(v0, v1, lane, &r0, &r1);
            if ((lane & 3) == 0 && !(ZERO_SKIP && r0 == 0.0f && r1 == 0.0f)) {
                const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                const uint32_t idx = corner_index<DENSE>(c, k, res, size, mask);
                vn_red_add_v2(grad_level + 2 * (size_t)idx, r0, r1);
            }
            return;
        }
        if (__popc(heads) <= kScanMaxHeads) leader = run_scan(v0, v1, heads, lane) && valid;   // enough sharing to pay for the scan
    }
    if (leader) {
        // the x / x+1 corners of a cell are neighbouring table entries whenever their indices differ only in bit 0: one
        // 16-byte red.v4 instead of two red.v2.  Dense levels: index x + y res + z res^2 -- i1 = i0 + 1, so the pair is
        // adjacent exactly when i0 is even (a wrap-around at the end of the slab never yields i0 ^ i1 == 1) and needs no
        // reordering.  (An exactly-zero corner value is added like any other: x + 0 = x, hash_encoder_half.py:212.)
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const uint32_t i0 = corner_index<DENSE>(c, k, res, size, mask);
            const uint32_t i1 = corner_index<DENSE>(c, k + 1, res, size, mask);
            if ((i0 ^ i1) == 1u) {
                if (DENSE) {
                    vn_red_add_v4(grad_level + 2 * (size_t)i0, v0[k], v1[k], v0[k + 1], v1[k + 1]);
                } else {
                    const bool lo0 = (i0 & 1u) == 0u;
                    vn_red_add_v4(grad_level + 2 * (size_t)(i0 & ~1u), lo0 ? v0[k] : v0[k + 1], lo0 ? v1[k] : v1[k + 1],
                                  lo0 ? v0[k + 1] : v0[k], lo0 ? v1[k + 1] : v1[k]);
                }
            } else {
                vn_red_add_v2(grad_level + 2 * (size_t)i0, v0[k], v1[k]);
                vn_red_add_v2(grad_level + 2 * (size_t)i1, v0[k + 1], v1[k + 1]);
            }
        }
    }
}
