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

template <typename DT, bool DENSE, bool AGG, bool ZERO_SKIP>
__device__ __forceinline__ void level_scatter(float* __restrict__ grad_level, const Cell& c, uint32_t res,
                                              uint32_t size, uint32_t mask, float d0, float d1, bool valid) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float v0[8], v1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float w = corner_weight(c, k);
        v0[k] = valid ? vn_mul(w, d0) : 0.0f;
        v1[k] = valid ? vn_mul(w, d1) : 0.0f;
    }
    bool leader = valid;
    if (AGG) {
        uint32_t p0 = __shfl_up_sync(full, c.g[0], 1), p1 = __shfl_up_sync(full, c.g[1], 1),
                 p2 = __shfl_up_sync(full, c.g[2], 1);
        int pv = __shfl_up_sync(full, (int)valid, 1);
        bool head = (lane == 0) || !valid || !pv || p0 != c.g[0] || p1 != c.g[1] || p2 != c.g[2];
        unsigned heads = __ballot_sync(full, head);
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
        if (__popc(heads) <= kScanMaxHeads) {  // warp-uniform: enough sharing to pay for the scan
            int seg_start = 31 - __clz(heads & (full >> (31 - lane)));
            // longest run in the warp bounds the number of scan steps that can do anything
            const int max_run = __reduce_max_sync(full, lane - seg_start + 1);
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                if (off >= max_run) break;
                bool take = (lane - off) >= seg_start;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float a = __shfl_up_sync(full, v0[k], off), b = __shfl_up_sync(full, v1[k], off);
                    if (take) { v0[k] += a; v1[k] += b; }
                }
            }
            bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
            leader = valid && tail;
        }
    }
    if (leader) {
        // the x / x+1 corners of a cell are neighbouring table entries whenever their indices
        // differ only in bit 0 (dense levels: x + ... with an even index; hashed levels: the x
        // prime is 1, so an even x gives h and h^1): one 16-byte red.v4 instead of two red.v2
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const uint32_t i0 = corner_index<DENSE>(c, k, res, size, mask);
            const uint32_t i1 = corner_index<DENSE>(c, k + 1, res, size, mask);
            if ((i0 ^ i1) == 1u) {
                const bool lo0 = (i0 & 1u) == 0u;
                vn_red_add_v4(grad_level + 2 * (size_t)(i0 & ~1u), lo0 ? v0[k] : v0[k + 1], lo0 ? v1[k] : v1[k + 1],
                              lo0 ? v0[k + 1] : v0[k], lo0 ? v1[k + 1] : v1[k]);
            } else {
                if (!(ZERO_SKIP && v0[k] == 0.0f && v1[k] == 0.0f))       // hash_encoder_half.py:212
                    vn_red_add_v2(grad_level + 2 * (size_t)i0, v0[k], v1[k]);
                if (!(ZERO_SKIP && v0[k + 1] == 0.0f && v1[k + 1] == 0.0f))
                    vn_red_add_v2(grad_level + 2 * (size_t)i1, v0[k + 1], v1[k + 1]);
            }
        }
    }
}

