// hash_encoder.cu -- multi-resolution hash-grid encoder, forward gather + backward scatter.
// Replaces the Taichi kernels of modules/hash_encoder.py:89-143 (+ autodiff, :264-277) and
// modules/hash_encoder_half.py:112-213.
//
// Mapping: one thread = one point x LPT consecutive levels; lanes of a warp are CONSECUTIVE
// points, i.e. consecutive samples along a ray, so at coarse levels all lanes hit the same 8
// corners (one L1 wavefront per load) and at fine levels the level slab is L2 resident.
// grid.y walks level groups slowest, so with LPT < levels the launch is level-major and one
// group's slab (<= 32 MB at T=2^22) stays in the 126 MB L2 while all points stream past it.
// Per-level constants are precomputed on the host (vn_hash_levels_init) and arrive in the
// kernel parameter bank.
#include "common.cuh"
#include "hash_common.cuh"
#include <math.h>
#include <string.h>

// a1. host geometry: hash_encoder.py:183-208, utils.py:19-42 (float64) and the kernel-side
// f32 constants of hash_encoder.py:73-80.
VN_API int vn_hash_levels_init(double base_res, double max_res, int levels, int64_t max_params,
                               vn_hash_levels_t* out) {
    VN_REQUIRE(out != nullptr, "vn_hash_levels_init: null output");
    VN_REQUIRE(levels >= 2 && levels <= VN_MAX_LEVELS, "vn_hash_levels_init: levels=%d out of [2,%d]", levels,
               VN_MAX_LEVELS);
    VN_REQUIRE(base_res > 0 && max_res >= base_res && max_params > 0, "vn_hash_levels_init: bad resolution/params");
    memset(out, 0, sizeof(*out));
    double log_b = log(max_res / base_res) / (double)(levels - 1);
    out->levels = levels;
    out->log_b = log_b;
    int64_t offset = 0;
    int begin_fast = levels;
    for (int i = 0; i < levels; ++i) {
        double r = ceil(base_res * exp((double)i * log_b) - 1.0) + 1.0;
        double full = r * r * r;
        int64_t aligned = (int64_t)((full + 7.0) / 8.0) * 8;
        int64_t size_i = aligned < max_params ? aligned : max_params;
        VN_REQUIRE(offset + size_i < (int64_t)1 << 30, "vn_hash_levels_init: table too large for i32 offsets");
        out->offsets[i] = (int32_t)offset;
        out->sizes[i] = (int32_t)size_i;
        if (full > (double)size_i && begin_fast == levels) begin_fast = i;
        offset += size_i;
        float sc = (float)base_res * expf((float)i * (float)log_b) - 1.0f;
        out->scales[i] = sc;
        out->res[i] = (uint32_t)ceilf(sc) + 1u;
    }
    out->begin_fast_hash_level = begin_fast;
    out->total_entries = offset;
    return VN_OK;
}

// ---- table element access ---------------------------------------------------------------
// (Measured dead end, profiles/r2_kbench.md: ld.global.nc.L1::no_allocate for the gathers -- the idea being that random
// 8-byte gathers have no reuse -- is 28 % SLOWER, 0.190 -> 0.243 ms at 1.3 M samples: the coarse levels and the
// neighbouring samples of a ray do reuse lines in L1.)
__device__ __forceinline__ float2 load_entry(const float2* t, uint32_t i) { return __ldg(t + i); }
__device__ __forceinline__ float4 load_pair(const float4* t, uint32_t i) { return __ldg(t + i); }
__device__ __forceinline__ float2 load_entry(const __half2* t, uint32_t i) {
    return __half22float2(__ldg(t + i));
}

// PAIR (fp32 tables): the x / x+1 corners of a cell are neighbouring entries whenever their
// indices differ only in bit 0 (see level_scatter) -- one 16-byte load instead of two 8-byte
// loads, i.e. half the L1 wavefronts for those lanes.  Values and summation order unchanged.
template <typename TT, bool DENSE, bool PAIR>
__device__ __forceinline__ void level_gather(const TT* __restrict__ tbl, const Cell& c, uint32_t res, uint32_t size,
                                             uint32_t mask, float& a0, float& a1) {
    float2 v[8];
    float w[8];
    if (PAIR && sizeof(TT) == sizeof(float2)) {
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const uint32_t i0 = corner_index<DENSE>(c, k, res, size, mask);
            const uint32_t i1 = corner_index<DENSE>(c, k + 1, res, size, mask);
            if ((i0 ^ i1) == 1u) {
                const float4 q = load_pair(reinterpret_cast<const float4*>(tbl), i0 >> 1);
                const bool lo0 = (i0 & 1u) == 0u;
                v[k] = lo0 ? make_float2(q.x, q.y) : make_float2(q.z, q.w);
                v[k + 1] = lo0 ? make_float2(q.z, q.w) : make_float2(q.x, q.y);
            } else {
                v[k] = load_entry(tbl, i0);
                v[k + 1] = load_entry(tbl, i1);
            }
            w[k] = corner_weight(c, k);
            w[k + 1] = corner_weight(c, k + 1);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = load_entry(tbl, corner_index<DENSE>(c, k, res, size, mask));
            w[k] = corner_weight(c, k);
        }
    }
    if (sizeof(TT) == sizeof(float2)) {
        a0 = 0.0f; a1 = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a0 = fmaf(w[k], v[k].x, a0); a1 = fmaf(w[k], v[k].y, a1); }
    } else {
        // hash_encoder_half.py:159: local_features(f16) += f16(w * table)
        __half h0 = __float2half_rn(0.0f), h1 = __float2half_rn(0.0f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            h0 = __hadd(h0, __float2half_rn(vn_mul(w[k], v[k].x)));
            h1 = __hadd(h1, __float2half_rn(vn_mul(w[k], v[k].y)));
        }
        a0 = __half2float(h0); a1 = __half2float(h1);
    }
}

// ---- forward ----------------------------------------------------------------------------
// PLANAR (f32 only, levels and LPT even): out is [levels/2][S] float4 -- plane p holds
// (level 2p f0, f1, level 2p+1 f0, f1) of every point, so a warp's store of one plane is one
// contiguous 512-byte run instead of 32 separate 16-byte pieces of 32 rows.
// CHUNK (LPT % 4 == 0, levels % 4 == 0): out is [levels/4][S] x 16 B -- plane c holds levels 4c .. 4c+3 of
// every point as 8 fp16 values, which is one row of one column chunk of the fused MLP's UMMA operand
// (umma.cuh): a 128-sample tile of a plane is 2 KB of contiguous, ready-to-multiply operand.  The fp32
// sums are rounded to fp16 exactly where torch.autocast rounds the Linear input.
template <typename TT, typename OT, int LPT, bool PLANAR, bool PAIR = false, bool CHUNK = false>
__global__ void __launch_bounds__(256) hash_fwd_kernel(const float* __restrict__ xyz, const TT* __restrict__ table,
                                                       OT* __restrict__ out, int64_t S,
                                                       const __grid_constant__ HashParams P) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const int level0 = blockIdx.y * LPT;
    const float x = __ldcg(xyz + 3 * i), y = __ldcg(xyz + 3 * i + 1), z = __ldcg(xyz + 3 * i + 2);
    float acc[2 * LPT];
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
        const int level = level0 + l;
        if (level < P.levels) {
            const Cell c = cell_of(x, y, z, P.scales[level]);
            const TT* tbl = table + P.offsets[level];
            if (level < P.begin_fast)
                level_gather<TT, true, PAIR>(tbl, c, P.res[level], P.sizes[level], 0u, acc[2 * l], acc[2 * l + 1]);
            else
                level_gather<TT, false, PAIR>(tbl, c, P.res[level], P.sizes[level], P.pow2mask[level], acc[2 * l],
                                        acc[2 * l + 1]);
        } else {
            acc[2 * l] = 0.0f; acc[2 * l + 1] = 0.0f;
        }
    }
    const int W = 2 * P.levels;
    if (CHUNK) {
        uint4* o = (uint4*)out + (int64_t)(level0 / 4) * S + i;
#pragma unroll
        for (int q = 0; q < LPT / 4; ++q) {
            if (level0 + 4 * q < P.levels) {
                __half2 h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(acc[8 * q + 2 * j], acc[8 * q + 2 * j + 1]);
                o[(int64_t)q * S] = *reinterpret_cast<const uint4*>(h);
            }
        }
    } else if (PLANAR) {
        float4* o = (float4*)out + (int64_t)(level0 / 2) * S + i;
#pragma unroll
        for (int q = 0; q < LPT / 2; ++q)
            if (level0 + 2 * q < P.levels) o[(int64_t)q * S] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    } else if (sizeof(OT) == 4) {
        float* o = (float*)out + i * W + 2 * level0;
        if (LPT % 2 == 0 && (W % 4) == 0 && level0 + LPT <= P.levels) {
#pragma unroll
            for (int q = 0; q < LPT / 2; ++q)
                ((float4*)o)[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        } else {
#pragma unroll
            for (int l = 0; l < LPT; ++l)
                if (level0 + l < P.levels) ((float2*)o)[l] = make_float2(acc[2 * l], acc[2 * l + 1]);
        }
    } else {
        __half2* o = (__half2*)out + i * P.levels + level0;
#pragma unroll
        for (int l = 0; l < LPT; ++l)
            if (level0 + l < P.levels) o[l] = __floats2half2_rn(acc[2 * l], acc[2 * l + 1]);
    }
}

// ---- backward: warp pre-reduction + level_scatter live in hash_common.cuh ------------------
// CHUNK (DT = __half): dout is the fused MLP's fp16 chunk-plane gradient [levels/4][S] x 16 B (VN_HASH_F16_CHUNKS)
template <typename DT, int LPT, bool AGG, bool ZERO_SKIP, bool PLANAR, int MINB = 1, bool CHUNK = false>
__global__ void __launch_bounds__(256, MINB) hash_bwd_kernel(const float* __restrict__ xyz, const DT* __restrict__ dout,
                                                       float* __restrict__ grad, int64_t S,
                                                       const __grid_constant__ HashParams P) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: see common.cuh
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < S;
    const int level0 = P.level_begin + blockIdx.y * LPT;
    float x = 0.f, y = 0.f, z = 0.f;
    float d[2 * LPT];
#pragma unroll
    for (int l = 0; l < 2 * LPT; ++l) d[l] = 0.0f;
    if (valid) {
        x = __ldcg(xyz + 3 * i); y = __ldcg(xyz + 3 * i + 1); z = __ldcg(xyz + 3 * i + 2);
        const int W = 2 * P.levels;
        if (PLANAR) {
            // level0 is even (level_begin and LPT even): plane level0/2 + q at index i
            const float4* dp = (const float4*)dout + (int64_t)(level0 / 2) * S + i;
#pragma unroll
            for (int q = 0; q < LPT / 2; ++q)
                if (level0 + 2 * q < P.level_end) {
                    float4 t = __ldcg(dp + (int64_t)q * S);
                    d[4 * q] = t.x; d[4 * q + 1] = t.y; d[4 * q + 2] = t.z; d[4 * q + 3] = t.w;
                }
        } else if (sizeof(DT) == 4) {
            const float* dp = (const float*)dout + i * W + 2 * level0;
            if (LPT % 2 == 0 && (W % 4) == 0 && level0 + LPT <= P.level_end) {
#pragma unroll
                for (int q = 0; q < LPT / 2; ++q) {
                    float4 t = __ldcg((const float4*)dp + q);
                    d[4 * q] = t.x; d[4 * q + 1] = t.y; d[4 * q + 2] = t.z; d[4 * q + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int l = 0; l < LPT; ++l)
                    if (level0 + l < P.level_end) { float2 t = __ldcg((const float2*)dp + l); d[2 * l] = t.x; d[2 * l + 1] = t.y; }
            }
        } else if (CHUNK) {
            // level l of point i: half2 number (l & 3) of the 16-byte element i of plane l >> 2
            if (LPT == 4 && (level0 & 3) == 0 && level0 + 4 <= P.level_end) {
                const uint4 u = __ldcg((const uint4*)dout + (int64_t)(level0 >> 2) * S + i);
                const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                for (int l = 0; l < 4; ++l) { const float2 t = __half22float2(h[l]); d[2 * l] = t.x; d[2 * l + 1] = t.y; }
            } else if (LPT == 2 && (level0 & 1) == 0 && level0 + 2 <= P.level_end) {
                const uint2 u = __ldcg((const uint2*)dout + ((int64_t)(level0 >> 2) * S + i) * 2 + ((level0 >> 1) & 1));
                const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                for (int l = 0; l < 2; ++l) { const float2 t = __half22float2(h[l]); d[2 * l] = t.x; d[2 * l + 1] = t.y; }
            } else {
#pragma unroll
                for (int l = 0; l < LPT; ++l) {
                    const int level = level0 + l;
                    if (level < P.level_end) {
                        const float2 t = __half22float2(__ldcg((const __half2*)dout + ((int64_t)(level >> 2) * S + i) * 4 + (level & 3)));
                        d[2 * l] = t.x; d[2 * l + 1] = t.y;
                    }
                }
            }
        } else {
            const __half2* dp = (const __half2*)dout + i * P.levels + level0;
#pragma unroll
            for (int l = 0; l < LPT; ++l)
                if (level0 + l < P.level_end) { float2 t = __half22float2(__ldcg(dp + l)); d[2 * l] = t.x; d[2 * l + 1] = t.y; }
        }
    }
#pragma unroll
    for (int l = 0; l < LPT; ++l) {
        const int level = level0 + l;
        if (level >= P.level_end) break;   // warp-uniform
        bool v = valid;
        if (ZERO_SKIP) v = v && !(d[2 * l] == 0.0f && d[2 * l + 1] == 0.0f);  // hash_encoder_half.py:210
        const Cell c = cell_of(x, y, z, P.scales[level]);
        float* gl = grad + 2 * (size_t)P.offsets[level];
        if (level < P.begin_fast)
            level_scatter<DT, true, AGG, ZERO_SKIP>(gl, c, P.res[level], P.sizes[level], 0u, d[2 * l], d[2 * l + 1], v);
        else
            level_scatter<DT, false, AGG, ZERO_SKIP>(gl, c, P.res[level], P.sizes[level], P.pow2mask[level], d[2 * l],
                                                     d[2 * l + 1], v);
    }
}

// ---- KAT kernel -------------------------------------------------------------------------
__global__ void hash_indices_kernel(const float* __restrict__ xyz, int64_t S, int32_t* __restrict__ idx,
                                    float* __restrict__ w, const __grid_constant__ HashParams P) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * P.levels) return;
    const int64_t i = t / P.levels;
    const int level = (int)(t % P.levels);
    const Cell c = cell_of(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], P.scales[level]);
    for (int k = 0; k < 8; ++k) {
        uint32_t h = (level < P.begin_fast) ? corner_index<true>(c, k, P.res[level], P.sizes[level], 0u)
                                            : corner_index<false>(c, k, P.res[level], P.sizes[level], P.pow2mask[level]);
        idx[t * 8 + k] = (int32_t)h;
        if (w) w[t * 8 + k] = corner_weight(c, k);
    }
}

__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = __ldg((const float4*)(src + i));
        ((__half2*)(dst + i))[0] = __floats2half2_rn(v.x, v.y);
        ((__half2*)(dst + i))[1] = __floats2half2_rn(v.z, v.w);
    } else {
        for (; i < n; ++i) dst[i] = __float2half_rn(src[i]);
    }
}

// ---- launchers --------------------------------------------------------------------------
// levels per thread.  Measured on B200 (tools/kbench.py, profiles/r1_kbench.md): 4 levels per
// thread is the best forward at both table sizes and the best backward while the table is L2
// resident; the backward prefers 2 when it is not (T = 2^22).  8 / 16 levels per thread touch
// every output row only once (less DRAM traffic) but lose more to occupancy than they gain.
static int pick_lpt(int flags, const vn_hash_levels_t* lv, size_t entry_bytes, bool backward) {
    if (!backward && (flags & VN_HASH_F16_CHUNKS)) return (flags & VN_HASH_LEVEL_GROUPS_8) ? 8 : ((flags & VN_HASH_LEVEL_GROUPS_16) ? 16 : 4);
    if (flags & VN_HASH_LEVEL_GROUPS_1) return 1;
    if (flags & VN_HASH_LEVEL_GROUPS_2) return 2;
    if (flags & VN_HASH_LEVEL_GROUPS_4) return 4;
    if (flags & VN_HASH_LEVEL_GROUPS_8) return 8;
    if (flags & VN_HASH_LEVEL_GROUPS_16) return 16;
    const size_t table_bytes = (size_t)lv->total_entries * entry_bytes;
    if (backward && table_bytes > ((size_t)96 << 20)) return 2;
    return 4;
}

template <typename TT, typename OT>
static int launch_fwd(const float* xyz, const TT* table, OT* out, int64_t S, const vn_hash_levels_t* lv, int flags,
                      cudaStream_t st) {
    HashParams P;
    int rc = make_params(lv, P);
    if (rc) return rc;
    if (S == 0) return VN_OK;
    VnProfScope prof(VN_K_HASH_FWD, S, st);
    const int lpt = pick_lpt(flags, lv, sizeof(TT), false);
    dim3 block(256), grid(vn_blocks(S, 256), (P.levels + lpt - 1) / lpt);
    if (flags & VN_HASH_F16_CHUNKS) {
        VN_REQUIRE(P.levels % 4 == 0 && lpt % 4 == 0, "hash fwd: the f16 chunk layout needs levels %% 4 == 0 and >= 4 levels per thread");
        VN_REQUIRE(vn_aligned(out, 16), "hash fwd: the f16 chunk layout needs a 16-byte aligned output");
        const bool pair = (flags & VN_HASH_PAIR_LOADS) != 0 && sizeof(TT) == 8;
        switch (lpt) {
            case 4: if (pair) vn_launch_pdl(hash_fwd_kernel<TT, OT, 4, false, true, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P);
                    else vn_launch_pdl(hash_fwd_kernel<TT, OT, 4, false, false, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
            case 8: if (pair) vn_launch_pdl(hash_fwd_kernel<TT, OT, 8, false, true, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P);
                    else vn_launch_pdl(hash_fwd_kernel<TT, OT, 8, false, false, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
            default: vn_launch_pdl(hash_fwd_kernel<TT, OT, 16, false, false, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
        }
        VN_CHECK_LAUNCH("hash_fwd_kernel<f16 chunks>");
        return VN_OK;
    }
    if (flags & VN_HASH_PLANAR) {
        VN_REQUIRE(sizeof(OT) == 4 && P.levels % 2 == 0 && lpt % 2 == 0,
                   "hash fwd: the planar layout needs f32 output, an even level count and >= 2 levels per thread");
        if constexpr (sizeof(OT) == 4) {
            const bool pair = (flags & VN_HASH_PAIR_LOADS) != 0;   // the level slabs are 16-byte aligned (sizes % 8 == 0)
            switch (lpt) {
                case 2: if (pair) vn_launch_pdl(hash_fwd_kernel<TT, OT, 2, true, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P);
                        else vn_launch_pdl(hash_fwd_kernel<TT, OT, 2, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
                case 4: if (pair) vn_launch_pdl(hash_fwd_kernel<TT, OT, 4, true, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P);
                        else vn_launch_pdl(hash_fwd_kernel<TT, OT, 4, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
                case 8: vn_launch_pdl(hash_fwd_kernel<TT, OT, 8, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
                default: vn_launch_pdl(hash_fwd_kernel<TT, OT, 16, true>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
            }
        }
        VN_CHECK_LAUNCH("hash_fwd_kernel<planar>");
        return VN_OK;
    }
    switch (lpt) {
        case 1: vn_launch_pdl(hash_fwd_kernel<TT, OT, 1, false>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
        case 2: vn_launch_pdl(hash_fwd_kernel<TT, OT, 2, false>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
        case 4: vn_launch_pdl(hash_fwd_kernel<TT, OT, 4, false>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
        case 8: vn_launch_pdl(hash_fwd_kernel<TT, OT, 8, false>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
        default: vn_launch_pdl(hash_fwd_kernel<TT, OT, 16, false>, dim3(grid), dim3(block), 0, st, xyz, table, out, S, P); break;
    }
    VN_CHECK_LAUNCH("hash_fwd_kernel");
    return VN_OK;
}

template <typename DT, bool ZERO_SKIP>
static int launch_bwd(const float* xyz, const DT* dout, float* grad, int64_t S, const vn_hash_levels_t* lv, int flags,
                      cudaStream_t st, int level_begin = 0, int level_end = -1) {
    HashParams P;
    int rc = make_params(lv, P);
    if (rc) return rc;
    if (level_end < 0) level_end = P.levels;
    VN_REQUIRE(level_begin >= 0 && level_begin <= level_end && level_end <= P.levels, "hash bwd: bad level range [%d,%d)",
               level_begin, level_end);
    P.level_begin = level_begin;
    P.level_end = level_end;
    if (S == 0 || level_begin == level_end) return VN_OK;
    VnProfScope prof(VN_K_HASH_BWD, S, st);
    const int lpt = pick_lpt(flags, lv, 8, true);
    const bool agg = !(flags & VN_HASH_NO_WARP_AGG);
    dim3 block(256), grid(vn_blocks(S, 256), (level_end - level_begin + lpt - 1) / lpt);
    if (sizeof(DT) == 2 && (flags & VN_HASH_F16_CHUNKS)) {
        VN_REQUIRE(P.levels % 4 == 0 && agg && vn_aligned(dout, 16), "hash bwd: the f16 chunk layout needs levels %% 4 == 0, "
                   "warp aggregation and a 16-byte aligned gradient");
        if constexpr (sizeof(DT) == 2) {
            switch (lpt) {
                case 2: vn_launch_pdl(hash_bwd_kernel<DT, 2, true, ZERO_SKIP, false, 5, true>, dim3(grid), dim3(block), 0, st, xyz, dout, grad, S, P); break;
                case 4: vn_launch_pdl(hash_bwd_kernel<DT, 4, true, ZERO_SKIP, false, 5, true>, dim3(grid), dim3(block), 0, st, xyz, dout, grad, S, P); break;
                default: VN_REQUIRE(false, "hash bwd: the f16 chunk layout takes 2 or 4 levels per thread");
            }
        }
        VN_CHECK_LAUNCH("hash_bwd_kernel<f16 chunks>");
        return VN_OK;
    }
    if (flags & VN_HASH_PLANAR) {
        VN_REQUIRE(sizeof(DT) == 4 && P.levels % 2 == 0 && lpt % 2 == 0 && level_begin % 2 == 0 && level_end % 2 == 0 && agg,
                   "hash bwd: the planar layout needs f32 gradients, even level counts / ranges and >= 2 levels per thread");
        if constexpr (sizeof(DT) == 4) {
            // VN_HASH_TIGHT_REGS: 48 registers (5 CTAs per SM instead of 4) at the price of a few spilled bytes
            const bool tight = (flags & VN_HASH_TIGHT_REGS) != 0;
            switch (lpt) {
                case 2: if (tight) vn_launch_pdl(hash_bwd_kernel<DT, 2, true, ZERO_SKIP, true, 5>, dim3(grid), dim3(block), 0, st, xyz, dout, grad, S, P);
                        else vn_launch_pdl(hash_bwd_kernel<DT, 2, true, ZERO_SKIP, true>, dim3(grid), dim3(block), 0, st, xyz, dout, grad, S, P); break;
                case 4: if (tight) vn_launch_pdl(hash_bwd_kernel<DT, 4, true, ZERO_SKIP, true, 5>, dim3(grid), dim3(block), 0, st, xyz, dout, grad, S, P);
                        else vn_launch_pdl(hash_bwd_kernel<DT, 4, true, ZERO_SKIP, true>, dim3(grid), dim3(block), 0, st, xyz, dout, grad, S, P); break;
                case 8: vn_launch_pdl(hash_bwd_kernel<DT, 8, true, ZERO_SKIP, true>, dim3(grid), dim3(block), 0, st, xyz, dout, grad, S, P); break;
                default: vn_launch_pdl(hash_bwd_kernel<DT, 16, true, ZERO_SKIP, true>, dim3(grid), dim3(block), 0, st, xyz, dout, grad, S, P); break;
            }
        }
        VN_CHECK_LAUNCH("hash_bwd_kernel<planar>");
        return VN_OK;
    }
#define VN_BWD(L, A) vn_launch_pdl(hash_bwd_kernel<DT, L, A, ZERO_SKIP, false>, dim3(grid), dim3(block), 0, st, xyz, dout, grad, S, P)
    if (agg) {
        switch (lpt) { case 1: VN_BWD(1, true); break; case 2: VN_BWD(2, true); break; case 4: VN_BWD(4, true); break;
                       case 8: VN_BWD(8, true); break; default: VN_BWD(16, true); }
    } else {
        switch (lpt) { case 1: VN_BWD(1, false); break; case 2: VN_BWD(2, false); break; case 4: VN_BWD(4, false); break;
                       case 8: VN_BWD(8, false); break; default: VN_BWD(16, false); }
    }
#undef VN_BWD
    VN_CHECK_LAUNCH("hash_bwd_kernel");
    return VN_OK;
}

#define VN_HASH_ARGCHECK(name, a, b, c)                                                           \
    VN_REQUIRE(S >= 0, name ": S < 0");                                                           \
    VN_REQUIRE(S == 0 || ((a) && (b) && (c)), name ": null pointer");                             \
    VN_REQUIRE(vn_aligned(a, 4) && vn_aligned(b, 8) && vn_aligned(c, 16), name ": misaligned buffer")

VN_API int vn_hash_encode_fwd_f32(const float* xyz, const float* table, float* out, int64_t S,
                                  const vn_hash_levels_t* lv, int flags, void* stream) {
    VN_HASH_ARGCHECK("vn_hash_encode_fwd_f32", xyz, table, out);
    return launch_fwd<float2, float>(xyz, (const float2*)table, out, S, lv, flags, (cudaStream_t)stream);
}

VN_API int vn_hash_encode_bwd_f32(const float* xyz, const float* dout, float* grad, int64_t S,
                                  const vn_hash_levels_t* lv, int flags, void* stream) {
    VN_HASH_ARGCHECK("vn_hash_encode_bwd_f32", xyz, grad, dout);
    // VN_HASH_SKIP_ZERO_GRADS: samples behind an opaque surface have an exactly zero gradient (the compositor
    // stops at T <= 1e-4); not scattering their zeros leaves every sum unchanged
    if (flags & VN_HASH_SKIP_ZERO_GRADS) return launch_bwd<float, true>(xyz, dout, grad, S, lv, flags, (cudaStream_t)stream);
    return launch_bwd<float, false>(xyz, dout, grad, S, lv, flags, (cudaStream_t)stream);
}

VN_API int vn_hash_encode_bwd_f32_levels(const float* xyz, const float* dout, float* grad, int64_t S,
                                         const vn_hash_levels_t* lv, int flags, int level_begin, int level_end,
                                         void* stream) {
    VN_HASH_ARGCHECK("vn_hash_encode_bwd_f32_levels", xyz, grad, dout);
    return launch_bwd<float, false>(xyz, dout, grad, S, lv, flags, (cudaStream_t)stream, level_begin, level_end);
}

VN_API int vn_hash_encode_fwd_f16(const float* xyz, const void* table_h, void* out_h, int64_t S,
                                  const vn_hash_levels_t* lv, int flags, void* stream) {
    VN_REQUIRE(S >= 0, "vn_hash_encode_fwd_f16: S < 0");
    VN_REQUIRE(S == 0 || (xyz && table_h && out_h), "vn_hash_encode_fwd_f16: null pointer");
    VN_REQUIRE(vn_aligned(table_h, 4) && vn_aligned(out_h, 4), "vn_hash_encode_fwd_f16: misaligned buffer");
    return launch_fwd<__half2, __half>(xyz, (const __half2*)table_h, (__half*)out_h, S, lv, flags, (cudaStream_t)stream);
}

VN_API int vn_hash_encode_bwd_f16(const float* xyz, const void* dout_h, float* grad, int64_t S,
                                  const vn_hash_levels_t* lv, int flags, void* stream) {
    VN_REQUIRE(S >= 0, "vn_hash_encode_bwd_f16: S < 0");
    VN_REQUIRE(S == 0 || (xyz && dout_h && grad), "vn_hash_encode_bwd_f16: null pointer");
    VN_REQUIRE(vn_aligned(dout_h, 4) && vn_aligned(grad, 8), "vn_hash_encode_bwd_f16: misaligned buffer");
    return launch_bwd<__half, true>(xyz, (const __half*)dout_h, grad, S, lv, flags, (cudaStream_t)stream);
}

VN_API int vn_f32_to_f16(const float* src, void* dst_h, int64_t n, void* stream) {
    VN_REQUIRE(n >= 0 && (n == 0 || (src && dst_h)), "vn_f32_to_f16: bad arguments");
    VN_REQUIRE(vn_aligned(src, 16) && vn_aligned(dst_h, 8), "vn_f32_to_f16: misaligned buffer");
    if (n == 0) return VN_OK;
    f32_to_f16_kernel<<<vn_blocks((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(src, (__half*)dst_h, n);
    VN_CHECK_LAUNCH("f32_to_f16_kernel");
    return VN_OK;
}

VN_API int vn_hash_indices(const float* xyz, int64_t S, const vn_hash_levels_t* lv, int32_t* idx, float* w,
                           void* stream) {
    VN_REQUIRE(S >= 0 && (S == 0 || (xyz && idx)), "vn_hash_indices: bad arguments");
    HashParams P;
    int rc = make_params(lv, P);
    if (rc) return rc;
    if (S == 0) return VN_OK;
    hash_indices_kernel<<<vn_blocks(S * P.levels, 256), 256, 0, (cudaStream_t)stream>>>(xyz, S, idx, w, P);
    VN_CHECK_LAUNCH("hash_indices_kernel");
    return VN_OK;
}
