// mlp_common.cuh -- pieces shared by the fused-MLP kernels (mlp_fused.cu: forward + the serial
// backward; mlp_bwd_pipe.cu: the pipelined three-chain backward).
#pragma once
#include "common.cuh"
#include "umma.cuh"
#include "hash_common.cuh"

namespace mlp {

constexpr int TILE = 128;

struct MlpArgs {
    const float* enc;      // [S,32] f32 (or fp16 when enc_half)
    const float* dirs;     // [S,3]
    const float* W[5];     // torch Linear layout [out,in] f32
    float* sigmas;         // [S]
    float* rgbs;           // [S,3]
    float* h_out;          // [S,16] or null: the density net's feature vector (return_feat)
    const float* dsigmas;  // [S]     (BWD)
    const float* drgbs;    // [S,3]   (BWD)
    float* denc;           // [S,32]  (BWD)
    float* dW[5];          // accumulated (BWD)
    int64_t S;
    int enc_half;          // enc_format == 1
    int enc_planar;        // enc_format == 2: enc / denc are [8][S] float4 planes (VN_HASH_PLANAR)
    int density_only;
    int enc_fmt;           // 0 f32 rows [S,32] | 1 f16 rows | 2 f32 planes [8][S] float4 | 3 f16 chunk planes [4][S] x 16 B
    int denc_fmt;          // 0 f32 rows | 2 f32 planes | 3 f16 chunk planes        (BWD)
};

__device__ __forceinline__ void load_weight(uint8_t* smem, int off, const float* __restrict__ W, int R, int C, int R_valid,
                                            int tid) {
    // [R x C] f32 row-major -> chunk-major fp16; one 16-byte chunk (8 columns of one row) per
    // thread and iteration: two LDG.128, one STS.128.  Rows >= R_valid are zero padding.
    const int chunks_per_row = C / 8;
    for (int i = tid; i < R * chunks_per_row; i += blockDim.x) {
        const int r = i / chunks_per_row, c = i % chunks_per_row;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < R_valid) {
            const float4* src = reinterpret_cast<const float4*>(W + (size_t)r * C + 8 * c);
            const float4 lo = __ldg(src), hi = __ldg(src + 1);
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        }
        umma::st_chunk(reinterpret_cast<__half*>(smem + off), R, r, c, v);
    }
}

__device__ __forceinline__ void sh16_half(float x, float y, float z, float* e) {
    // spherical_harmonics.py:16-42
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
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


// the hash encoder's side of the fused backward (vn_mlp_bwd_scatter): d(enc) is scattered into `grad` instead of stored
struct ScatterArgs {
    const float* xyz;      // [S,3] unit-cube positions (the hash forward's input)
    float* grad;           // table gradient [total_entries, 2] f32, accumulated with red.global.add
    int round_f16;         // d(enc) rounded to fp16 before the scatter (half-precision encoder)
    float* found_inf;      // optional: set to 1 when a gradient contribution (table or weights) is not finite
    HashParams P;
};

int launch_mlp_bwd_pipe(const MlpArgs& a, const ScatterArgs* hs, cudaStream_t st);   // mlp_bwd_pipe.cu

}  // namespace mlp
