// mlp_fused.cu -- fully fused density / colour MLPs of the NGP model on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM).  Replaces the cuBLAS GEMMs + element-wise
// launches of modules/networks.py:134-164, 195-282 and modules/spherical_harmonics.py.
#include "common.cuh"
#include "umma.cuh"
#include "mlp_common.cuh"
#include <stdlib.h>

// ---- UMMA self-test: one 128 x N x K product in each operand mode the fused kernels use ----
// mode 0 (forward):  D[128,N] = A[128,K] * B[N,K]^T          A, B K-major
// mode 1 (dgrad):    D[128,N] = A[128,K] * B[K,N]            A K-major, B stored [K x N] -> MN-major
// mode 2 (wgrad):    D[128,N] = A[K,128]^T * B[K,N]          A stored [K x 128], B stored [K x N], both MN-major
__global__ void __launch_bounds__(128) umma_selftest_kernel(int mode, int N, int K, const __half* __restrict__ A,
                                                            const __half* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    // stored shapes [rows x cols]
    const int Ar = (mode == 2) ? K : 128, Ac = (mode == 2) ? 128 : K;
    const int Br = (mode == 0) ? N : K, Bc = (mode == 0) ? K : N;
    __half* sA = reinterpret_cast<__half*>(smem);
    __half* sB = reinterpret_cast<__half*>(smem + (size_t)Ar * Ac * 2);
    for (int i = tid; i < Ar * Ac; i += 128) {
        const int r = i / Ac, c = i % Ac;
        *reinterpret_cast<__half*>(reinterpret_cast<char*>(sA) + (size_t)(c / 8) * Ar * 16 + r * 16 + (c % 8) * 2) = A[i];
    }
    for (int i = tid; i < Br * Bc; i += 128) {
        const int r = i / Bc, c = i % Bc;
        *reinterpret_cast<__half*>(reinterpret_cast<char*>(sB) + (size_t)(c / 8) * Br * 16 + r * 16 + (c % 8) * 2) = B[i];
    }
    if (warp == 0) umma::tmem_alloc(&tmem_base, 64);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = umma::instr_desc_f16(128, N, mode == 2 ? 1 : 0, mode == 0 ? 0 : 1);
        for (int k = 0; k < K / 16; ++k) {
            const uint64_t ad = (mode == 2) ? umma::desc_mnmajor(umma::smem_u32(sA), Ar, k) : umma::desc_kmajor(umma::smem_u32(sA), Ar, k);
            const uint64_t bd = (mode == 0) ? umma::desc_kmajor(umma::smem_u32(sB), Br, k) : umma::desc_mnmajor(umma::smem_u32(sB), Br, k);
            umma::mma_f16(tm, ad, bd, idesc, k > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        umma::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        umma::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tm, 64);
}

VN_API int vn_umma_selftest(int mode, int N, int K, const void* A, const void* B, float* D, void* stream) {
    VN_REQUIRE(mode >= 0 && mode <= 2 && N >= 16 && N <= 64 && N % 16 == 0 && K >= 16 && K <= 128 && K % 16 == 0,
               "vn_umma_selftest: unsupported shape");
    VN_REQUIRE(A && B && D, "vn_umma_selftest: null pointer");
    const size_t smem = (size_t)128 * K * 2 + (size_t)((mode == 0) ? N * K : K * N) * 2;
    VN_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, N, K, (const __half*)A, (const __half*)B, D);
    VN_CHECK_LAUNCH("umma_selftest_kernel");
    return VN_OK;
}

// =========================================================================================
// Fused NGP MLPs.  One CTA = 128 threads = one tile of 128 samples; thread t owns sample row
// t = TMEM lane t.  Per layer: the activation tile (fp16, chunk-major, see umma.cuh) is the A
// operand, the weight matrix (fp16, chunk-major, resident in shared memory for the whole
// kernel) the B operand, the fp32 accumulator lives in TMEM; the epilogue (ReLU / exp /
// sigmoid / SH concat) reads its own row with tcgen05.ld and writes the next layer's operand.
//
//   enc[32] -W1-> relu[64] -W2-> h[16]; sigma = exp(h0)          networks.py:142-145
//   [SH16((d/|d|+1)/2) | h16] -W3-> relu[64] -W4-> relu[64] -W5-> sigmoid[3]   networks.py:160-162
//
// Backward (same kernel, BWD = true) recomputes the forward per tile, then chains the five
// dgrad products (weights read as MN-major operands) and accumulates the five weight
// gradients in TMEM across all tiles of the CTA (activations / gradients read as MN-major
// operands, M = 64), flushing them with one atomicAdd pass at the end.
// =========================================================================================
namespace {
using namespace mlp;
// shared-memory operand buffers (bytes)
constexpr int SZ_X0 = TILE * 32 * 2, SZ_H = TILE * 64 * 2, SZ_16 = TILE * 16 * 2;
constexpr int OFF_X0 = 0;
constexpr int OFF_H1 = OFF_X0 + SZ_X0;
constexpr int OFF_IN2 = OFF_H1 + SZ_H;       // 32 cols: [SH16 | h16]
constexpr int OFF_H3 = OFF_IN2 + SZ_X0;
constexpr int OFF_H4 = OFF_H3 + SZ_H;
constexpr int OFF_W1 = OFF_H4 + SZ_H;        // [64 x 32]
constexpr int OFF_W2 = OFF_W1 + 64 * 32 * 2; // [16 x 64]
constexpr int OFF_W3 = OFF_W2 + 16 * 64 * 2; // [64 x 32]
constexpr int OFF_W4 = OFF_W3 + 64 * 32 * 2; // [64 x 64]
constexpr int OFF_W5 = OFF_W4 + 64 * 64 * 2; // [16 x 64], rows 3..15 zero
constexpr int SMEM_FULL = OFF_W5 + 16 * 64 * 2;
constexpr int OFF_D5 = SMEM_FULL;            // [128 x 16] d(out5)
constexpr int OFF_DH = OFF_D5 + SZ_16;       // [128 x 16] d(h)
constexpr int OFF_G = OFF_DH + SZ_16;        // [128 x 64] dH4 / dH3 / dH1 (reused)
constexpr int SMEM_BWD = OFF_G + SZ_H;
// forward-only kernel: H1 / H3 / H4 are never live together -> one buffer, 52 KB, 4 CTAs per SM
constexpr int FOFF_X0 = 0, FOFF_H = FOFF_X0 + SZ_X0, FOFF_IN2 = FOFF_H + SZ_H, FOFF_W1 = FOFF_IN2 + SZ_X0;
constexpr int FOFF_W2 = FOFF_W1 + 64 * 32 * 2, FOFF_W3 = FOFF_W2 + 16 * 64 * 2, FOFF_W4 = FOFF_W3 + 64 * 32 * 2;
constexpr int FOFF_W5 = FOFF_W4 + 64 * 64 * 2;
constexpr int SMEM_FWD = FOFF_W5 + 16 * 64 * 2;

template <bool BWD> struct Lay {
    static constexpr int X0 = BWD ? OFF_X0 : FOFF_X0, H1 = BWD ? OFF_H1 : FOFF_H, IN2 = BWD ? OFF_IN2 : FOFF_IN2;
    static constexpr int H3 = BWD ? OFF_H3 : FOFF_H, H4 = BWD ? OFF_H4 : FOFF_H;
    static constexpr int W1 = BWD ? OFF_W1 : FOFF_W1, W2 = BWD ? OFF_W2 : FOFF_W2, W3 = BWD ? OFF_W3 : FOFF_W3;
    static constexpr int W4 = BWD ? OFF_W4 : FOFF_W4, W5 = BWD ? OFF_W5 : FOFF_W5;
};
// TMEM columns
constexpr int TC_TMP = 0;                    // 64-column scratch accumulator
constexpr int TC_DW1 = 64, TC_DW3 = 96, TC_DW4 = 128, TC_DW2T = 192, TC_DW5T = 208;   // wgrad accumulators
constexpr int TMEM_FWD = 64, TMEM_BWD = 256;

// this thread's row of a 128-row operand buffer: 8 fp16 of column chunk c
__device__ __forceinline__ uint4 ld_chunk(const uint8_t* smem, int off, int r, int c) {
    return *reinterpret_cast<const uint4*>(smem + off + c * TILE * 16 + r * 16);
}
__device__ __forceinline__ void st_row8(uint8_t* smem, int off, int r, int c, const float* v8) {
    umma::st_chunk(reinterpret_cast<__half*>(smem + off), TILE, r, c, v8);
}

struct Pipe {
    uint64_t* bar;
    uint32_t phase;
    uint32_t tm;      // TMEM base
    uint32_t lane;    // (warp * 32) << 16
};

// all threads: operands written -> visible to the tensor core; then thread 0 may issue
__device__ __forceinline__ void publish() {
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
}
__device__ __forceinline__ void wait_mma(Pipe& p) {
    umma::mbar_wait(p.bar, p.phase);
    p.phase ^= 1u;
    umma::fence_after_sync();
}
// read N (multiple of 16) accumulator columns of this thread's row
template <int N>
__device__ __forceinline__ void read_acc(const Pipe& p, int col, float* v) {
#pragma unroll
    for (int c = 0; c < N; c += 16) umma::tmem_ld16(p.tm + p.lane + (uint32_t)(col + c), v + c);
    umma::tmem_ld_wait();
}

// forward: D[128,N] = A[128,K] * W[N,K]^T
__device__ __forceinline__ void mma_fwd(uint32_t sbase, int offA, int offW, int N, int K, uint32_t tm_col) {
    const uint32_t idesc = umma::instr_desc_f16(128, N, 0, 0);
    for (int k = 0; k < K / 16; ++k)
        umma::mma_f16(tm_col, umma::desc_kmajor(sbase + offA, TILE, k), umma::desc_kmajor(sbase + offW, N, k), idesc, k > 0);
}
// dgrad: D[128,N=in] = dY[128,K=out] * W[out,in]      (W stored [Wrows x N], read MN-major)
__device__ __forceinline__ void mma_dgrad(uint32_t sbase, int offdY, int offW, int Wrows, int N, int K, uint32_t tm_col) {
    const uint32_t idesc = umma::instr_desc_f16(128, N, 0, 1);
    for (int k = 0; k < K / 16; ++k)
        umma::mma_f16(tm_col, umma::desc_kmajor(sbase + offdY, TILE, k), umma::desc_mnmajor(sbase + offW, Wrows, k), idesc, k > 0);
}
// wgrad: D[64,N] (+)= P[128 samples, 64]^T * Q[128 samples, N]   (both read MN-major, K = 128 samples)
__device__ __forceinline__ void mma_wgrad(uint32_t sbase, int offP, int offQ, int N, uint32_t tm_col, bool first_tile) {
    const uint32_t idesc = umma::instr_desc_f16(64, N, 1, 1);
    for (int k = 0; k < TILE / 16; ++k)
        umma::mma_f16(tm_col, umma::desc_mnmajor(sbase + offP, TILE, k), umma::desc_mnmajor(sbase + offQ, TILE, k), idesc,
                      !(first_tile && k == 0));
}

__device__ __forceinline__ void relu_store(uint8_t* smem, int off, int r, const float* v, int ncols) {
    for (int c = 0; c < ncols / 8; ++c) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = fmaxf(v[8 * c + j], 0.0f);
        st_row8(smem, off, r, c, t);
    }
}
// g * (act > 0) -> fp16 operand buffer
__device__ __forceinline__ void mask_store(uint8_t* smem, int off_dst, int off_act, int r, const float* g) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 a = ld_chunk(smem, off_act, r, c);
        const __half2* h = reinterpret_cast<const __half2*>(&a);
        float t[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            t[2 * j] = f.x > 0.0f ? g[8 * c + 2 * j] : 0.0f;
            t[2 * j + 1] = f.y > 0.0f ? g[8 * c + 2 * j + 1] : 0.0f;
        }
        st_row8(smem, off_dst, r, c, t);
    }
}

// 256 threads per tile: thread (row = tid & 127, half = tid >> 7) owns one half of the columns
// of sample row `row` in every epilogue (warps w and w + 4 both address TMEM lanes 32*(w&3)..),
// which doubles the warps in flight per tile and halves the serial epilogue work per thread.
constexpr int NTHREADS = 256;

// relu -> fp16 -> this thread's column half of a 64-wide operand buffer
__device__ __forceinline__ void relu_store_half(uint8_t* smem, int off, int r, int half, const float* v32) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = fmaxf(v32[8 * c + j], 0.0f);
        st_row8(smem, off, r, 4 * half + c, t);
    }
}
// g * (act > 0) -> fp16, this thread's column half
__device__ __forceinline__ void mask_store_half(uint8_t* smem, int off_dst, int off_act, int r, int half, const float* g32) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 a = ld_chunk(smem, off_act, r, 4 * half + c);
        const __half2* h = reinterpret_cast<const __half2*>(&a);
        float t[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            t[2 * j] = f.x > 0.0f ? g32[8 * c + 2 * j] : 0.0f;
            t[2 * j + 1] = f.y > 0.0f ? g32[8 * c + 2 * j + 1] : 0.0f;
        }
        st_row8(smem, off_dst, r, 4 * half + c, t);
    }
}

// CHUNKS (forward only): the encoding arrives as fp16 operand chunk planes (enc_fmt 3 / 5) and nothing else has to be
// supported: the prefetched tile is kept as the raw 16-byte chunks (8 registers instead of 16 + conversions), which
// brings the kernel under 64 registers -> 4 CTAs per SM instead of 3 for the latency-bound layer chain.
template <bool BWD, bool CHUNKS = false>
__global__ void __launch_bounds__(NTHREADS, BWD ? 2 : (CHUNKS ? 4 : 3)) mlp_kernel(const MlpArgs a) {
    vn_pdl_trigger();                         // PDL: the wait follows the prologue
    using L = Lay<BWD>;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & (TILE - 1), half = tid >> 7;
    const uint32_t sbase = umma::smem_u32(smem);

    load_weight(smem, L::W1, a.W[0], 64, 32, 64, tid);
    load_weight(smem, L::W2, a.W[1], 16, 64, 16, tid);
    load_weight(smem, L::W3, a.W[2], 64, 32, 64, tid);
    load_weight(smem, L::W4, a.W[3], 64, 64, 64, tid);
    load_weight(smem, L::W5, a.W[4], 16, 64, 3, tid);
    if (warp == 0) umma::tmem_alloc(&tmem_base, BWD ? TMEM_BWD : TMEM_FWD);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    publish();
    umma::fence_after_sync();
    Pipe p{&bar, 0u, tmem_base, (uint32_t)((warp & 3) * 32) << 16};
    const uint32_t tmp = p.tm + TC_TMP;

    const int64_t n_tiles = (a.S + TILE - 1) / TILE;
    // software prefetch: the NEXT tile's half enc row / direction (and, for the backward, its
    // output gradients) are loaded into registers while the current tile runs the layer chain
    // NOTE on the loads below: everything a PREDECESSOR KERNEL of the stream produces (enc, dsigmas, drgbs) is read with
    // ld.global.cg (__ldcg), never with the non-coherent ld.global.nc (__ldg).  An "invariant" load may be scheduled above
    // griddepcontrol.wait -- the asm memory clobber does not order it -- and then reads the buffer while the predecessor is
    // still writing it.  That happened to the 64-register instantiation of this kernel (intermittent wrong losses inside
    // the PDL-chained step, none with VN_PDL=0; profiles/r2_kbench.md).
    struct Staged { float4 e[CHUNKS ? 2 : 4]; float d[3]; float dsig; float drgb[3]; uint4 sh; };
    auto fetch = [&](int64_t tile, Staged& st) {
        const int64_t s = tile * TILE + row;
        const bool valid = tile < n_tiles && s < a.S;
        if (CHUNKS) {
            // raw chunks 2*half, 2*half+1 of this row (bit patterns travel in the float4 registers)
            st.e[0] = make_float4(0.f, 0.f, 0.f, 0.f); st.e[1] = st.e[0];
            if (valid) {
                const float4* src = reinterpret_cast<const float4*>(a.enc) + (int64_t)(2 * half) * a.S + s;
                st.e[0] = __ldcg(src); st.e[1] = __ldcg(src + a.S);
            }
        } else if (valid) {
            if (a.enc_fmt == 3 || a.enc_fmt == 5) {
                // f16 chunk planes [4][S] x 16 B: chunks 2*half, 2*half+1 of this row
                const uint4* src = reinterpret_cast<const uint4*>(a.enc) + (int64_t)(2 * half) * a.S + s;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint4 u = __ldcg(src + (int64_t)q * a.S);
                    const __half2* h = reinterpret_cast<const __half2*>(&u);
                    const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
                    st.e[2 * q] = make_float4(f0.x, f0.y, f1.x, f1.y);
                    st.e[2 * q + 1] = make_float4(f2.x, f2.y, f3.x, f3.y);
                }
            } else if (a.enc_half) {
                const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.enc) + s * 32 + 16 * half);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint4 u = __ldcg(src + q);
                    const __half2* h = reinterpret_cast<const __half2*>(&u);
                    const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
                    st.e[2 * q] = make_float4(f0.x, f0.y, f1.x, f1.y);
                    st.e[2 * q + 1] = make_float4(f2.x, f2.y, f3.x, f3.y);
                }
            } else if (a.enc_planar) {
                // planes 4*half .. 4*half+3 at index s: consecutive rows = consecutive float4
                const float4* src = reinterpret_cast<const float4*>(a.enc) + (int64_t)(4 * half) * a.S + s;
#pragma unroll
                for (int q = 0; q < 4; ++q) st.e[q] = __ldcg(src + (int64_t)q * a.S);
            } else {
                const float4* src = reinterpret_cast<const float4*>(a.enc + s * 32 + 16 * half);
#pragma unroll
                for (int q = 0; q < 4; ++q) st.e[q] = __ldcg(src + q);
            }
        } else {
#pragma unroll
            for (int q = 0; q < (CHUNKS ? 2 : 4); ++q) st.e[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        st.d[0] = 1.0f; st.d[1] = 0.0f; st.d[2] = 0.0f;
        st.sh = make_uint4(0u, 0u, 0u, 0u);
        if (a.enc_fmt == 5) {          // planes 4, 5: the direction encoding, already an operand chunk (vn_march_train_expand_sh)
            if (valid && !a.density_only) st.sh = __ldcg(reinterpret_cast<const uint4*>(a.enc) + (int64_t)(4 + half) * a.S + s);
        } else if (valid && !a.density_only) { st.d[0] = __ldg(a.dirs + 3 * s); st.d[1] = __ldg(a.dirs + 3 * s + 1); st.d[2] = __ldg(a.dirs + 3 * s + 2); }
        st.dsig = 0.0f; st.drgb[0] = st.drgb[1] = st.drgb[2] = 0.0f;
        if (BWD && valid && half == 0) {
            st.dsig = __ldcg(a.dsigmas + s);
            if (!a.density_only) { st.drgb[0] = __ldcg(a.drgbs + 3 * s); st.drgb[1] = __ldcg(a.drgbs + 3 * s + 1); st.drgb[2] = __ldcg(a.drgbs + 3 * s + 2); }
        }
    };
    Staged cur;
    // everything above read only the weights (final since the previous optimiser step, which every
    // predecessor in the stream has transitively waited for) -- from here on the producer's outputs
    vn_pdl_wait();
    fetch(blockIdx.x, cur);
    bool first_tile = true;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first_tile = false) {
        const int64_t s = tile * TILE + row;
        const bool valid = s < a.S;
        // ---- stage inputs: half enc row -> X0, half of SH(dir) -> IN2[:, 0:16] ------------------
        {
            if (CHUNKS) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    *reinterpret_cast<float4*>(smem + L::X0 + (2 * half + c) * (TILE * 16) + row * 16) = cur.e[c];
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float v[8] = {cur.e[2 * c].x, cur.e[2 * c].y, cur.e[2 * c].z, cur.e[2 * c].w,
                                        cur.e[2 * (CHUNKS ? 0 : c) + 1].x, cur.e[2 * (CHUNKS ? 0 : c) + 1].y,
                                        cur.e[2 * (CHUNKS ? 0 : c) + 1].z, cur.e[2 * (CHUNKS ? 0 : c) + 1].w};
                    st_row8(smem, L::X0, row, 2 * half + c, v);
                }
            }
            if (!a.density_only && a.enc_fmt == 5) {
                *reinterpret_cast<uint4*>(smem + L::IN2 + half * (TILE * 16) + row * 16) = cur.sh;
            } else if (!a.density_only) {
                const float dx = cur.d[0], dy = cur.d[1], dz = cur.d[2];
                const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);          // networks.py:160
                float e[16];
                sh16_half((dx / nrm + 1.0f) * 0.5f, (dy / nrm + 1.0f) * 0.5f, (dz / nrm + 1.0f) * 0.5f, e);   // :161
                float eh[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) eh[j] = half ? e[8 + j] : e[j];
                st_row8(smem, L::IN2, row, half, eh);
            }
        }
        const float my_dsig = cur.dsig;
        const float my_drgb[3] = {cur.drgb[0], cur.drgb[1], cur.drgb[2]};
        fetch(tile + gridDim.x, cur);                                         // prefetch the next tile
        publish();
        // ---- L1: 32 -> 64, ReLU -----------------------------------------------------------
        if (tid == 0) { umma::fence_after_sync(); mma_fwd(sbase, L::X0, L::W1, 64, 32, tmp); umma::commit(&bar); }
        wait_mma(p);
        float acc[32];
        read_acc<32>(p, TC_TMP + 32 * half, acc);
        relu_store_half(smem, L::H1, row, half, acc);
        publish();
        // ---- L2: 64 -> 16; sigma = exp(h0) ------------------------------------------------
        if (tid == 0) { umma::fence_after_sync(); mma_fwd(sbase, L::H1, L::W2, 16, 64, tmp); umma::commit(&bar); }
        wait_mma(p);
        read_acc<16>(p, TC_TMP, acc);
        const float h0 = acc[0];
        if (!BWD && valid && half == 0) {
            a.sigmas[s] = expf(h0);                                            // TruncExp fwd, networks.py:23
            if (a.h_out) {
                float4* hd = reinterpret_cast<float4*>(a.h_out + s * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) hd[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            }
        }
        float rgb[3] = {0.f, 0.f, 0.f};
        if (!a.density_only) {
            {
                float hh[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) hh[j] = half ? acc[8 + j] : acc[j];
                st_row8(smem, L::IN2, row, 2 + half, hh);
            }
            publish();
            // ---- L3: [SH | h] 32 -> 64, ReLU ---------------------------------------------
            if (tid == 0) { umma::fence_after_sync(); mma_fwd(sbase, L::IN2, L::W3, 64, 32, tmp); umma::commit(&bar); }
            wait_mma(p);
            read_acc<32>(p, TC_TMP + 32 * half, acc);
            relu_store_half(smem, L::H3, row, half, acc);
            publish();
            // ---- L4: 64 -> 64, ReLU ------------------------------------------------------
            if (tid == 0) { umma::fence_after_sync(); mma_fwd(sbase, L::H3, L::W4, 64, 64, tmp); umma::commit(&bar); }
            wait_mma(p);
            read_acc<32>(p, TC_TMP + 32 * half, acc);
            relu_store_half(smem, L::H4, row, half, acc);
            publish();
            // ---- L5: 64 -> 3 (padded to 16), sigmoid -------------------------------------
            if (tid == 0) { umma::fence_after_sync(); mma_fwd(sbase, L::H4, L::W5, 16, 64, tmp); umma::commit(&bar); }
            wait_mma(p);
            read_acc<16>(p, TC_TMP, acc);
#pragma unroll
            for (int c = 0; c < 3; ++c) rgb[c] = 1.0f / (1.0f + expf(-acc[c]));
            if (!BWD && valid && half == 0) { a.rgbs[3 * s] = rgb[0]; a.rgbs[3 * s + 1] = rgb[1]; a.rgbs[3 * s + 2] = rgb[2]; }
        }
        if (!BWD) { umma::fence_before_sync(); continue; }

        // =============================== backward =========================================
        float dh_sigma = 0.0f;
        if (valid && half == 0) dh_sigma = my_dsig * expf(fminf(fmaxf(h0, -15.0f), 15.0f));   // TruncExp bwd, networks.py:28
        if (!a.density_only) {
            // d(out5) = drgb * rgb * (1 - rgb) in columns 0..2 (chunk 0); chunk 1 is zero padding
            float d5[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) d5[j] = 0.0f;
            if (valid && half == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) d5[c] = my_drgb[c] * rgb[c] * (1.0f - rgb[c]);
            }
            st_row8(smem, OFF_D5, row, half, d5);
            publish();
            if (tid == 0) {
                umma::fence_after_sync();
                mma_dgrad(sbase, OFF_D5, L::W5, 16, 64, 16, tmp);                        // dH4raw = d5 * W5
                mma_wgrad(sbase, L::H4, OFF_D5, 16, p.tm + TC_DW5T, first_tile);         // dW5^T += H4^T d5
                umma::commit(&bar);
            }
            wait_mma(p);
            read_acc<32>(p, TC_TMP + 32 * half, acc);
            mask_store_half(smem, OFF_G, L::H4, row, half, acc);                         // dH4
            publish();
            if (tid == 0) {
                umma::fence_after_sync();
                mma_dgrad(sbase, OFF_G, L::W4, 64, 64, 64, tmp);                         // dH3raw = dH4 * W4
                mma_wgrad(sbase, OFF_G, L::H3, 64, p.tm + TC_DW4, first_tile);           // dW4 += dH4^T H3
                umma::commit(&bar);
            }
            wait_mma(p);
            read_acc<32>(p, TC_TMP + 32 * half, acc);
            mask_store_half(smem, OFF_G, L::H3, row, half, acc);                         // dH3 (dH4 no longer needed)
            publish();
            if (tid == 0) {
                umma::fence_after_sync();
                mma_dgrad(sbase, OFF_G, L::W3, 64, 32, 64, tmp);                         // dIN2raw = dH3 * W3
                mma_wgrad(sbase, OFF_G, L::IN2, 32, p.tm + TC_DW3, first_tile);          // dW3 += dH3^T [SH|h]
                umma::commit(&bar);
            }
            wait_mma(p);
            read_acc<16>(p, TC_TMP + 16, acc);                                           // cols 16..31 = d(h)
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
        }
        {
            float dh[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) dh[j] = valid ? (half ? acc[8 + j] : acc[j]) : 0.0f;
            if (half == 0) dh[0] += dh_sigma;
            st_row8(smem, OFF_DH, row, half, dh);
        }
        publish();
        if (tid == 0) {
            umma::fence_after_sync();
            mma_dgrad(sbase, OFF_DH, L::W2, 16, 64, 16, tmp);                            // dH1raw = dh * W2
            mma_wgrad(sbase, L::H1, OFF_DH, 16, p.tm + TC_DW2T, first_tile);             // dW2^T += H1^T dh
            umma::commit(&bar);
        }
        wait_mma(p);
        read_acc<32>(p, TC_TMP + 32 * half, acc);
        mask_store_half(smem, OFF_G, L::H1, row, half, acc);                             // dH1
        publish();
        if (tid == 0) {
            umma::fence_after_sync();
            mma_dgrad(sbase, OFF_G, L::W1, 64, 32, 64, tmp);                             // d(enc) = dH1 * W1
            mma_wgrad(sbase, OFF_G, L::X0, 32, p.tm + TC_DW1, first_tile);               // dW1 += dH1^T enc
            umma::commit(&bar);
        }
        wait_mma(p);
        read_acc<16>(p, TC_TMP + 16 * half, acc);
        if (valid) {
            if (a.enc_planar) {
                float4* dst = reinterpret_cast<float4*>(a.denc) + (int64_t)(4 * half) * a.S + s;
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[(int64_t)q * a.S] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            } else {
                float4* dst = reinterpret_cast<float4*>(a.denc + s * 32 + 16 * half);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            }
        }
        umma::fence_before_sync();
    }

    if (BWD && !first_tile) {
        // flush the weight-gradient accumulators: M = 64 layout -> row 16*warp + lane (lane < 16)
        __syncthreads();
        umma::fence_after_sync();
        if (warp < 4) {
            const int wrow = 16 * warp + lane;
            float v[64];
            // dW1 [64 out x 32 in]
            read_acc<32>(p, TC_DW1, v);
            if (lane < 16) for (int j = 0; j < 32; ++j) atomicAdd(a.dW[0] + wrow * 32 + j, v[j]);
            // dW2^T [64 in x 16 out] -> dW2 [16 x 64]
            read_acc<16>(p, TC_DW2T, v);
            if (lane < 16) for (int j = 0; j < 16; ++j) atomicAdd(a.dW[1] + j * 64 + wrow, v[j]);
            if (!a.density_only) {
                read_acc<32>(p, TC_DW3, v);
                if (lane < 16) for (int j = 0; j < 32; ++j) atomicAdd(a.dW[2] + wrow * 32 + j, v[j]);
                read_acc<64>(p, TC_DW4, v);
                if (lane < 16) for (int j = 0; j < 64; ++j) atomicAdd(a.dW[3] + wrow * 64 + j, v[j]);
                read_acc<16>(p, TC_DW5T, v);
                if (lane < 16) for (int j = 0; j < 3; ++j) atomicAdd(a.dW[4] + j * 64 + wrow, v[j]);
            }
        }
        umma::fence_before_sync();
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(p.tm, BWD ? TMEM_BWD : TMEM_FWD);
}

// VN_MLP_FWD4 = CTAs per SM of the slim chunk-format forward (0: use the generic 3-CTA kernel)
static const int g_mlp_fwd4 = []() { const char* e = getenv("VN_MLP_FWD4"); return e ? atoi(e) : 4; }();

int launch_mlp(bool bwd, const MlpArgs& a, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        VN_CUDA(cudaFuncSetAttribute(mlp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
        VN_CUDA(cudaFuncSetAttribute(mlp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
        VN_CUDA(cudaFuncSetAttribute(mlp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD));
        attr_set = true;
    }
    VnProfScope prof(bwd ? VN_K_MLP_BWD : VN_K_MLP_FWD, a.S, st);
    const int64_t n_tiles = (a.S + TILE - 1) / TILE;
    const int per_sm = bwd ? 2 : 3;
    int64_t grid = (int64_t)vn_sm_count() * per_sm;
    if (grid > n_tiles) grid = n_tiles;
    const bool chunks = !bwd && !a.density_only && (a.enc_fmt == 3 || a.enc_fmt == 5) && g_mlp_fwd4 > 0;
    if (chunks) {
        grid = (int64_t)vn_sm_count() * g_mlp_fwd4;
        if (grid > n_tiles) grid = n_tiles;
    }
    if (bwd)         vn_launch_pdl(mlp_kernel<true>, dim3((unsigned)grid), dim3(NTHREADS), SMEM_BWD, st, a);
    else if (chunks) vn_launch_pdl(mlp_kernel<false, true>, dim3((unsigned)grid), dim3(NTHREADS), SMEM_FWD, st, a);
    else             vn_launch_pdl(mlp_kernel<false>, dim3((unsigned)grid), dim3(NTHREADS), SMEM_FWD, st, a);
    VN_CHECK_LAUNCH(bwd ? "mlp_kernel<bwd>" : "mlp_kernel<fwd>");
    return VN_OK;
}

}  // namespace

static bool mlp_pipe_from_env() { const char* e = getenv("VN_MLP_PIPE"); return !(e && e[0] == '0'); }
static bool g_mlp_pipe = mlp_pipe_from_env();

VN_API int vn_mlp_fwd(const void* enc, int enc_format, const float* dirs, const float* W1, const float* W2, const float* W3,
                      const float* W4, const float* W5, int64_t S, int density_only, float* sigmas, float* rgbs,
                      float* h_out, void* stream) {
    VN_REQUIRE(S >= 0, "vn_mlp_fwd: S < 0");
    if (S == 0) return VN_OK;
    VN_REQUIRE(enc && W1 && W2 && sigmas, "vn_mlp_fwd: null pointer");
    VN_REQUIRE((enc_format >= 0 && enc_format <= 3) || enc_format == 5,
               "vn_mlp_fwd: enc_format must be 0 (f32 rows), 1 (f16 rows), 2 (f32 planes), 3 (f16 chunk planes) or 5 (f16 chunk "
               "planes + SH planes)");
    VN_REQUIRE(density_only || ((dirs || enc_format == 5) && W3 && W4 && W5 && rgbs), "vn_mlp_fwd: null colour-network pointer");
    VN_REQUIRE(vn_aligned(enc, 16), "vn_mlp_fwd: enc must be 16-byte aligned");
    VN_REQUIRE(vn_aligned(W1, 16) && vn_aligned(W2, 16) && vn_aligned(W3, 16) && vn_aligned(W4, 16) && vn_aligned(W5, 16),
               "vn_mlp_fwd: weight matrices must be 16-byte aligned");
    MlpArgs a{};
    a.enc = (const float*)enc; a.enc_half = enc_format == 1; a.enc_planar = enc_format == 2; a.enc_fmt = enc_format; a.dirs = dirs;
    a.W[0] = W1; a.W[1] = W2; a.W[2] = density_only ? W1 : W3; a.W[3] = density_only ? W1 : W4; a.W[4] = density_only ? W1 : W5;
    a.sigmas = sigmas; a.rgbs = rgbs; a.h_out = h_out; a.S = S; a.density_only = density_only;
    return launch_mlp(false, a, (cudaStream_t)stream);
}

VN_API int vn_mlp_bwd(const void* enc, int enc_format, const float* dirs, const float* W1, const float* W2, const float* W3,
                      const float* W4, const float* W5, int64_t S, int density_only, const float* dsigmas,
                      const float* drgbs, float* denc, float* dW1, float* dW2, float* dW3, float* dW4, float* dW5,
                      void* stream) {
    VN_REQUIRE(S >= 0, "vn_mlp_bwd: S < 0");
    if (S == 0) return VN_OK;
    VN_REQUIRE(enc && W1 && W2 && dsigmas && denc && dW1 && dW2, "vn_mlp_bwd: null pointer");
    const int base_fmt = enc_format & ~VN_MLP_DENC_F16;
    const bool denc_f16 = (enc_format & VN_MLP_DENC_F16) != 0;
    VN_REQUIRE(enc_format >= 0 && ((base_fmt >= 0 && base_fmt <= 3) || base_fmt == 5) && (!denc_f16 || base_fmt == 3 || base_fmt == 5),
               "vn_mlp_bwd: enc_format must be 0 (f32 rows), 1 (f16 rows), 2 (f32 planes), 3 (f16 chunk planes in, f32 planes out) "
               "or 5 (f16 chunk + SH planes in, f32 planes out); + VN_MLP_DENC_F16 (formats 3, 5): f16 chunk planes out");
    VN_REQUIRE(!(density_only && enc_format >= 3), "vn_mlp_bwd: the density-only backward takes enc_format 0..2");
    VN_REQUIRE(density_only || ((dirs || base_fmt == 5) && W3 && W4 && W5 && drgbs && dW3 && dW4 && dW5), "vn_mlp_bwd: null colour-network pointer");
    VN_REQUIRE(vn_aligned(enc, 16) && vn_aligned(denc, 16), "vn_mlp_bwd: enc/denc must be 16-byte aligned");
    VN_REQUIRE(vn_aligned(W1, 16) && vn_aligned(W2, 16) && vn_aligned(W3, 16) && vn_aligned(W4, 16) && vn_aligned(W5, 16),
               "vn_mlp_bwd: weight matrices must be 16-byte aligned");
    MlpArgs a{};
    a.enc = (const float*)enc; a.enc_half = enc_format == 1; a.enc_planar = enc_format == 2; a.dirs = dirs;
    a.W[0] = W1; a.W[1] = W2; a.W[2] = density_only ? W1 : W3; a.W[3] = density_only ? W1 : W4; a.W[4] = density_only ? W1 : W5;
    a.dsigmas = dsigmas; a.drgbs = drgbs; a.denc = denc;
    a.dW[0] = dW1; a.dW[1] = dW2; a.dW[2] = dW3; a.dW[3] = dW4; a.dW[4] = dW5;
    a.S = S; a.density_only = density_only;
    a.enc_fmt = base_fmt;
    a.denc_fmt = denc_f16 ? 3 : (base_fmt >= 2 ? 2 : 0);
    // the pipelined three-chain kernel (mlp_bwd_pipe.cu) is the backward; the serial kernel remains for the
    // density-only variant and as the A/B baseline (VN_MLP_PIPE=0)
    if (!density_only && (g_mlp_pipe || enc_format >= 3)) return launch_mlp_bwd_pipe(a, nullptr, (cudaStream_t)stream);
    return launch_mlp(true, a, (cudaStream_t)stream);
}

// MLP backward fused with the hash encoder's backward: d(enc) is scattered into the table gradient from inside the
// kernel (mlp_bwd_pipe.cu, SCATTER) instead of being written out for vn_hash_encode_bwd_*.
VN_API int vn_mlp_bwd_scatter(const void* enc, int enc_format, const float* dirs, const float* W1, const float* W2,
                              const float* W3, const float* W4, const float* W5, int64_t S, const float* dsigmas,
                              const float* drgbs, const float* xyz, const vn_hash_levels_t* lv, int round_f16,
                              float* table_grad, float* dW1, float* dW2, float* dW3, float* dW4, float* dW5, float* found_inf,
                              void* stream) {
    VN_REQUIRE(S >= 0, "vn_mlp_bwd_scatter: S < 0");
    if (S == 0) return VN_OK;
    VN_REQUIRE(enc && W1 && W2 && W3 && W4 && W5 && dsigmas && drgbs && xyz && lv && table_grad && dW1 && dW2 && dW3 && dW4 && dW5,
               "vn_mlp_bwd_scatter: null pointer");
    VN_REQUIRE(enc_format == 3 || enc_format == 5, "vn_mlp_bwd_scatter: enc_format must be 3 (f16 chunk planes) or 5 (+ SH planes)");
    VN_REQUIRE(dirs || enc_format == 5, "vn_mlp_bwd_scatter: dirs is required unless enc_format is 5");
    VN_REQUIRE(lv->levels == 16, "vn_mlp_bwd_scatter: the fused backward is built for 16 levels x 2 features (32-wide encoding)");
    VN_REQUIRE(vn_aligned(enc, 16) && vn_aligned(table_grad, 16) && vn_aligned(xyz, 4), "vn_mlp_bwd_scatter: misaligned buffer");
    VN_REQUIRE(vn_aligned(W1, 16) && vn_aligned(W2, 16) && vn_aligned(W3, 16) && vn_aligned(W4, 16) && vn_aligned(W5, 16),
               "vn_mlp_bwd_scatter: weight matrices must be 16-byte aligned");
    MlpArgs a{};
    a.enc = (const float*)enc; a.dirs = dirs;
    a.W[0] = W1; a.W[1] = W2; a.W[2] = W3; a.W[3] = W4; a.W[4] = W5;
    a.dsigmas = dsigmas; a.drgbs = drgbs; a.denc = nullptr;
    a.dW[0] = dW1; a.dW[1] = dW2; a.dW[2] = dW3; a.dW[3] = dW4; a.dW[4] = dW5;
    a.S = S; a.enc_fmt = enc_format; a.denc_fmt = round_f16 ? 3 : 2;
    ScatterArgs hs{};
    hs.xyz = xyz; hs.grad = table_grad; hs.round_f16 = round_f16 ? 1 : 0; hs.found_inf = found_inf;
    int rc = make_params(lv, hs.P);
    if (rc) return rc;
    return launch_mlp_bwd_pipe(a, &hs, (cudaStream_t)stream);
}
