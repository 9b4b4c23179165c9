// mlp_bwd_pipe.cu -- backward of the fused NGP MLPs (modules/networks.py:134-164, 195-282 under
// torch.autocast(float16)) as a warp-specialised, software-pipelined tcgen05 kernel.
//
// Same mathematics, operand layouts and rounding points as mlp_kernel<true> (mlp_fused.cu); what
// changes is how the per-tile latency chain (10 dependent MMA rounds) is hidden and shortened:
//
//   * ONE persistent CTA per SM: 3 tile chains x 256 epilogue threads + 4 MMA-issuing warps.
//     Each chain owns one 128-sample tile at a time; while chain A waits for its MMAs, chains B
//     and C run their epilogues, so three tiles are in flight per SM.  The 20 KB of fp16 weights
//     and the five weight-gradient accumulators in TMEM are shared by the chains.
//   * the A operand of every K = 64 / K = 16 product on the critical path lives in TENSOR MEMORY:
//     the epilogue packs its accumulator row to fp16 and writes it back with tcgen05.st, the next
//     round's tcgen05.mma reads it from there (no shared-memory round trip, no proxy fence), and
//     the request for that MMA goes out right after the tcgen05.st.
//   * the shared-memory copies of the activations / gradients serve only the weight gradients.
//     Their stores, the wait for the previous weight-gradient product (in-place overwrite) and
//     the request for the next one happen in the shadow of the main MMA: the weight-gradient
//     stream (warp 27, one thread => the accumulating MMAs are ordered) runs beside the critical
//     path instead of inside it.
//   * warps 24..26: one issuer per chain for the MMAs the chain waits for -- a straight-line
//     schedule with immediate descriptors that blocks on the chain's request counter (a first
//     version with ONE dispatching issuer spent ~960 cycles per round in bookkeeping and was the
//     bottleneck: profiles/r2_mlp_bwd.md).
//   * inputs arrive by bulk async copies (cp.async.bulk -> mbarrier complete_tx) one tile ahead:
//     the encoding as fp16 in the UMMA core-matrix layout straight from the hash kernel
//     (enc_fmt 3 / 5: [4][S] x 16 B chunk planes -- four 2 KB copies are the whole X0 operand),
//     with enc_fmt 5 also the direction encoding (two more planes, evaluated once per RAY by
//     vn_march_train_expand_sh) straight into the [SH | h] operand, the output gradients as flat
//     runs.  No register staging, no cvt / st.shared pass, no per-sample SH evaluation.
//   * 60 KB of activations per tile instead of 88: masked gradients overwrite their activations
//     in place, d(h) overwrites h inside the [SH | h] operand, and H1 is RECOMPUTED (one extra
//     32->64 MMA in the shadow of the dgrad3 round) instead of being kept.
//   * epilogues work on packed halves: cvt.rn.relu.f16x2.f32 for the ReLU, a half2 compare mask
//     + AND for the ReLU backward.
//
//   * SCATTER = true (vn_mlp_bwd_scatter): d(enc) never leaves the SM.  After r10 every chain thread (sample row, column
//     half) reads its 8 levels of d(enc) back from tensor memory one level at a time and runs the hash encoder's
//     warp-aggregated gradient scatter (hash_common.cuh: lanes of a warp are consecutive samples of a ray) straight into
//     the table gradient.  The red traffic of one chain runs under the MMA rounds of the other two: the LSU / L2-atomic
//     bound scatter and the latency-bound MMA chain share the SM instead of running back to back as two kernels, and
//     the 2 x 128 B / sample round trip of d(enc) through HBM disappears.
//
// Main rounds per tile (TMP = the chain's TMEM accumulator, A = its TMEM operand columns):
//   r1  TMP[0:64]  = X0(smem) W1^T      -> A = H1 = relu          ; smem HA = H1
//   r2  TMP[0:16]  = A W2^T             -> smem IN2[:,16:32] = h, keep dsigma*exp(h0)
//   r3  TMP[0:64]  = IN2(smem) W3^T     -> A = H3 = relu          ; smem HB = H3
//   r4  TMP[0:64]  = A W4^T             -> A = H4 = relu          ; smem HA = H4
//   r5  TMP[0:16]  = A W5^T             -> A = D5 = drgb rgb (1-rgb) ; smem D5       => wgrad dW5^T += HA^T D5
//   r6  TMP[0:64]  = A W5               -> A = dH4 = TMP * (HA>0) ; smem HA = dH4     => wgrad dW4   += HA^T HB
//   r7  TMP[0:64]  = A W4               -> A = dH3 = TMP * (HB>0) ; smem HB = dH3     => wgrad dW3   += HB^T IN2
//   r8  TMP[64:80] = A W3[:,16:32] ; TMP[0:64] = X0 W1^T
//                                       -> A = dh (+ sigma term)  ; smem IN2[:,16:32] = dh, HA = H1  => wgrad dW2^T += HA^T dh
//   r9  TMP[0:64]  = A W2               -> A = dH1 = TMP * (HA>0) ; smem HA = dH1     => wgrad dW1   += HA^T X0
//   r10 TMP[0:32]  = A W1               -> d(enc) to global
#include "mlp_common.cuh"
#include "hash_common.cuh"
#include <stdlib.h>

namespace mlp {
namespace {

constexpr int NCH = 3;                          // tile chains per CTA
constexpr int CH_THREADS = 256;                 // (row, column half) threads of a chain
constexpr int MMA_WARP = NCH * CH_THREADS / 32; // warps 24..26: forward / dgrad MMAs of chain 0..2 (the critical path)
constexpr int WG_WARP = MMA_WARP + NCH;         // warp 27: weight-gradient MMAs (all accumulating MMAs from one thread)
constexpr int NTHR = NCH * CH_THREADS + 128;    // 896: warps 24..26 = one forward / dgrad issuer per chain
// per-chain shared memory (bytes)
constexpr int C_X0 = 0;                         // 2 x [128 x 32] fp16 (double-buffered input tile)
constexpr int C_IN2 = C_X0 + 2 * 8192;          // [128 x 32]
constexpr int C_HA = C_IN2 + 8192;              // [128 x 64]
constexpr int C_HB = C_HA + 16384;              // [128 x 64]
constexpr int C_D5 = C_HB + 16384;              // [128 x 16]
constexpr int C_AUX = C_D5 + 4096;              // 2 x (drgb 1536 | dsig 512 | dirs 1536)
constexpr int AUX_DRGB = 0, AUX_DSIG = 1536, AUX_DIRS = 2048, AUX_SIZE = 3584;
constexpr int C_SIZE = C_AUX + 2 * AUX_SIZE;    // 68 608
static_assert(C_SIZE % 1024 == 0, "chain stride must keep the operand alignment");
// shared weights
constexpr int W_BASE = NCH * C_SIZE;
constexpr int W1_OFF = W_BASE;                  // [64 x 32]
constexpr int W2_OFF = W1_OFF + 64 * 32 * 2;    // [16 x 64]
constexpr int W3_OFF = W2_OFF + 16 * 64 * 2;    // [64 x 32]
constexpr int W4_OFF = W3_OFF + 64 * 32 * 2;    // [64 x 64]
constexpr int W5_OFF = W4_OFF + 64 * 64 * 2;    // [16 x 64], rows 3..15 zero
constexpr int SMEM_PIPE = W5_OFF + 16 * 64 * 2; // 226 304 bytes
static_assert(SMEM_PIPE <= 227 * 1024 - 256, "shared memory budget");
// TMEM columns: chain scratch c * 112 (+0: 64-wide accumulator, +64: 16-wide d(h), +80: 32 columns = the fp16
// [128 x 64] activation / gradient that is the A operand of the next round, written by the epilogue), then the
// shared weight-gradient accumulators (M = 64 layout)
constexpr int T_CH = 112, T_A = 80;
constexpr int T_DW1 = NCH * T_CH, T_DW3 = T_DW1 + 32, T_DW4 = T_DW3 + 32, T_DW2T = T_DW4 + 64, T_DW5T = T_DW2T + 16;
static_assert(T_DW5T + 16 <= 512, "TMEM budget");
constexpr int ROUNDS = 10;

// ---- small PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void red_release_inc(uint32_t saddr) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(saddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t saddr) {      // release.cta: no MEMBAR in SASS (SYNCS.ARRIVE)
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr) : "memory");
}
__device__ __forceinline__ bool mbar_test_a(uint32_t saddr, uint32_t parity) {     // non-blocking
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(saddr), "r"(parity)
        : "memory");
    return done != 0;
}
// mbarrier wait on a precomputed shared-window address (keeps the cvta / S2R out of the round loop)
__device__ __forceinline__ void mbar_wait_a(uint32_t saddr, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(saddr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(umma::smem_u32(p)) : "memory");
    return v;
}
// two fp32 -> packed fp16 (lo in bits 0..15), round to nearest; RELU: negative -> +0
template <bool RELU>
__device__ __forceinline__ uint32_t pack2(uint32_t lo_bits, uint32_t hi_bits) {
    uint32_t d;
    if (RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
    else      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
    return d;
}
// 0xffff per half where act > 0
__device__ __forceinline__ uint32_t gt0_mask(uint32_t act2) {
    const __half2 z = __float2half2_rn(0.0f);
    return __hgt2_mask(*reinterpret_cast<const __half2*>(&act2), z);
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld<2>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

// ---- descriptors with everything but the chain base folded into immediates ------------------------
// a16 = shared-memory byte address >> 4 (< 2^14 for the whole 227 KB window, so adding k-steps never
// carries into the LBO field).  K-major operand of R stored rows: LBO = R*16 (next 8 columns), SBO =
// 128 (next 8 rows), one K=16 step = 2*LBO.  MN-major: LBO = 128, SBO = R*16, one K step = 256 B.
template <int R>
__device__ __forceinline__ uint64_t dK(uint32_t a16, int k) {
    const uint32_t lo = a16 + (uint32_t)(k * 2 * R + (R << 16));
    return ((uint64_t)(0x4000u | 8u) << 32) | lo;
}
template <int R>
__device__ __forceinline__ uint64_t dMN(uint32_t a16, int k) {
    const uint32_t lo = a16 + (uint32_t)(k * 16 + (8 << 16));
    return ((uint64_t)(0x4000u | (uint32_t)R) << 32) | lo;
}

// ---- the MMA issuers: one round of one chain, round number known at compile time ---------------------
// cb16 / x16: chain base and this tile's X0 buffer (>> 4); w16: weights; tm: the chain's TMEM scratch.
// A operands of the K = 64 / K = 16 products come from tensor memory (tm + T_A, 8 columns per K = 16 step).
template <int R>
__device__ __forceinline__ void issue_main(uint32_t cb16, uint32_t x16, uint32_t w16, uint32_t tm) {
    constexpr uint32_t IN2 = C_IN2 >> 4;
    constexpr uint32_t W1 = (W1_OFF - W_BASE) >> 4, W2 = (W2_OFF - W_BASE) >> 4, W3 = (W3_OFF - W_BASE) >> 4,
                       W4 = (W4_OFF - W_BASE) >> 4, W5 = (W5_OFF - W_BASE) >> 4;
    constexpr uint32_t I_F64 = umma::instr_desc_f16(128, 64, 0, 0), I_F16 = umma::instr_desc_f16(128, 16, 0, 0);
    constexpr uint32_t I_D64 = umma::instr_desc_f16(128, 64, 0, 1), I_D32 = umma::instr_desc_f16(128, 32, 0, 1),
                       I_D16 = umma::instr_desc_f16(128, 16, 0, 1);
    const uint32_t ta = tm + T_A;
    if (R == 0) {          // r1: H1raw = X0 W1^T
#pragma unroll
        for (int k = 0; k < 2; ++k) umma::mma_f16(tm, dK<128>(x16, k), dK<64>(w16 + W1, k), I_F64, k > 0);
    } else if (R == 1) {   // r2: h = H1 W2^T
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_f16_ts(tm, ta + 8 * k, dK<16>(w16 + W2, k), I_F16, k > 0);
    } else if (R == 2) {   // r3: H3raw = IN2 W3^T
#pragma unroll
        for (int k = 0; k < 2; ++k) umma::mma_f16(tm, dK<128>(cb16 + IN2, k), dK<64>(w16 + W3, k), I_F64, k > 0);
    } else if (R == 3) {   // r4: H4raw = H3 W4^T
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_f16_ts(tm, ta + 8 * k, dK<64>(w16 + W4, k), I_F64, k > 0);
    } else if (R == 4) {   // r5: out = H4 W5^T
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_f16_ts(tm, ta + 8 * k, dK<16>(w16 + W5, k), I_F16, k > 0);
    } else if (R == 5) {   // r6: dH4raw = D5 W5              (W5 stored [16 x 64], read MN-major)
        umma::mma_f16_ts(tm, ta, dMN<16>(w16 + W5, 0), I_D64, false);
    } else if (R == 6) {   // r7: dH3raw = dH4 W4
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_f16_ts(tm, ta + 8 * k, dMN<64>(w16 + W4, k), I_D64, k > 0);
    } else if (R == 7) {   // r8: d(h)raw = dH3 W3[:,16:32] ; H1raw = X0 W1^T (recomputed)
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_f16_ts(tm + 64, ta + 8 * k, dMN<64>(w16 + W3 + 128, k), I_D16, k > 0);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma::mma_f16(tm, dK<128>(x16, k), dK<64>(w16 + W1, k), I_F64, k > 0);
    } else if (R == 8) {   // r9: dH1raw = dh W2
        umma::mma_f16_ts(tm, ta, dMN<16>(w16 + W2, 0), I_D64, false);
    } else {               // r10: d(enc) = dH1 W1
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_f16_ts(tm, ta + 8 * k, dMN<64>(w16 + W1, k), I_D32, k > 0);
    }
}
// the five weight-gradient products of a tile (P = 0 .. 4 in request order); acc0: accumulate into what is there
template <int P>
__device__ __forceinline__ void issue_wgrad(uint32_t cb16, uint32_t x16, uint32_t tmem0, bool acc0) {
    constexpr uint32_t IN2 = C_IN2 >> 4, HA = C_HA >> 4, HB = C_HB >> 4, D5 = C_D5 >> 4;
    constexpr uint32_t I_W64 = umma::instr_desc_f16(64, 64, 1, 1), I_W32 = umma::instr_desc_f16(64, 32, 1, 1),
                       I_W16 = umma::instr_desc_f16(64, 16, 1, 1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool acc = k > 0 || acc0;
        if (P == 0)      umma::mma_f16(tmem0 + T_DW5T, dMN<128>(cb16 + HA, k), dMN<128>(cb16 + D5, k), I_W16, acc);         // dW5^T += H4^T D5
        else if (P == 1) umma::mma_f16(tmem0 + T_DW4, dMN<128>(cb16 + HA, k), dMN<128>(cb16 + HB, k), I_W64, acc);          // dW4 += dH4^T H3
        else if (P == 2) umma::mma_f16(tmem0 + T_DW3, dMN<128>(cb16 + HB, k), dMN<128>(cb16 + IN2, k), I_W32, acc);         // dW3 += dH3^T IN2
        else if (P == 3) umma::mma_f16(tmem0 + T_DW2T, dMN<128>(cb16 + HA, k), dMN<128>(cb16 + IN2 + 256, k), I_W16, acc);  // dW2^T += H1^T dh
        else             umma::mma_f16(tmem0 + T_DW1, dMN<128>(cb16 + HA, k), dMN<128>(x16, k), I_W32, acc);                // dW1 += dH1^T X0
    }
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                 : "memory");
}
__device__ __forceinline__ void st_chunks4(uint8_t* buf, int row, int half, const uint32_t* h) {   // 4 chunks of a 64-wide buffer
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(buf + (4 * half + c) * (TILE * 16) + row * 16) = make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 u;
    u.x = pack2<false>(__float_as_uint(v[0]), __float_as_uint(v[1])); u.y = pack2<false>(__float_as_uint(v[2]), __float_as_uint(v[3]));
    u.z = pack2<false>(__float_as_uint(v[4]), __float_as_uint(v[5])); u.w = pack2<false>(__float_as_uint(v[6]), __float_as_uint(v[7]));
    return u;
}

#ifdef VN_MLP_TIMING
__device__ unsigned long long g_tim[64];
#define TIM_DECL unsigned long long tim_last = clock64(); const bool tim_on = (blockIdx.x == 0 && ct == 0 && c == 0);
#define TIM(slot) do { if (tim_on) { unsigned long long now = clock64(); g_tim[slot] += now - tim_last; tim_last = now; } } while (0)
#else
#define TIM_DECL
#define TIM(slot) do {} while (0)
#endif

struct Shared {
    uint64_t in_full[NCH][2];   // bulk copies of a tile's X0 / output gradients (/ dirs) landed (tx count)
    uint64_t sh_full[NCH];      // enc_fmt 5: the tile's SH chunks landed in IN2
    uint64_t done[NCH];         // the chain's current main round has completed (tcgen05.commit of warp 24 + c)
    uint64_t done_wg[NCH];      // the chain's current weight-gradient product has completed (warp 27)
    uint64_t req[NCH];          // per chain: its 8 warps have requested the next main round (phase per round)
    uint64_t req_wg[NCH];       // per chain: its 8 warps have requested the next weight-gradient product
    uint32_t tmem_base;
};

template <bool SCATTER>
__global__ void __launch_bounds__(NTHR, 1) mlp_bwd_pipe_kernel(const MlpArgs a, const __grid_constant__ ScatterArgs hs) {
    vn_pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ Shared sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = umma::smem_u32(smem);

    // ---- prologue (overlaps the predecessor's tail under PDL): weights, barriers, TMEM ------------
    load_weight(smem, W1_OFF, a.W[0], 64, 32, 64, tid);
    load_weight(smem, W2_OFF, a.W[1], 16, 64, 16, tid);
    load_weight(smem, W3_OFF, a.W[2], 64, 32, 64, tid);
    load_weight(smem, W4_OFF, a.W[3], 64, 64, 64, tid);
    load_weight(smem, W5_OFF, a.W[4], 16, 64, 3, tid);
    // chunk 1 of every chain's D5 operand (columns 8..15 of d(out5)) is zero for ever
    for (int i = tid; i < NCH * TILE; i += NTHR)
        *reinterpret_cast<uint4*>(smem + (i / TILE) * C_SIZE + C_D5 + TILE * 16 + (i % TILE) * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (warp == MMA_WARP) umma::tmem_alloc(&sh.tmem_base, 512);
    if (tid == 0) {
        for (int c = 0; c < NCH; ++c) {
            umma::mbar_init(&sh.in_full[c][0], 1); umma::mbar_init(&sh.in_full[c][1], 1); umma::mbar_init(&sh.sh_full[c], 1);
            umma::mbar_init(&sh.done[c], 1); umma::mbar_init(&sh.done_wg[c], 1);
            umma::mbar_init(&sh.req[c], 8); umma::mbar_init(&sh.req_wg[c], 8);
        }
        umma::mbar_fence_init();
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem0 = sh.tmem_base;

    const int64_t n_tiles = (a.S + TILE - 1) / TILE;
    const int64_t stride = (int64_t)NCH * gridDim.x;
    const bool chunks = a.enc_fmt == 3 || a.enc_fmt == 5;     // X0 arrives as bulk-copied fp16 operand chunks
    const bool fmt5 = a.enc_fmt == 5;                          // ... and so does the direction encoding
    // flat 16-byte-aligned input arrays can be bulk-copied tile-wise (full tiles only)
    const bool bulk_small = ((reinterpret_cast<uintptr_t>(a.drgbs) | reinterpret_cast<uintptr_t>(a.dsigmas) |
                              (fmt5 ? (uintptr_t)0 : reinterpret_cast<uintptr_t>(a.dirs))) & 15u) == 0;

    if (warp >= MMA_WARP) {
        // ============================ MMA issuers ==============================================
        // The whole warp runs the uniform control flow; one elected lane issues (no per-instruction election loop).
        const uint32_t s16 = sbase >> 4, w16 = (sbase + W_BASE) >> 4;
        if (warp < WG_WARP) {
            // chain c's own issuer: straight-line (tile, round) schedule, blocks on the chain's request counter
            const int c = warp - MMA_WARP;
            const int64_t first = (int64_t)c * gridDim.x + blockIdx.x;
            const uint32_t nt = first < n_tiles ? (uint32_t)((n_tiles - first + stride - 1) / stride) : 0u;
            const uint32_t req_c = umma::smem_u32(&sh.req[c]), done_c = umma::smem_u32(&sh.done[c]);
            const uint32_t cb16 = s16 + (uint32_t)c * (C_SIZE >> 4), tmc = tmem0 + (uint32_t)c * T_CH;
            for (uint32_t t = 0; t < nt; ++t) {      // ROUNDS is even: the request phase parity of round R is R & 1
                const uint32_t x16 = cb16 + (t & 1u) * (8192u >> 4);
#define VN_MAIN_ROUND(R)                                                                                              \
    {                                                                                                                 \
        mbar_wait_a(req_c, (R) & 1u);                                                                                 \
        umma::fence_after_sync();                                                                                     \
        if (elect_one()) {                                                                                            \
            issue_main<R>(cb16, x16, w16, tmc);                                                                       \
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done_c) : "memory"); \
        }                                                                                                             \
        __syncwarp();                                                                                                 \
    }
                VN_MAIN_ROUND(0) VN_MAIN_ROUND(1) VN_MAIN_ROUND(2) VN_MAIN_ROUND(3) VN_MAIN_ROUND(4)
                VN_MAIN_ROUND(5) VN_MAIN_ROUND(6) VN_MAIN_ROUND(7) VN_MAIN_ROUND(8) VN_MAIN_ROUND(9)
#undef VN_MAIN_ROUND
            }
        } else {
            // the weight-gradient products of all chains: serves whichever chain has requested its next product
            uint32_t nt[NCH], wt[NCH], wp[NCH];
            uint32_t remaining = 0, started = 0u;
            const uint32_t req_wg0 = umma::smem_u32(&sh.req_wg[0]), done_wg0 = umma::smem_u32(&sh.done_wg[0]);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int64_t first = (int64_t)c * gridDim.x + blockIdx.x;
                nt[c] = first < n_tiles ? (uint32_t)((n_tiles - first + stride - 1) / stride) : 0u;
                wt[c] = 0u; wp[c] = 0u;
                remaining += 5u * nt[c];
            }
            while (remaining) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (wt[c] < nt[c]) {
                        // phase number of this request = wt * 5 + wp; a chain requests product k+1 only after it has waited
                        // for product k, so the barrier is never more than one phase ahead of this test.  (mbarrier.test_wait
                        // instead of polling a counter with ld.shared: a tight ld.shared spin measurably slows the chains.)
                        if (mbar_test_a(req_wg0 + 8u * c, (wt[c] * 5u + wp[c]) & 1u)) {
                            umma::fence_after_sync();
                            const uint32_t cb16 = s16 + (uint32_t)c * (C_SIZE >> 4);
                            const uint32_t x16 = cb16 + (wt[c] & 1u) * (8192u >> 4);
                            const uint32_t bit = 1u << wp[c];
                            const bool acc0 = (started & bit) != 0u;
                            if (elect_one()) {
                                switch (wp[c]) {
                                    case 0: issue_wgrad<0>(cb16, x16, tmem0, acc0); break;
                                    case 1: issue_wgrad<1>(cb16, x16, tmem0, acc0); break;
                                    case 2: issue_wgrad<2>(cb16, x16, tmem0, acc0); break;
                                    case 3: issue_wgrad<3>(cb16, x16, tmem0, acc0); break;
                                    default: issue_wgrad<4>(cb16, x16, tmem0, acc0); break;
                                }
                                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done_wg0 + 8u * c) : "memory");
                            }
                            __syncwarp();
                            started |= bit;
                            --remaining;
                            if (++wp[c] == 5u) { wp[c] = 0u; ++wt[c]; }
                        }
                    }
                }
            }
        }
    } else {
        // ============================ tile chains ==============================================
        const int c = warp >> 3, q = warp & 3, half = (warp >> 2) & 1;
        const int row = 32 * q + lane;
        const int ct = tid - c * CH_THREADS;
        uint8_t* cs = smem + c * C_SIZE;
        const uint32_t cs32 = sbase + (uint32_t)c * C_SIZE;
        const uint32_t tm = tmem0 + (uint32_t)c * T_CH + ((uint32_t)(32 * q) << 16);
        const uint32_t ta = tm + T_A + 16 * half;      // this thread's half row of the TMEM-resident A operand
        const uint32_t done = umma::smem_u32(&sh.done[c]), done_wg = umma::smem_u32(&sh.done_wg[c]);
        const uint32_t req = umma::smem_u32(&sh.req[c]), req_wg = umma::smem_u32(&sh.req_wg[c]);
        const uint32_t in_full0 = umma::smem_u32(&sh.in_full[c][0]), sh_full = umma::smem_u32(&sh.sh_full[c]);
        uint32_t done_phase = 0u, wg_phase = 0u, in_phase = 0u;   // in_phase: bit p = parity to wait for on in_full[c][p]
        uint8_t* const bufA = cs + C_HA;
        uint8_t* const bufB = cs + C_HB;

        // request the next main round: operands in tensor memory are ordered by the tcgen05 fences; SMEM = this
        // epilogue also wrote a shared-memory operand of that round
        auto post_main = [&](bool smem_operand) {
            if (smem_operand) umma::fence_async_smem();
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(req);
        };
        // request the next weight-gradient product (both operands in shared memory)
        auto post_wg = [&]() {
            umma::fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(req_wg);
        };
        auto wait_done = [&]() {
            mbar_wait_a(done, done_phase);
            done_phase ^= 1u;
            umma::fence_after_sync();
        };
        // the previous weight-gradient product has read its operands: they may be overwritten
        auto wait_wgrad = [&]() {
            mbar_wait_a(done_wg, wg_phase);
            wg_phase ^= 1u;
        };
        auto tile_rows = [&](int64_t t) { const int64_t left = a.S - t * TILE; return (int)(left < TILE ? left : TILE); };
        auto issue_loads = [&](int64_t t, int p) {      // one thread per chain: X0 chunks + flat output gradients (+ dirs)
            const int rows = tile_rows(t);
            const bool small = bulk_small && rows == TILE;
            if (!chunks && !small) return;
            const uint32_t bar = in_full0 + 8u * p;
            const uint32_t small_bytes = fmt5 ? 2048u : (uint32_t)AUX_SIZE;
            mbar_expect_tx(bar, (chunks ? 4u * rows * 16u : 0u) + (small ? small_bytes : 0u));
            if (chunks) {
                const uint4* src = reinterpret_cast<const uint4*>(a.enc) + t * TILE;
#pragma unroll
                for (int k = 0; k < 4; ++k) bulk_g2s(cs32 + C_X0 + p * 8192 + k * 2048, src + (int64_t)k * a.S, rows * 16u, bar);
            }
            if (small) {
                const uint32_t aux = cs32 + C_AUX + p * AUX_SIZE;
                bulk_g2s(aux + AUX_DRGB, a.drgbs + t * TILE * 3, 1536u, bar);
                bulk_g2s(aux + AUX_DSIG, a.dsigmas + t * TILE, 512u, bar);
                if (!fmt5) bulk_g2s(aux + AUX_DIRS, a.dirs + t * TILE * 3, 1536u, bar);
            }
        };
        auto issue_sh = [&](int64_t t) {                // enc_fmt 5: planes 4, 5 -> chunks 0, 1 of IN2
            const uint32_t rows = (uint32_t)tile_rows(t);
            mbar_expect_tx(sh_full, 2u * rows * 16u);
            const uint4* src = reinterpret_cast<const uint4*>(a.enc) + 4 * a.S + t * TILE;
            bulk_g2s(cs32 + C_IN2, src, rows * 16u, sh_full);
            bulk_g2s(cs32 + C_IN2 + 2048, src + a.S, rows * 16u, sh_full);
        };

        const int64_t first = (int64_t)c * gridDim.x + blockIdx.x;
        // everything above read only the weights; from here on the producers' outputs
        vn_pdl_wait();
        if (ct == 0 && first < n_tiles) { issue_loads(first, 0); if (fmt5) issue_sh(first); }
        int it = 0;
        TIM_DECL
        for (int64_t tile = first; tile < n_tiles; tile += stride, ++it) {
            const int p = it & 1;
            const int rows = tile_rows(tile);
            const bool small = bulk_small && rows == TILE;
            const int64_t s = tile * TILE + row;
            const bool valid = row < rows;
            const uint8_t* aux = cs + C_AUX + p * AUX_SIZE;
            uint8_t* x0 = cs + C_X0 + p * 8192;
            if (chunks || small) {
                mbar_wait_a(in_full0 + 8u * p, (in_phase >> p) & 1u);
                in_phase ^= 1u << p;
            }
            // ---- stage: X0 and SH(dir) -> IN2[:, 0:16], unless they arrived by bulk copy --------------
            if (chunks) {
                if (!valid) {          // tail tile: rows the copy did not write must be finite (they meet zero gradients in the wgrads)
                    *reinterpret_cast<uint4*>(x0 + (2 * half) * (TILE * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(x0 + (2 * half + 1) * (TILE * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
                float e[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) e[k] = 0.0f;
                if (valid) {
                    if (a.enc_fmt == 1) {
                        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.enc) + s * 32 + 16 * half);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint4 u = __ldcg(src + k);
                            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); e[8 * k + 2 * j] = f.x; e[8 * k + 2 * j + 1] = f.y; }
                        }
                    } else {
                        const float4* src = a.enc_fmt == 2 ? reinterpret_cast<const float4*>(a.enc) + (int64_t)(4 * half) * a.S + s
                                                           : reinterpret_cast<const float4*>(a.enc + s * 32 + 16 * half);
                        const int64_t step = a.enc_fmt == 2 ? a.S : 1;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 f = __ldcg(src + k * step);
                            e[4 * k] = f.x; e[4 * k + 1] = f.y; e[4 * k + 2] = f.z; e[4 * k + 3] = f.w;
                        }
                    }
                }
                *reinterpret_cast<uint4*>(x0 + (2 * half) * (TILE * 16) + row * 16) = pack8(e);
                *reinterpret_cast<uint4*>(x0 + (2 * half + 1) * (TILE * 16) + row * 16) = pack8(e + 8);
            }
            if (fmt5) {
                mbar_wait_a(sh_full, (uint32_t)(it & 1));
                if (!valid) *reinterpret_cast<uint4*>(cs + C_IN2 + half * (TILE * 16) + row * 16) = make_uint4(0u, 0u, 0u, 0u);
            } else {
                float dx = 1.0f, dy = 0.0f, dz = 0.0f;
                if (small) {
                    const float* d = reinterpret_cast<const float*>(aux + AUX_DIRS) + 3 * row;
                    dx = d[0]; dy = d[1]; dz = d[2];
                } else if (valid) {
                    dx = __ldg(a.dirs + 3 * s); dy = __ldg(a.dirs + 3 * s + 1); dz = __ldg(a.dirs + 3 * s + 2);
                }
                const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);                                    // networks.py:160
                float e[16];
                sh16_half((dx / nrm + 1.0f) * 0.5f, (dy / nrm + 1.0f) * 0.5f, (dz / nrm + 1.0f) * 0.5f, e);   // :161
                float eh[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) eh[k] = half ? e[8 + k] : e[k];
                *reinterpret_cast<uint4*>(cs + C_IN2 + half * (TILE * 16) + row * 16) = pack8(eh);
            }
            post_main(true);                                               // -> r1 (X0 from shared memory)
            TIM(2);
            // ---- r1: H1 -> A (for r2), shared HA (for nothing yet: H1 is recomputed for its weight gradient) ----
            wait_done();
            TIM(3);
            {
                uint32_t r[32], h[16];
                tmem_ld<32>(tm + 32 * half, r);
                umma::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = pack2<true>(r[2 * j], r[2 * j + 1]);
                umma::tmem_st16(ta, h);
                umma::tmem_st_wait();
            }
            post_main(false);                                              // -> r2
            TIM(4);
            // ---- r2: h -> IN2[:,16:32]; sigma-branch gradient seed -------------------------------------
            wait_done();
            TIM(5);
            float dh_sigma = 0.0f;
            {
                uint32_t r[8];
                tmem_ld<8>(tm + 8 * half, r);
                umma::tmem_ld_wait();
                uint4 u;
                u.x = pack2<false>(r[0], r[1]); u.y = pack2<false>(r[2], r[3]); u.z = pack2<false>(r[4], r[5]); u.w = pack2<false>(r[6], r[7]);
                *reinterpret_cast<uint4*>(cs + C_IN2 + (2 + half) * (TILE * 16) + row * 16) = u;
                if (half == 0 && valid) {
                    const float dsig = small ? reinterpret_cast<const float*>(aux + AUX_DSIG)[row] : __ldcg(a.dsigmas + s);
                    dh_sigma = dsig * expf(fminf(fmaxf(__uint_as_float(r[0]), -15.0f), 15.0f));          // TruncExp bwd, networks.py:28
                }
            }
            post_main(true);                                               // -> r3 (IN2 from shared memory)
            TIM(6);
            // ---- r3: H3 -> A (r4), shared HB (dW4) -------------------------------------------------------
            wait_done();
            TIM(7);
            {
                uint32_t r[32], h[16];
                tmem_ld<32>(tm + 32 * half, r);
                umma::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = pack2<true>(r[2 * j], r[2 * j + 1]);
                umma::tmem_st16(ta, h);
                umma::tmem_st_wait();
                post_main(false);                                          // -> r4
                st_chunks4(bufB, row, half, h);                            // in the shadow of r4
            }
            TIM(8);
            // ---- r4: H4 -> A (r5), shared HA (dW5, ReLU mask of r6) --------------------------------------
            wait_done();
            TIM(9);
            {
                uint32_t r[32], h[16];
                tmem_ld<32>(tm + 32 * half, r);
                umma::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = pack2<true>(r[2 * j], r[2 * j + 1]);
                umma::tmem_st16(ta, h);
                umma::tmem_st_wait();
                post_main(false);                                          // -> r5
                st_chunks4(bufA, row, half, h);
            }
            TIM(10);
            // ---- r5: rgb -> d(out5) -> A (r6), shared D5 (dW5) -------------------------------------------
            wait_done();
            TIM(11);
            if (half == 0) {
                uint32_t r[8];
                tmem_ld<8>(tm, r);
                umma::tmem_ld_wait();
                float d5[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (valid) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float g = small ? reinterpret_cast<const float*>(aux + AUX_DRGB)[3 * row + k] : __ldcg(a.drgbs + 3 * s + k);
                        const float rgb = __fdividef(1.0f, 1.0f + __expf(-__uint_as_float(r[k])));
                        d5[k] = g * rgb * (1.0f - rgb);
                    }
                }
                const uint4 u = pack8(d5);
                const uint32_t w[8] = {u.x, u.y, u.z, u.w, 0u, 0u, 0u, 0u};     // K = 16: columns 8..15 are zero
                tmem_st8(tm + T_A, w);
                umma::tmem_st_wait();
                post_main(false);                                          // -> r6
                *reinterpret_cast<uint4*>(cs + C_D5 + row * 16) = u;
            } else {
                post_main(false);
                // the half-1 warps have nothing else to do in this round: one of their threads prefetches the next tile
                // (its buffers were released by the previous tile's last weight-gradient product)
                if (ct == CH_THREADS / 2 && tile + stride < n_tiles) issue_loads(tile + stride, p ^ 1);
            }
            post_wg();                                                     // => dW5^T += H4^T D5
            TIM(12);
            // ---- r6: dH4 -> A (r7), shared HA in place (dW4) ---------------------------------------------
            {
                uint4 act[4];          // this thread's own activation chunks (ReLU mask): loaded in the shadow of the MMA
#pragma unroll
                for (int k = 0; k < 4; ++k) act[k] = *reinterpret_cast<const uint4*>(bufA + (4 * half + k) * (TILE * 16) + row * 16);
                wait_done();
                TIM(13);
                uint32_t r[32], h[16];
                tmem_ld<32>(tm + 32 * half, r);
                umma::tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    h[4 * k] = pack2<false>(r[8 * k], r[8 * k + 1]) & gt0_mask(act[k].x);
                    h[4 * k + 1] = pack2<false>(r[8 * k + 2], r[8 * k + 3]) & gt0_mask(act[k].y);
                    h[4 * k + 2] = pack2<false>(r[8 * k + 4], r[8 * k + 5]) & gt0_mask(act[k].z);
                    h[4 * k + 3] = pack2<false>(r[8 * k + 6], r[8 * k + 7]) & gt0_mask(act[k].w);
                }
                umma::tmem_st16(ta, h);
                umma::tmem_st_wait();
                post_main(false);                                          // -> r7
                wait_wgrad();                                              // dW5 has read H4 (HA) and D5
                st_chunks4(bufA, row, half, h);
            }
            post_wg();                                                     // => dW4 += dH4^T H3
            TIM(14);
            // ---- r7: dH3 -> A (r8), shared HB in place (dW3) ---------------------------------------------
            {
                uint4 act[4];          // this thread's own activation chunks (ReLU mask): loaded in the shadow of the MMA
#pragma unroll
                for (int k = 0; k < 4; ++k) act[k] = *reinterpret_cast<const uint4*>(bufB + (4 * half + k) * (TILE * 16) + row * 16);
                wait_done();
                TIM(15);
                uint32_t r[32], h[16];
                tmem_ld<32>(tm + 32 * half, r);
                umma::tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    h[4 * k] = pack2<false>(r[8 * k], r[8 * k + 1]) & gt0_mask(act[k].x);
                    h[4 * k + 1] = pack2<false>(r[8 * k + 2], r[8 * k + 3]) & gt0_mask(act[k].y);
                    h[4 * k + 2] = pack2<false>(r[8 * k + 4], r[8 * k + 5]) & gt0_mask(act[k].z);
                    h[4 * k + 3] = pack2<false>(r[8 * k + 6], r[8 * k + 7]) & gt0_mask(act[k].w);
                }
                umma::tmem_st16(ta, h);
                umma::tmem_st_wait();
                post_main(false);                                          // -> r8
                wait_wgrad();                                              // dW4 has read dH4 (HA) and H3 (HB)
                st_chunks4(bufB, row, half, h);
            }
            post_wg();                                                     // => dW3 += dH3^T IN2
            TIM(16);
            // ---- r8: d(h) -> A (r9), shared IN2[:,16:32]; H1 again -> shared HA (dW2, ReLU mask of r9) -----
            wait_done();
            TIM(17);
            {
                uint32_t r8[8];
                tmem_ld<8>(tm + 64 + 8 * half, r8);
                uint32_t r[32], h[16];
                tmem_ld<32>(tm + 32 * half, r);
                umma::tmem_ld_wait();
                float dh[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) dh[k] = valid ? __uint_as_float(r8[k]) : 0.0f;
                if (half == 0) dh[0] += dh_sigma;
                const uint4 u = pack8(dh);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
                tmem_st4(tm + T_A + 4 * half, w);                          // K = 16: half 0 -> columns 0..3, half 1 -> 4..7
                umma::tmem_st_wait();
                post_main(false);                                          // -> r9
#pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = pack2<true>(r[2 * j], r[2 * j + 1]);
                wait_wgrad();                                              // dW3 has read dH3 (HB) and IN2 (and dW4 is older: HA is free)
                if (fmt5 && ct == 0 && tile + stride < n_tiles) issue_sh(tile + stride);   // IN2[:, 0:16] is free: next tile's SH
                *reinterpret_cast<uint4*>(cs + C_IN2 + (2 + half) * (TILE * 16) + row * 16) = u;
                st_chunks4(bufA, row, half, h);
            }
            post_wg();                                                     // => dW2^T += H1^T dh
            TIM(18);
            // ---- r9: dH1 -> A (r10), shared HA in place (dW1) --------------------------------------------
            {
                uint4 act[4];          // this thread's own activation chunks (ReLU mask): loaded in the shadow of the MMA
#pragma unroll
                for (int k = 0; k < 4; ++k) act[k] = *reinterpret_cast<const uint4*>(bufA + (4 * half + k) * (TILE * 16) + row * 16);
                wait_done();
                TIM(19);
                uint32_t r[32], h[16];
                tmem_ld<32>(tm + 32 * half, r);
                umma::tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    h[4 * k] = pack2<false>(r[8 * k], r[8 * k + 1]) & gt0_mask(act[k].x);
                    h[4 * k + 1] = pack2<false>(r[8 * k + 2], r[8 * k + 3]) & gt0_mask(act[k].y);
                    h[4 * k + 2] = pack2<false>(r[8 * k + 4], r[8 * k + 5]) & gt0_mask(act[k].z);
                    h[4 * k + 3] = pack2<false>(r[8 * k + 6], r[8 * k + 7]) & gt0_mask(act[k].w);
                }
                umma::tmem_st16(ta, h);
                umma::tmem_st_wait();
                post_main(false);                                          // -> r10
                wait_wgrad();                                              // dW2 has read H1 (HA) and dh
                st_chunks4(bufA, row, half, h);
            }
            post_wg();                                                     // => dW1 += dH1^T X0
            TIM(20);
            // ---- r10: d(enc) ------------------------------------------------------------------------
            float px = 0.0f, py = 0.0f, pz = 0.0f;
            if (SCATTER && valid) {        // issued before the wait: the load latency hides behind the MMA
                px = __ldg(hs.xyz + 3 * s); py = __ldg(hs.xyz + 3 * s + 1); pz = __ldg(hs.xyz + 3 * s + 2);
            }
            wait_done();
            TIM(21);
            if (SCATTER) {
                // this thread: sample `row`, levels half, half + 2, .. (TMEM columns 2 * level, + 1).  The two column halves
                // take the even / the odd levels so that both get the same mix of coarse (few atomics per warp) and fine
                // (one run per lane) levels: the chain waits for the slower of its warps.
#pragma unroll 1
                for (int l = 0; l < 8; ++l) {
                    const int level = 2 * l + half;
                    uint32_t r2[2];
                    tmem_ld<2>(tm + 2 * level, r2);
                    umma::tmem_ld_wait();
                    float d0 = __uint_as_float(r2[0]), d1 = __uint_as_float(r2[1]);
                    if (hs.round_f16) {        // half-precision encoder: d(enc) is an fp16 tensor (hash_encoder_half.py:344)
                        const uint32_t u = pack2<false>(r2[0], r2[1]);
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u));
                        d0 = f.x; d1 = f.y;
                    }
                    // GradScaler's inf check (trainer.py:140) rides here: a table-gradient entry is non-finite iff one of
                    // its contributions is (|d(enc)| <= 64 * 65504^2 otherwise, so no sum of < 2^31 finite terms overflows)
                    if (hs.found_inf && valid && !(fabsf(d0) + fabsf(d1) < INFINITY)) *hs.found_inf = 1.0f;
                    const bool v = valid && !(d0 == 0.0f && d1 == 0.0f);           // hash_encoder_half.py:210
                    float* gl = hs.grad + 2 * (size_t)hs.P.offsets[level];
                    const Cell cl = cell_of(px, py, pz, hs.P.scales[level]);
                    if (level < hs.P.begin_fast)
                        level_scatter<float, true, true, true>(gl, cl, hs.P.res[level], hs.P.sizes[level], 0u, d0, d1, v);
                    else
                        level_scatter<float, false, true, true>(gl, cl, hs.P.res[level], hs.P.sizes[level], hs.P.pow2mask[level], d0, d1, v);
                }
            } else {
                uint32_t r[16];
                tmem_ld<16>(tm + 16 * half, r);
                umma::tmem_ld_wait();
                if (valid) {
                    if (a.denc_fmt == 2) {
                        float4* dst = reinterpret_cast<float4*>(a.denc) + (int64_t)(4 * half) * a.S + s;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            dst[(int64_t)k * a.S] = make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]),
                                                               __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
                    } else if (a.denc_fmt == 3) {
                        uint4* dst = reinterpret_cast<uint4*>(a.denc) + (int64_t)(2 * half) * a.S + s;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            uint4 u;
                            u.x = pack2<false>(r[8 * k], r[8 * k + 1]); u.y = pack2<false>(r[8 * k + 2], r[8 * k + 3]);
                            u.z = pack2<false>(r[8 * k + 4], r[8 * k + 5]); u.w = pack2<false>(r[8 * k + 6], r[8 * k + 7]);
                            dst[(int64_t)k * a.S] = u;
                        }
                    } else {
                        float4* dst = reinterpret_cast<float4*>(a.denc + s * 32 + 16 * half);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            dst[k] = make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]),
                                                 __uint_as_float(r[4 * k + 3]));
                    }
                }
            }
            wait_wgrad();                                                  // dW1 has read dH1 (HA) and this tile's X0
            TIM(22);
            umma::fence_before_sync();
        }
    }

    // ---- flush the weight-gradient accumulators (M = 64 layout: row i sits in TMEM lane 32*(i/16) + i%16) ----
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (blockIdx.x < n_tiles && warp < 24) {
        // warps with the same (warp & 3) address the same TMEM lanes; the six warp groups split the columns
        const int q = warp & 3, grp = warp >> 2;
        const int wrow = 16 * q + lane;
        const uint32_t tl = tmem0 + ((uint32_t)(32 * q) << 16);
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
        if (grp == 0) {                       // dW1 [64 out x 32 in]
            tmem_ld<32>(tl + T_DW1, v); umma::tmem_ld_wait();
            if (lane < 16) for (int j = 0; j < 32; ++j) atomicAdd(a.dW[0] + wrow * 32 + j, __uint_as_float(v[j]));
        } else if (grp == 1) {                // dW3 [64 x 32]
            tmem_ld<32>(tl + T_DW3, v); umma::tmem_ld_wait();
            if (lane < 16) for (int j = 0; j < 32; ++j) atomicAdd(a.dW[2] + wrow * 32 + j, __uint_as_float(v[j]));
        } else if (grp == 2 || grp == 3) {    // dW4 [64 x 64], 32 columns each
            const int c0 = 32 * (grp - 2);
            tmem_ld<32>(tl + T_DW4 + c0, v); umma::tmem_ld_wait();
            if (lane < 16) for (int j = 0; j < 32; ++j) atomicAdd(a.dW[3] + wrow * 64 + c0 + j, __uint_as_float(v[j]));
        } else if (grp == 4) {                // dW2^T [64 in x 16 out] -> dW2 [16 x 64]
            tmem_ld<16>(tl + T_DW2T, v); umma::tmem_ld_wait();
            if (lane < 16) for (int j = 0; j < 16; ++j) atomicAdd(a.dW[1] + j * 64 + wrow, __uint_as_float(v[j]));
        } else {                              // dW5^T [64 in x 16] -> dW5 [3 x 64]
            tmem_ld<8>(tl + T_DW5T, v); umma::tmem_ld_wait();
            if (lane < 16) for (int j = 0; j < 3; ++j) atomicAdd(a.dW[4] + j * 64 + wrow, __uint_as_float(v[j]));
        }
        if (SCATTER && hs.found_inf && lane < 16) {   // the weight gradients' share of the inf check (lanes 16..31 hold no rows)
            float m = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) m += fabsf(__uint_as_float(v[j]));
            if (!(m < INFINITY)) *hs.found_inf = 1.0f;
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) umma::tmem_dealloc(tmem0, 512);
}

}  // namespace

#ifdef VN_MLP_TIMING
extern "C" __attribute__((visibility("default"))) int vn_mlp_debug_timing(unsigned long long* out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, g_tim, sizeof(unsigned long long) * 64);
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_tim, z, sizeof(z)); }
    return 0;
}
#endif

int launch_mlp_bwd_pipe(const MlpArgs& a, const ScatterArgs* hs, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        VN_CUDA(cudaFuncSetAttribute(mlp_bwd_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_PIPE));
        VN_CUDA(cudaFuncSetAttribute(mlp_bwd_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_PIPE));
        attr_set = true;
    }
    VnProfScope prof(hs ? VN_K_MLP_BWD_SCATTER : VN_K_MLP_BWD, a.S, st);
    const int64_t n_tiles = (a.S + TILE - 1) / TILE;
    int64_t grid = vn_sm_count();
    if (grid > n_tiles) grid = n_tiles;
    if (hs) {
        vn_launch_pdl(mlp_bwd_pipe_kernel<true>, dim3((unsigned)grid), dim3(NTHR), SMEM_PIPE, st, a, *hs);
    } else {
        static const ScatterArgs none{};
        vn_launch_pdl(mlp_bwd_pipe_kernel<false>, dim3((unsigned)grid), dim3(NTHR), SMEM_PIPE, st, a, none);
    }
    VN_CHECK_LAUNCH("mlp_bwd_pipe_kernel");
    return VN_OK;
}

}  // namespace mlp
