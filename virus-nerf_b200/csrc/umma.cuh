// umma.cuh -- thin inline-PTX layer over the Blackwell 5th-gen tensor cores (tcgen05):
// TMEM allocation, shared-memory matrix descriptors (no-swizzle canonical layouts),
// instruction descriptors for kind::f16, MMA issue, commit -> mbarrier, TMEM loads.
//
// Shared-memory operand layout used throughout ("chunk-major core matrices"): a [R x C] fp16
// matrix is stored as C/8 column chunks; a chunk holds R rows x 8 elements (16 bytes per
// row), rows contiguous:  addr(r, c) = (c/8) * (R*16) + r*16 + (c%8)*2.
// Every 8 rows x 16 bytes = 128 contiguous bytes form one UMMA core matrix, so the same buffer
// is readable
//   * as a K-major operand  (MN = rows, K = cols): SBO = 128 (next 8 rows), LBO = R*16 (next
//     8 columns); one K=16 step advances the start address by 2*LBO;
//   * as an MN-major operand (MN = cols, K = rows): SBO = R*16 (next 8 columns), LBO = 128
//     (next 8 rows); one K=16 step advances the start address by 256 bytes.
// (canonical INTERLEAVE layouts of cute/atom/mma_traits_sm100.hpp:make_umma_desc)
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM ------------------------------------------------------------------------------
// one full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *dst
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier --------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies
// tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- descriptors -----------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version bit set
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor for kind::f16: fp16 x fp16 -> fp32, M x N, operand major-ness
// (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                    /* c_format = F32 */
           | (0u << 7) | (0u << 10)     /* a_format = b_format = F16 */
           | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows on the 128 lanes, K fp16 values packed two per 32-bit
// column, K-major) is read from tensor memory instead of shared memory; one K=16 step = 8 columns
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// chunk-major operand helpers (see file header).  R = rows of the stored matrix.
// K-major use: MN = rows (128 for activations, `out` for weights), K = cols
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr, int R, int kstep) {
    const uint32_t lbo = (uint32_t)R * 16u;
    return smem_desc(saddr + (uint32_t)kstep * 2u * lbo, lbo, 128u);
}
// MN-major use: MN = cols, K = rows
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, int R, int kstep) {
    return smem_desc(saddr + (uint32_t)kstep * 256u, 128u, (uint32_t)R * 16u);
}

// ---- TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns -----------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// store 8 fp16 values (one 16-byte chunk) of row r, column chunk c of an [R x C] operand buffer
__device__ __forceinline__ void st_chunk(__half* buf, int R, int r, int c, const float* v8) {
    uint4 u;
    __half2 h0 = __floats2half2_rn(v8[0], v8[1]), h1 = __floats2half2_rn(v8[2], v8[3]);
    __half2 h2 = __floats2half2_rn(v8[4], v8[5]), h3 = __floats2half2_rn(v8[6], v8[7]);
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(reinterpret_cast<char*>(buf) + (size_t)c * R * 16 + (size_t)r * 16) = u;
}

}  // namespace umma
