// mlp_fused.cu -- fully fused density / colour MLPs of the NGP model on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM).  Replaces the cuBLAS GEMMs + element-wise
// launches of modules/networks.py:134-164, 195-282 and modules/spherical_harmonics.py.
#include "common.cuh"
#include "umma.cuh"

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
