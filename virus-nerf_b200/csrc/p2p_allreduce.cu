// p2p_allreduce.cu -- sum-allreduce of the flat gradient buffer over NVLink peer memory.
//
// Data-parallel training exchanges ONE dense fp32 buffer per step (hash-table + MLP gradients,
// 45.7 MB at T = 2^19).  Every rank maps every peer's buffer (CUDA IPC) and runs a two-shot
// allreduce with plain peer loads through NVLink 5 / NVSwitch:
//   barrier -> reduce-scatter (rank r sums slice r over all ranks, fixed rank order, and writes
//   it into its own buffer) -> barrier -> all-gather (rank r copies the reduced slices of its
//   peers) -> barrier.
// Each slice is reduced by exactly one rank in a fixed order, so all replicas end up with
// bit-identical gradients.  Barriers are flag arrays in peer memory (monotonic epochs, system-
// scope fences); every spin has a clock64() timeout that sets an error word instead of hanging.
#include "common.cuh"
#include <string.h>

#define VN_P2P_MAX_RANKS 8

struct P2PCtx {
    int rank, world;
    float* bufs[VN_P2P_MAX_RANKS];     // peer-mapped gradient buffers (bufs[rank] = local)
    int* flags[VN_P2P_MAX_RANKS];      // peer-mapped flag arrays, [world] ints each
    int* err;                          // local device int
    int epoch;
};
static P2PCtx g_ctx;
static bool g_ctx_ready = false;

__global__ void p2p_barrier_kernel(P2PCtx c, int value) {
    const int p = threadIdx.x;
    if (p < c.world) {
        __threadfence_system();
        volatile int* dst = c.flags[p] + c.rank;           // my arrival, in peer p's array
        *dst = value;
        volatile int* src = c.flags[c.rank] + p;           // peer p's arrival, in my array
        const long long t0 = clock64();
        while (*src < value) {
            if (clock64() - t0 > (long long)4e9) { *c.err = 1; break; }    // ~2 s: never hang the GPU
        }
        __threadfence_system();
    }
}

__global__ void __launch_bounds__(512) p2p_reduce_scatter_kernel(P2PCtx c, int64_t n4, int64_t chunk4) {
    const int64_t lo = (int64_t)c.rank * chunk4;
    const int64_t hi = min(lo + chunk4, n4);
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < VN_P2P_MAX_RANKS; ++p) {
            if (p < c.world) {
                const float4 v = __ldcv(reinterpret_cast<const float4*>(c.bufs[p]) + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<float4*>(c.bufs[c.rank])[i] = acc;
    }
}

__global__ void __launch_bounds__(512) p2p_all_gather_kernel(P2PCtx c, int64_t n4, int64_t chunk4) {
    // blockIdx.y walks the peers (skipping self)
    int p = blockIdx.y;
    if (p >= c.rank) ++p;
    const int64_t lo = (int64_t)p * chunk4;
    const int64_t hi = min(lo + chunk4, n4);
    const float4* src = reinterpret_cast<const float4*>(c.bufs[p]);
    float4* dst = reinterpret_cast<float4*>(c.bufs[c.rank]);
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __ldcv(src + i);
}

// export: the 64-byte cudaIpcMemHandle_t of the allocation that contains `ptr` and ptr's offset
// inside it (the caching allocator of the host framework sub-allocates from larger blocks)
VN_API int vn_ipc_get_handle(const void* ptr, void* h_handle64_out, int64_t* h_offset_out) {
    VN_REQUIRE(ptr && h_handle64_out && h_offset_out, "vn_ipc_get_handle: null argument");
    typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
    static GetRangeFn get_range = nullptr;
    if (!get_range) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        VN_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
        VN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "vn_ipc_get_handle: cuMemGetAddressRange unavailable");
        get_range = (GetRangeFn)fn;
    }
    unsigned long long base = 0;
    size_t size = 0;
    const int rc = get_range(&base, &size, (unsigned long long)(uintptr_t)ptr);
    VN_REQUIRE(rc == 0, "vn_ipc_get_handle: cuMemGetAddressRange failed (%d)", rc);
    cudaIpcMemHandle_t h;
    VN_CUDA(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
    memcpy(h_handle64_out, &h, sizeof(h));
    *h_offset_out = (int64_t)((unsigned long long)(uintptr_t)ptr - base);
    return VN_OK;
}

VN_API int vn_ipc_open(const void* h_handle64, int64_t offset_bytes, void** h_ptr_out) {
    VN_REQUIRE(h_handle64 && h_ptr_out, "vn_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, sizeof(h));
    void* base = nullptr;
    VN_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *h_ptr_out = (char*)base + offset_bytes;
    return VN_OK;
}

VN_API int vn_p2p_init(int rank, int world, void* const* h_bufs, void* const* h_flags, int* err_dev) {
    VN_REQUIRE(world >= 1 && world <= VN_P2P_MAX_RANKS && rank >= 0 && rank < world, "vn_p2p_init: bad rank/world");
    VN_REQUIRE(h_bufs && h_flags && err_dev, "vn_p2p_init: null argument");
    g_ctx.rank = rank; g_ctx.world = world; g_ctx.err = err_dev; g_ctx.epoch = 0;
    for (int p = 0; p < world; ++p) {
        VN_REQUIRE(h_bufs[p] && h_flags[p], "vn_p2p_init: null peer pointer");
        VN_REQUIRE(vn_aligned(h_bufs[p], 16), "vn_p2p_init: buffers must be 16-byte aligned");
        g_ctx.bufs[p] = (float*)h_bufs[p];
        g_ctx.flags[p] = (int*)h_flags[p];
    }
    g_ctx_ready = true;
    return VN_OK;
}

VN_API int vn_p2p_allreduce(int64_t n, void* stream) {
    VN_REQUIRE(g_ctx_ready, "vn_p2p_allreduce: vn_p2p_init has not been called");
    VN_REQUIRE(n >= 0 && n % 4 == 0, "vn_p2p_allreduce: n must be a multiple of 4 floats");
    if (n == 0 || g_ctx.world == 1) return VN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n4 = n / 4;
    const int64_t chunk4 = (n4 + g_ctx.world - 1) / g_ctx.world;
    const int sms = vn_sm_count();
    const int e = g_ctx.epoch;
    g_ctx.epoch += 3;
    p2p_barrier_kernel<<<1, 32, 0, st>>>(g_ctx, e + 1);                      // every rank's gradients are complete
    VN_CHECK_LAUNCH("p2p_barrier_kernel");
    p2p_reduce_scatter_kernel<<<sms, 512, 0, st>>>(g_ctx, n4, chunk4);
    VN_CHECK_LAUNCH("p2p_reduce_scatter_kernel");
    p2p_barrier_kernel<<<1, 32, 0, st>>>(g_ctx, e + 2);                      // every slice is reduced
    VN_CHECK_LAUNCH("p2p_barrier_kernel");
    dim3 grid((unsigned)((sms + g_ctx.world - 2) / (g_ctx.world - 1)), (unsigned)(g_ctx.world - 1));
    p2p_all_gather_kernel<<<grid, 512, 0, st>>>(g_ctx, n4, chunk4);
    VN_CHECK_LAUNCH("p2p_all_gather_kernel");
    p2p_barrier_kernel<<<1, 32, 0, st>>>(g_ctx, e + 3);                      // nobody reads my buffer any more
    VN_CHECK_LAUNCH("p2p_barrier_kernel");
    return VN_OK;
}
