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
#include <stdlib.h>

#define VN_P2P_MAX_RANKS 8
#define VN_P2P_MBOX 8                  // floats per rank and mailbox slot

struct P2PCtx {
    int rank, world;
    float* bufs[VN_P2P_MAX_RANKS];     // peer-mapped gradient buffers (bufs[rank] = local)
    int* flags[VN_P2P_MAX_RANKS];      // peer-mapped flag arrays, [world] ints each
    float* pbufs[VN_P2P_MAX_RANKS];    // peer-mapped parameter buffers (fused optimiser), or null
    float* mbox[VN_P2P_MAX_RANKS];     // peer-mapped mailboxes, [2][world][VN_P2P_MBOX] floats each, or null
    float* mc_grad;                    // NVLink-multicast (NVLS) address of the gradient buffers of all ranks, or null
    float* mc_params;                  // ... of the parameter buffers
    int* err;                          // local device int
    int epoch;
    int small_ops;                     // number of mailbox exchanges so far (mailbox parity)
};
static P2PCtx g_ctx;
static bool g_ctx_ready = false;

// spin budget of every cross-rank wait, in SM clocks (VN_P2P_TIMEOUT_MS in the environment, default 20 s at ~2 GHz).
// A time-out sets *err (the engine reads it back on the step's host sync and raises) and makes the step a no-op
// (treated like an overflow: nothing is applied to incomplete data).
static long long timeout_from_env() {
    const char* e = getenv("VN_P2P_TIMEOUT_MS");
    double ms = e ? atof(e) : 20000.0;
    if (!(ms > 0.0)) ms = 20000.0;
    return (long long)(ms * 2.0e6);
}
static long long g_timeout_clocks = timeout_from_env();
// per-context device scratch of vn_p2p_step: {go epoch, done-CTA counter, global found_inf, unused}
static int* g_step_sync = nullptr;

__global__ void p2p_barrier_kernel(P2PCtx c, int value, long long timeout) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: may be pre-launched behind the kernel that produces the data
    const int p = threadIdx.x;
    if (p < c.world) {
        __threadfence_system();
        volatile int* dst = c.flags[p] + c.rank;           // my arrival, in peer p's array
        *dst = value;
        volatile int* src = c.flags[c.rank] + p;           // peer p's arrival, in my array
        const long long t0 = clock64();
        while (*src < value) {
            if (clock64() - t0 > timeout) { *c.err = 1; break; }    // never hang the GPU; the engine raises on *err
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


// flag barrier that also carries a small payload: rank r deposits n <= VN_P2P_MBOX floats in
// slot [parity][r] of every peer's mailbox before it signals; after the wait each rank combines
// the `world` slots of its own mailbox in rank order (sum, or max when `use_max`), so every
// rank obtains the bit-identical result.  Two mailbox parities: a slot can only be overwritten
// two exchanges later, i.e. after its reader has passed another barrier.
__global__ void p2p_exchange_kernel(P2PCtx c, int value, int parity, float* data, int n, int use_max, long long timeout) {
    vn_pdl_trigger(); vn_pdl_wait();          // PDL: `data` comes from the preceding kernel
    const int p = threadIdx.x;
    if (p < c.world) {
        float* slot = c.mbox[p] + ((size_t)parity * c.world + c.rank) * VN_P2P_MBOX;
        for (int k = 0; k < n; ++k) reinterpret_cast<volatile float*>(slot)[k] = data[k];
        __threadfence_system();
        volatile int* dst = c.flags[p] + c.rank;
        *dst = value;
        volatile int* src = c.flags[c.rank] + p;
        const long long t0 = clock64();
        while (*src < value) {
            if (clock64() - t0 > timeout) { *c.err = 1; break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x < n) {
        const volatile float* mine = c.mbox[c.rank] + (size_t)parity * c.world * VN_P2P_MBOX + threadIdx.x;
        float acc = use_max ? -INFINITY : 0.0f;
        for (int q = 0; q < c.world; ++q) {
            const float v = mine[(size_t)q * VN_P2P_MBOX];
            acc = use_max ? fmaxf(acc, v) : acc + v;
        }
        data[threadIdx.x] = acc;
    }
}

// fused optimiser, phase A: reduce-scatter of this rank's slice (fixed rank order -> the same
// sums the two-shot allreduce produces) into the local gradient buffer + GradScaler inf check
__global__ void __launch_bounds__(512) p2p_reduce_check_kernel(P2PCtx c, int64_t n4, int64_t chunk4, float* found_inf) {
    const int64_t lo = (int64_t)c.rank * chunk4;
    const int64_t hi = min(lo + chunk4, n4);
    bool bad = false;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < VN_P2P_MAX_RANKS; ++p) {
            if (p < c.world) {
                const float4 v = __ldcv(reinterpret_cast<const float4*>(c.bufs[p]) + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        bad |= !(isfinite(acc.x) && isfinite(acc.y) && isfinite(acc.z) && isfinite(acc.w));
        reinterpret_cast<float4*>(c.bufs[c.rank])[i] = acc;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found_inf = 1.0f;
}

// fused optimiser, phase B: Adam on this rank's slice (its m / v never leave the rank) and
// all-gather of the UPDATED PARAMETERS by pushing them into every replica over NVLink
__global__ void __launch_bounds__(512) p2p_adam_push_kernel(P2PCtx c, int64_t n4, int64_t chunk4, float* __restrict__ m,
                                                            float* __restrict__ v, AdamCfg cfg,
                                                            const float* __restrict__ found_inf,
                                                            const float* __restrict__ scale_dev) {
    if (*found_inf != 0.0f) return;                       // GradScaler.step skips the step on every rank alike
    cfg.inv_scale = 1.0f / *scale_dev;
    const int64_t lo = (int64_t)c.rank * chunk4;
    const int64_t hi = min(lo + chunk4, n4);
    const float4* g4 = reinterpret_cast<const float4*>(c.bufs[c.rank]);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        float4 P = reinterpret_cast<const float4*>(c.pbufs[c.rank])[i], M = m4[i], V = v4[i];
        const float4 G = g4[i];
        adam1(P.x, G.x, M.x, V.x, cfg); adam1(P.y, G.y, M.y, V.y, cfg);
        adam1(P.z, G.z, M.z, V.z, cfg); adam1(P.w, G.w, M.w, V.w, cfg);
        m4[i] = M; v4[i] = V;
#pragma unroll
        for (int q = 0; q < VN_P2P_MAX_RANKS; ++q)
            if (q < c.world) reinterpret_cast<float4*>(c.pbufs[q])[i] = P;
    }
}

__global__ void p2p_scaler_update_kernel(float* scale, int32_t* tracker, float* found_inf, float growth, float backoff,
                                         int interval) {
    if (*found_inf != 0.0f) { *scale = *scale * backoff; *tracker = 0; }
    else {
        const int t = *tracker + 1;
        if (t == interval) { const float grown = *scale * growth; if (isfinite(grown)) *scale = grown; *tracker = 0; }   // torch _amp_update_scale_
        else *tracker = t;
    }
    *found_inf = 0.0f;
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
    // ONE exchange context per process: a second initialisation would silently re-route the first owner's kernels
    VN_REQUIRE(!g_ctx_ready, "vn_p2p_init: the peer-memory exchange is already owned (vn_p2p_shutdown releases it)");
    memset(&g_ctx, 0, sizeof(g_ctx));
    g_ctx.rank = rank; g_ctx.world = world; g_ctx.err = err_dev; g_ctx.epoch = 0; g_ctx.small_ops = 0;
    for (int p = 0; p < world; ++p) {
        VN_REQUIRE(h_bufs[p] && h_flags[p], "vn_p2p_init: null peer pointer");
        VN_REQUIRE(vn_aligned(h_bufs[p], 16), "vn_p2p_init: buffers must be 16-byte aligned");
        g_ctx.bufs[p] = (float*)h_bufs[p];
        g_ctx.flags[p] = (int*)h_flags[p];
    }
    if (!g_step_sync) VN_CUDA(cudaMalloc(&g_step_sync, 4 * sizeof(int)));
    VN_CUDA(cudaMemset(g_step_sync, 0, 4 * sizeof(int)));
    g_ctx_ready = true;
    return VN_OK;
}

// NVLink-multicast addresses of the gradient / parameter buffers (one address that stands for the same offset in every
// rank's buffer; e.g. torch.distributed._symmetric_memory's multicast_ptr): vn_p2p_step then reduces and broadcasts
// through the switch.  NULL switches back to peer loads / stores.
VN_API int vn_p2p_set_multicast(void* mc_grad, void* mc_params) {
    VN_REQUIRE(g_ctx_ready, "vn_p2p_set_multicast: vn_p2p_init has not been called");
    VN_REQUIRE((mc_grad == nullptr) == (mc_params == nullptr), "vn_p2p_set_multicast: both addresses or none");
    VN_REQUIRE(vn_aligned(mc_grad, 16) && vn_aligned(mc_params, 16), "vn_p2p_set_multicast: addresses must be 16-byte aligned");
    g_ctx.mc_grad = (float*)mc_grad; g_ctx.mc_params = (float*)mc_params;
    return VN_OK;
}

VN_API int vn_p2p_shutdown(void) {
    g_ctx_ready = false;
    memset(&g_ctx, 0, sizeof(g_ctx));
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
    p2p_barrier_kernel<<<1, 32, 0, st>>>(g_ctx, e + 1, g_timeout_clocks);                      // every rank's gradients are complete
    VN_CHECK_LAUNCH("p2p_barrier_kernel");
    p2p_reduce_scatter_kernel<<<sms, 512, 0, st>>>(g_ctx, n4, chunk4);
    VN_CHECK_LAUNCH("p2p_reduce_scatter_kernel");
    p2p_barrier_kernel<<<1, 32, 0, st>>>(g_ctx, e + 2, g_timeout_clocks);                      // every slice is reduced
    VN_CHECK_LAUNCH("p2p_barrier_kernel");
    dim3 grid((unsigned)((sms + g_ctx.world - 2) / (g_ctx.world - 1)), (unsigned)(g_ctx.world - 1));
    p2p_all_gather_kernel<<<grid, 512, 0, st>>>(g_ctx, n4, chunk4);
    VN_CHECK_LAUNCH("p2p_all_gather_kernel");
    p2p_barrier_kernel<<<1, 32, 0, st>>>(g_ctx, e + 3, g_timeout_clocks);                      // nobody reads my buffer any more
    VN_CHECK_LAUNCH("p2p_barrier_kernel");
    return VN_OK;
}

VN_API int vn_p2p_attach(void* const* h_pbufs, void* const* h_mbox) {
    VN_REQUIRE(g_ctx_ready, "vn_p2p_attach: vn_p2p_init has not been called");
    VN_REQUIRE(h_mbox != nullptr, "vn_p2p_attach: null mailbox list");
    for (int p = 0; p < g_ctx.world; ++p) {
        VN_REQUIRE(h_mbox[p] && vn_aligned(h_mbox[p], 16), "vn_p2p_attach: null / misaligned mailbox");
        g_ctx.mbox[p] = (float*)h_mbox[p];
        if (h_pbufs) {
            VN_REQUIRE(h_pbufs[p] && vn_aligned(h_pbufs[p], 16), "vn_p2p_attach: null / misaligned parameter buffer");
            g_ctx.pbufs[p] = (float*)h_pbufs[p];
        }
    }
    return VN_OK;
}

VN_API int vn_p2p_allreduce_small(float* data, int n, int use_max, void* stream) {
    VN_REQUIRE(g_ctx_ready && g_ctx.mbox[0], "vn_p2p_allreduce_small: vn_p2p_init / vn_p2p_attach have not been called");
    VN_REQUIRE(data && n >= 1 && n <= VN_P2P_MBOX, "vn_p2p_allreduce_small: n must be in [1,%d]", VN_P2P_MBOX);
    if (g_ctx.world == 1) return VN_OK;
    const int e = ++g_ctx.epoch;
    const int parity = (g_ctx.small_ops++) & 1;
    vn_launch_pdl(p2p_exchange_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, g_ctx, e, parity, data, n, use_max, g_timeout_clocks);
    VN_CHECK_LAUNCH("p2p_exchange_kernel");
    return VN_OK;
}

VN_API int vn_p2p_reduce_adam(int64_t n, float* m, float* v, double lr, double beta1, double beta2, double eps, int step,
                              float* found_inf, float* scale_dev, int32_t* growth_tracker, void* stream) {
    VN_REQUIRE(g_ctx_ready && g_ctx.mbox[0] && g_ctx.pbufs[0],
               "vn_p2p_reduce_adam: vn_p2p_init / vn_p2p_attach (with parameter buffers) have not been called");
    VN_REQUIRE(n >= 0 && n % 4 == 0 && step >= 1, "vn_p2p_reduce_adam: n must be a multiple of 4 floats, step >= 1");
    VN_REQUIRE(m && v && found_inf && scale_dev && growth_tracker, "vn_p2p_reduce_adam: null pointer");
    VN_REQUIRE(vn_aligned(m, 16) && vn_aligned(v, 16), "vn_p2p_reduce_adam: m / v must be 16-byte aligned");
    if (n == 0) return VN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n4 = n / 4;
    const int64_t chunk4 = (n4 + g_ctx.world - 1) / g_ctx.world;
    const int sms = vn_sm_count();
    const AdamCfg c = vn_make_adam_cfg(lr, beta1, beta2, eps, step, 1.0f);
    const int e = g_ctx.epoch;
    g_ctx.epoch += 3;
    const int parity = (g_ctx.small_ops++) & 1;
    vn_launch_pdl(p2p_barrier_kernel, dim3(1), dim3(32), 0, st, g_ctx, e + 1, g_timeout_clocks);                      // every rank's gradients are complete
    VN_CHECK_LAUNCH("p2p_barrier_kernel");
    p2p_reduce_check_kernel<<<sms, 512, 0, st>>>(g_ctx, n4, chunk4, found_inf);
    VN_CHECK_LAUNCH("p2p_reduce_check_kernel");
    // every slice is reduced (nobody reads a peer's gradient after this) + global OR of the inf flags
    p2p_exchange_kernel<<<1, 32, 0, st>>>(g_ctx, e + 2, parity, found_inf, 1, 1, g_timeout_clocks);
    VN_CHECK_LAUNCH("p2p_exchange_kernel");
    p2p_adam_push_kernel<<<sms, 512, 0, st>>>(g_ctx, n4, chunk4, m, v, c, found_inf, scale_dev);
    VN_CHECK_LAUNCH("p2p_adam_push_kernel");
    p2p_barrier_kernel<<<1, 32, 0, st>>>(g_ctx, e + 3, g_timeout_clocks);                      // every replica holds every updated slice
    VN_CHECK_LAUNCH("p2p_barrier_kernel");
    p2p_scaler_update_kernel<<<1, 1, 0, st>>>(scale_dev, growth_tracker, found_inf, 2.0f, 0.5f, 2000);
    VN_CHECK_LAUNCH("p2p_scaler_update_kernel");
    return VN_OK;
}


// ---- the sharded optimiser as ONE kernel -----------------------------------------------------------------------------
// vn_p2p_reduce_adam above is six launches with three cross-rank barriers, its inbound reduce and its outbound parameter
// push never overlap, and its inf check needs a second exchange in the middle.  Here the inf flag is an INPUT (the fused
// backward kernel -- or vn_grad_check -- has evaluated it on the rank's own gradient: every contribution is bounded, so a
// sum over <= 8 ranks of finite gradients is finite) and rides in the start barrier's mailbox, which leaves one pass:
//   start barrier (+ max of the inf flags) -> for every float4 of this rank's slice: sum over the ranks' gradients
//   (fixed rank order, peer loads) -> Adam with rank-local m / v -> store the new parameters into EVERY replica (peer
//   stores) -> end barrier -> GradScaler / step-count update.
// Loads of later elements are in flight while the stores of earlier ones drain, so both NVLink directions are busy for
// the whole pass (a rank also serves its peers' loads while it loads, so either direction carries 2 (n-1)/n of the buffer
// per step: 80 MB at 8 ranks, i.e. >= 0.09 ms at the 900 GB/s of NVLink 5; measured 0.172 ms, NCCL's allreduce of the same
// buffer alone 0.209 ms).  CTA 0 runs the start barrier and releases the others through a device-scope flag; the LAST CTA to
// finish (atomic counter) runs the end barrier and the scalar update, so the kernel's completion == "every replica
// holds every updated slice and nobody reads my gradient buffer any more".
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// NVLS: the sum over the ranks is ONE multimem.ld_reduce per float4 -- the NVSwitch adds the replicas' values and returns
// the result, so 16 bytes cross this GPU's link instead of 16 (n - 1) -- and the new parameters reach every replica with
// ONE multimem.st that the switch replicates.  Per step and direction a GPU then moves ~2 x 45.7 MB / n instead of
// 2 x 45.7 MB (n - 1) / n.  The order of the additions inside the switch is not the fixed rank order of the peer-load
// path: the replicas stay bit-identical to EACH OTHER (every slice is reduced once, by its owner), but agree with
// allreduce + dense Adam only to rounding.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

template <bool NVLS>
__global__ void __launch_bounds__(512) p2p_step_kernel(P2PCtx c, int64_t n4, int64_t chunk4, float* __restrict__ m,
                                                       float* __restrict__ v, AdamCfg cfg, float* found_inf, float* scale_dev,
                                                       int32_t* tracker, float* opt_state, double lr, double beta1, double beta2,
                                                       int epoch, int parity, int* sync, long long timeout) {
    vn_pdl_trigger(); vn_pdl_wait();          // this rank's gradients (and its inf flag) are complete
    __shared__ float s_found;
    __shared__ int s_last;
    int* go = sync; int* done = sync + 1; float* gfound = reinterpret_cast<float*>(sync + 2);
    if (blockIdx.x == 0) {
        // ---- start barrier + max of the inf flags (mailbox slot [parity][rank] of every peer)
        const int p = threadIdx.x;
        if (p < c.world) {
            float* slot = c.mbox[p] + ((size_t)parity * c.world + c.rank) * VN_P2P_MBOX;
            *reinterpret_cast<volatile float*>(slot) = *found_inf;
            __threadfence_system();
            *reinterpret_cast<volatile int*>(c.flags[p] + c.rank) = epoch + 1;
            volatile int* src = c.flags[c.rank] + p;
            const long long t0 = clock64();
            while (*src < epoch + 1) {
                if (clock64() - t0 > timeout) { *c.err = 1; break; }
            }
            __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float f = 0.0f;
            const volatile float* mine = c.mbox[c.rank] + (size_t)parity * c.world * VN_P2P_MBOX;
            for (int q = 0; q < c.world; ++q) f = fmaxf(f, mine[(size_t)q * VN_P2P_MBOX]);
            if (*reinterpret_cast<volatile int*>(c.err) != 0) f = 1.0f;        // timed out: apply nothing to incomplete data
            *gfound = f;
            __threadfence();
            st_release_gpu(go, epoch + 1);
        }
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_acquire_gpu(go) < epoch + 1) {
            if (clock64() - t0 > 2 * timeout) break;                            // CTA 0 has set *err already
        }
        s_found = *reinterpret_cast<volatile float*>(gfound);
    }
    __syncthreads();
    const bool skip = s_found != 0.0f;
    if (!skip) {
        cfg.inv_scale = 1.0f / *scale_dev;
        if (opt_state) { cfg.step_size = opt_state[0]; cfg.bc2_sqrt = opt_state[1]; }
        const int64_t lo = (int64_t)c.rank * chunk4;
        const int64_t hi = min(lo + chunk4, n4);
        float4* m4 = reinterpret_cast<float4*>(m);
        float4* v4 = reinterpret_cast<float4*>(v);
        // (Measured dead end, profiles/r2_dp.md: four elements per thread with all loads / switch reductions issued before
        // the first use -- 0.141 -> 0.164 ms through NVLS at 8 ranks, 0.102 -> 0.106 ms with peer loads at 2.)
        for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
            float4 G = make_float4(0.f, 0.f, 0.f, 0.f);
            if (NVLS) {
                G = multimem_ld_reduce_add(reinterpret_cast<const float4*>(c.mc_grad) + i);
            } else {
#pragma unroll
                for (int p = 0; p < VN_P2P_MAX_RANKS; ++p) {
                    if (p < c.world) {
                        const float4 g = __ldcv(reinterpret_cast<const float4*>(c.bufs[p]) + i);
                        G.x += g.x; G.y += g.y; G.z += g.z; G.w += g.w;
                    }
                }
            }
            float4 P = reinterpret_cast<const float4*>(c.pbufs[c.rank])[i], M = m4[i], V = v4[i];
            adam1(P.x, G.x, M.x, V.x, cfg); adam1(P.y, G.y, M.y, V.y, cfg);
            adam1(P.z, G.z, M.z, V.z, cfg); adam1(P.w, G.w, M.w, V.w, cfg);
            m4[i] = M; v4[i] = V;
            if (NVLS) {
                multimem_st(reinterpret_cast<float4*>(c.mc_params) + i, P);
            } else {
#pragma unroll
                for (int q = 0; q < VN_P2P_MAX_RANKS; ++q)
                    if (q < c.world) reinterpret_cast<float4*>(c.pbufs[q])[i] = P;
            }
        }
    }
    // ---- the last CTA to get here closes the step
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(done, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int p = threadIdx.x;
    if (p < c.world) {
        __threadfence_system();
        *reinterpret_cast<volatile int*>(c.flags[p] + c.rank) = epoch + 2;
        volatile int* src = c.flags[c.rank] + p;
        const long long t0 = clock64();
        while (*src < epoch + 2) {
            if (clock64() - t0 > timeout) { *c.err = 1; break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *done = 0;
        // GradScaler.update + Adam's step count (vn_scaler_update_dev)
        if (skip) { *scale_dev = *scale_dev * 0.5f; *tracker = 0; }
        else {
            if (opt_state) vn_opt_state_set(opt_state, __float_as_int(opt_state[2]) + 1, lr, beta1, beta2);
            const int t = *tracker + 1;
            if (t == 2000) { const float grown = *scale_dev * 2.0f; if (isfinite(grown)) *scale_dev = grown; *tracker = 0; }
            else *tracker = t;
        }
        *found_inf = 0.0f;
    }
}

VN_API int vn_p2p_step(int64_t n, float* m, float* v, double lr, double beta1, double beta2, double eps, float* opt_state,
                       float* found_inf, float* scale_dev, int32_t* growth_tracker, void* stream) {
    VN_REQUIRE(g_ctx_ready && g_ctx.mbox[0] && g_ctx.pbufs[0],
               "vn_p2p_step: vn_p2p_init / vn_p2p_attach (with parameter buffers) have not been called");
    VN_REQUIRE(n >= 0 && n % 4 == 0, "vn_p2p_step: n must be a multiple of 4 floats");
    VN_REQUIRE(m && v && opt_state && found_inf && scale_dev && growth_tracker, "vn_p2p_step: null pointer");
    VN_REQUIRE(vn_aligned(m, 16) && vn_aligned(v, 16), "vn_p2p_step: m / v must be 16-byte aligned");
    if (n == 0) return VN_OK;
    const int64_t n4 = n / 4;
    const int64_t chunk4 = (n4 + g_ctx.world - 1) / g_ctx.world;
    const AdamCfg c = vn_make_adam_cfg(lr, beta1, beta2, eps, 1, 1.0f);      // step-dependent fields come from opt_state
    const int e = g_ctx.epoch;
    g_ctx.epoch += 2;
    const int parity = (g_ctx.small_ops++) & 1;
    // all CTAs must be co-resident (grid-wide flags): one per SM; half of them reduce, half push (VN_P2P_STEP_SPLIT=0: every
    // thread does both)
    // all CTAs must be co-resident (grid-wide flags): one per SM.  (Measured at 2 and 8 ranks, profiles/r2_dp.md: splitting
    // the CTAs into reducers and pushers, or running "all loads, grid barrier, all stores", is not faster -- both NVLink
    // directions are busy in either phase, because a rank serves its peers' loads while it loads.)
    if (g_ctx.mc_grad && g_ctx.mc_params)
        vn_launch_pdl(p2p_step_kernel<true>, dim3((unsigned)vn_sm_count()), dim3(512), 0, (cudaStream_t)stream, g_ctx, n4, chunk4, m, v,
                      c, found_inf, scale_dev, growth_tracker, opt_state, lr, beta1, beta2, e, parity, g_step_sync, g_timeout_clocks);
    else
        vn_launch_pdl(p2p_step_kernel<false>, dim3((unsigned)vn_sm_count()), dim3(512), 0, (cudaStream_t)stream, g_ctx, n4, chunk4, m, v,
                      c, found_inf, scale_dev, growth_tracker, opt_state, lr, beta1, beta2, e, parity, g_step_sync, g_timeout_clocks);
    VN_CHECK_LAUNCH("p2p_step_kernel");
    return VN_OK;
}
