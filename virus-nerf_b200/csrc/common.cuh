// common.cuh -- shared host/device helpers for libvirusnerf_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/virusnerf.h"

#define VN_API extern "C" __attribute__((visibility("default")))

// ---- error plumbing (thread-local message, negative return codes, no exceptions) ----
void vn_set_error(const char* fmt, ...);

#define VN_REQUIRE(cond, ...)                       \
    do {                                            \
        if (!(cond)) {                              \
            vn_set_error(__VA_ARGS__);              \
            return VN_EINVAL;                       \
        }                                           \
    } while (0)

extern unsigned long long g_vn_launches;   // kernels launched by this library (api.cu)

#define VN_CHECK_LAUNCH(name)                                                        \
    do {                                                                             \
        ++g_vn_launches;                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            vn_set_error("%s: CUDA error %d (%s)", name, (int)e__, cudaGetErrorString(e__)); \
            return VN_ELAUNCH;                                                       \
        }                                                                            \
    } while (0)

#define VN_CUDA(call)                                                                \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            vn_set_error("%s: CUDA error %d (%s)", #call, (int)e__, cudaGetErrorString(e__)); \
            return VN_ELAUNCH;                                                       \
        }                                                                            \
    } while (0)

// ---- optional in-stream kernel timing (CUDA events around a launch; see vn_profile_*) ----
enum { VN_K_HASH_FWD = 0, VN_K_HASH_BWD, VN_K_MLP_FWD, VN_K_MLP_BWD, VN_K_MARCH_COUNT, VN_K_MARCH_WRITE, VN_K_COMP_FWD,
       VN_K_COMP_BWD, VN_K_ADAM, VN_K_MLP_BWD_SCATTER, VN_K_COUNT };
extern unsigned g_vn_profiling;           // bit k set: kernel id k is timed
void vn_prof_begin(int kernel_id, int64_t size, cudaStream_t st);
void vn_prof_end(cudaStream_t st);
struct VnProfScope {
    cudaStream_t st; bool on;
    VnProfScope(int id, int64_t size, cudaStream_t s) : st(s), on((g_vn_profiling >> id) & 1u) { if (on) vn_prof_begin(id, size, st); }
    ~VnProfScope() { if (on) vn_prof_end(st); }
};

// ---- programmatic dependent launch (PDL): the kernels of the train step are launched with
// programmaticStreamSerialization so that the launch latency and the prologue of kernel k+1
// (index math, weight staging, TMEM allocation) overlap the tail of kernel k.  Every such kernel
// executes vn_pdl_wait() before it touches memory a predecessor may still be writing (or reading)
// and vn_pdl_trigger() as early as possible.  VN_PDL=0 in the environment turns the attribute off.
// RULE: data that a predecessor kernel of the stream produces is read with ld.global.cg (__ldcg) or plain loads, never
// with the non-coherent ld.global.nc (__ldg): an invariant load is not ordered by the asm memory clobber of
// griddepcontrol.wait and may be scheduled above it (observed in round 2: profiles/r2_kbench.md).  __ldg stays for data
// that is older than the previous kernel (tables after the optimiser step, rays, positions, measurement targets).
extern bool g_vn_pdl;
template <typename... KArgs, typename... Args>
static inline void vn_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_vn_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through cudaGetLastError()
}
#ifdef __CUDACC__
__device__ __forceinline__ void vn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void vn_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

static inline bool vn_aligned(const void* p, size_t a) { return ((uintptr_t)p % a) == 0; }
static inline unsigned vn_blocks(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }
int vn_sm_count();

// ---- constants, modules/utils.py:12-16 ----
#define VN_NEAR_DISTANCE 0.01f
#define VN_SQRT3_MAX_SAMPLES ((float)(1.7320508075688772 / 1024.0))
#define VN_SQRT3_2 ((float)(1.7320508075688772 * 2.0))

// ---- device helpers.  Index-critical float expressions use __fmul_rn/__fadd_rn so nvcc
// cannot contract them into FMAs: the oracle is defined without contraction. ----
__device__ __forceinline__ float vn_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float vn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float vn_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float vn_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float vn_clamp(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

// modules/utils.py:54-57
__device__ __forceinline__ float vn_calc_dt(float t, float esf, float dt_max) {
    return vn_clamp(vn_mul(t, esf), VN_SQRT3_MAX_SAMPLES, dt_max);
}
// modules/utils.py:60-75
__device__ __forceinline__ int vn_frexp_bit(float x) {
    int exponent = 0;
    if (x != 0.0f) {
        uint32_t bits = __float_as_uint(x);
        exponent = (int)((bits & 0x7f800000u) >> 23) - 127;
        // frac = 1.mantissa in [1,2): "frac > 1" <=> mantissa != 0 ("frac < 0.5" never holds)
        if ((bits & 0x7fffffu) != 0u) exponent += 1;
    }
    return exponent;
}
// modules/utils.py:95-117
__device__ __forceinline__ uint32_t vn_expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t vn_morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return vn_expand_bits(x) | (vn_expand_bits(y) << 1) | (vn_expand_bits(z) << 2);
}
__device__ __forceinline__ uint32_t vn_morton3D_invert(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}
// float -> u32, saturating (negative / NaN -> 0): PTX cvt.rzi.u32.f32 semantics
__device__ __forceinline__ uint32_t vn_f2u(float x) { return __float2uint_rz(x); }

__device__ __forceinline__ void vn_red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void vn_red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// ---- Adam update of one element (torch.optim.Adam single-tensor path + GradScaler.unscale_),
// shared by the single-GPU optimiser (optim.cu) and the sharded one fused with the DP exchange
// (p2p_allreduce.cu) so that both produce the same bits.  Explicit fma / mul / add: no
// context-dependent contraction.
struct AdamCfg { float inv_scale, beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt; };

// the f32 constants torch's single-tensor Adam hands to its kernels: python-double arithmetic on the
// hyper-parameters first, ONE rounding to f32 when the scalar meets the f32 tensor -- in particular
// 1 - beta2 = f32(1 - 0.999) = 0.001f, not 1.0f - 0.999f (which is 1.3e-5 smaller)
#include <math.h>
static inline AdamCfg vn_make_adam_cfg(double lr, double beta1, double beta2, double eps, int step, float inv_scale) {
    AdamCfg c;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    c.inv_scale = inv_scale;
    c.beta1 = (float)beta1; c.beta2 = (float)beta2;                 // exp_avg_sq.mul_(beta2)
    c.omb1 = (float)(1.0 - beta1); c.omb2 = (float)(1.0 - beta2);   // lerp_(grad, 1 - beta1); addcmul_(..., value=1 - beta2)
    c.eps = (float)eps;
    c.step_size = (float)(lr / bc1);                                // addcdiv_(..., value=-step_size)
    c.bc2_sqrt = (float)sqrt(bc2);                                  // exp_avg_sq.sqrt() / bias_correction2_sqrt
    return c;
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, const AdamCfg& c) {
    g = __fmul_rn(g, c.inv_scale);                                   // GradScaler.unscale_
    m = __fmaf_rn(__fsub_rn(g, m), c.omb1, m);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = __fmaf_rn(__fmul_rn(c.omb2, g), g, __fmul_rn(v, c.beta2));   // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), c.bc2_sqrt), c.eps);   // (sqrt / bias_correction2_sqrt).add_(eps)
    p = __fsub_rn(p, __fmul_rn(c.step_size, __fdiv_rn(m, denom)));   // param.addcdiv_(exp_avg, denom, -step_size)
}

// Adam's step count on the device (optim.cu "optimiser step count on the device"): opt_state = {lr / (1 - beta1^t),
// sqrt(1 - beta2^t) for the NEXT step t, bit pattern of the int32 count of applied steps, unused}
#ifdef __CUDACC__
__device__ __forceinline__ void vn_opt_state_set(float* opt_state, int applied, double lr, double beta1, double beta2) {
    const double t = (double)(applied + 1);
    opt_state[0] = (float)(lr / (1.0 - pow(beta1, t)));
    opt_state[1] = (float)sqrt(1.0 - pow(beta2, t));
    opt_state[2] = __int_as_float(applied);
}
#endif
