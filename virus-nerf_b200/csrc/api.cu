// api.cu -- error plumbing and device queries of the C ABI (include/virusnerf.h).
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";
unsigned long long g_vn_launches = 0;

void vn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int vn_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

VN_API const char* vn_last_error(void) { return g_err; }
VN_API int vn_abi_version(void) { return VN_ABI_VERSION; }
VN_API int64_t vn_launch_count(void) { return (int64_t)g_vn_launches; }

VN_API int vn_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
    int dev = 0;
    VN_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    VN_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_len > 0) { strncpy(name, prop.name, (size_t)name_len - 1); name[name_len - 1] = 0; }
    return VN_OK;
}

// ---- in-stream kernel timing -----------------------------------------------------------------
unsigned g_vn_profiling = 0u;
static bool pdl_from_env() { const char* e = getenv("VN_PDL"); return !(e && e[0] == '0'); }
bool g_vn_pdl = pdl_from_env();
namespace {
struct ProfRec { int id; int64_t size; cudaEvent_t e0, e1; };
const int kMaxRecs = 8192;
ProfRec g_recs[kMaxRecs];
int g_nrecs = 0, g_nevents = 0;
}

void vn_prof_begin(int kernel_id, int64_t size, cudaStream_t st) {
    if (g_nrecs >= kMaxRecs) { g_vn_profiling = 0u; return; }
    ProfRec& r = g_recs[g_nrecs];
    if (g_nrecs >= g_nevents) { cudaEventCreate(&r.e0); cudaEventCreate(&r.e1); ++g_nevents; }
    r.id = kernel_id; r.size = size;
    cudaEventRecord(r.e0, st);
}
void vn_prof_end(cudaStream_t st) {
    if (g_nrecs >= kMaxRecs) return;
    cudaEventRecord(g_recs[g_nrecs].e1, st);
    ++g_nrecs;
}

VN_API int vn_profile_enable(int on) { g_vn_profiling = on ? 0xffffffffu : 0u; if (on) g_nrecs = 0; return VN_OK; }
VN_API int vn_profile_enable_mask(unsigned mask) { g_vn_profiling = mask; if (mask) g_nrecs = 0; return VN_OK; }
VN_API int vn_set_pdl(int on) { g_vn_pdl = on != 0; return VN_OK; }
VN_API int vn_profile_count(void) { return g_nrecs; }
VN_API int vn_profile_get(int i, int* kernel_id, int64_t* size, float* ms) {
    VN_REQUIRE(i >= 0 && i < g_nrecs && kernel_id && size && ms, "vn_profile_get: bad index");
    VN_CUDA(cudaEventSynchronize(g_recs[i].e1));
    VN_CUDA(cudaEventElapsedTime(ms, g_recs[i].e0, g_recs[i].e1));
    *kernel_id = g_recs[i].id; *size = g_recs[i].size;
    return VN_OK;
}
