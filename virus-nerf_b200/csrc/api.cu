// api.cu -- error plumbing and device queries of the C ABI (include/virusnerf.h).
#include "common.cuh"
#include <string.h>

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
