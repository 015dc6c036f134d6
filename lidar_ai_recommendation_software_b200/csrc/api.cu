// Error plumbing and device queries of the lidar_b200 C ABI (include/lidar_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace lidar {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    return LIDAR_ERR_CUDA;
}

struct DevInfo {
    int sms = 0;
    size_t smem = 0;
    bool valid = false;
};
static DevInfo g_dev[64];

static DevInfo& dev_info() {
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 64) d = 0;
    DevInfo& di = g_dev[d];
    if (!di.valid) {
        int sms = 0, smem = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d);
        cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, d);
        di.sms = sms > 0 ? sms : 148;
        di.smem = smem > 0 ? (size_t)smem : (size_t)48 * 1024;
        di.valid = true;
    }
    return di;
}

int sm_count() { return dev_info().sms; }
size_t smem_optin() { return dev_info().smem; }

}  // namespace lidar

extern "C" {

const char* lidar_last_error(void) { return lidar::g_err; }

int lidar_abi_version(void) { return LIDAR_ABI_VERSION; }

int lidar_device_props(int device, int* sm_count, int* cc, size_t* smem_optin) {
    cudaDeviceProp p;
    LIDAR_CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc) *cc = p.major * 10 + p.minor;
    if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
    return LIDAR_OK;
}

}  // extern "C"
