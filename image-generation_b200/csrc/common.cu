// Error plumbing and device checks shared by every entry point of libb200grbm.so.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

namespace b200grbm {

static thread_local char g_error[512] = "";

char *error_buffer() { return g_error; }

int32_t fail(int32_t code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int32_t check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    snprintf(g_error, sizeof(g_error), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    // clear the sticky-less error so the next call starts clean
    (void)cudaGetLastError();
    return (int32_t)e;
}

struct DeviceInfo {
    int device = -1;
    int sm = 0, major = 0, minor = 0, smem_optin = 0;
};

static int32_t query(DeviceInfo &d)
{
    static thread_local DeviceInfo cache;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(B200GRBM_ENODEVICE, "no CUDA device available (%s); libb200grbm has no CPU fallback",
                    cudaGetErrorString(e));
    }
    if (cache.device != dev) {
        DeviceInfo t;
        t.device = dev;
        B200_CUDA(cudaDeviceGetAttribute(&t.sm, cudaDevAttrMultiProcessorCount, dev));
        B200_CUDA(cudaDeviceGetAttribute(&t.major, cudaDevAttrComputeCapabilityMajor, dev));
        B200_CUDA(cudaDeviceGetAttribute(&t.minor, cudaDevAttrComputeCapabilityMinor, dev));
        B200_CUDA(cudaDeviceGetAttribute(&t.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cache = t;
    }
    d = cache;
    return 0;
}

int32_t require_device()
{
    DeviceInfo d;
    B200_TRY(query(d));
    if (d.major != 10)
        return fail(B200GRBM_ENODEVICE, "device %d is sm_%d%d; libb200grbm is built for sm_100a only", d.device,
                    d.major, d.minor);
    return 0;
}

int32_t sm_count()
{
    DeviceInfo d;
    if (query(d) != 0) return 0;
    return d.sm;
}

}  // namespace b200grbm

using namespace b200grbm;

extern "C" const char *b200grbm_last_error(void) { return error_buffer(); }

extern "C" int32_t b200grbm_abi_version(void) { return B200GRBM_ABI_VERSION; }

extern "C" int32_t b200grbm_device_info(int32_t *sm, int32_t *major, int32_t *minor, int32_t *smem_optin)
{
    DeviceInfo d;
    B200_TRY(query(d));
    if (sm) *sm = d.sm;
    if (major) *major = d.major;
    if (minor) *minor = d.minor;
    if (smem_optin) *smem_optin = d.smem_optin;
    return 0;
}
