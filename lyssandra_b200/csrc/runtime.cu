// runtime.cu — error string, version, device probing.
#include "common.cuh"
#include <string.h>

namespace lys {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

}  // namespace lys

extern "C" int lys_version(void) { return 100; }   // 0.1.0

extern "C" const char* lys_last_error(void) { return lys::g_err; }

// fingerprint of the sources this library was built from (set by lyssandra_b200/_build.py; the loader compares it with
// the tree it runs in, so a stale library is rebuilt instead of being loaded with mismatched signatures)
#ifndef LYS_FINGERPRINT
#define LYS_FINGERPRINT "unknown"
#endif
static const char g_fingerprint[] = "LYSFP:" LYS_FINGERPRINT;      // the marker lets the loader find it without dlopen
extern "C" const char* lys_build_fingerprint(void) { return g_fingerprint + 6; }

extern "C" int lys_device_info(int device, int* sm_count, int* cc_major, int* cc_minor)
{
    int n = 0;
    LYS_CUDA(cudaGetDeviceCount(&n));
    LYS_CHECK_ARG(device >= 0 && device < n, "lys_device_info: device %d out of range (count %d)", device, n);
    int sms = 0, maj = 0, min = 0;
    LYS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    LYS_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, device));
    LYS_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, device));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (maj != 10) {
        lys::set_error("lys_device_info: device %d is sm_%d%d; this library is built for sm_100a only", device, maj, min);
        return LYS_EUNSUPPORTED;
    }
    return LYS_OK;
}
