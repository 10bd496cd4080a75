// cs_api.cu — library-wide plumbing: error message, device selection, shared math.
#include "cs_common.cuh"

#include <stdarg.h>

static thread_local char g_err[1024] = "";

void cs_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *cs_last_error(void) { return g_err; }
extern "C" int cs_version(void) { return 100; }

extern "C" int cs_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cs_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return -1;
    }
    return n;
}

extern "C" int cs_set_device(int device)
{
    CS_CUDA(cudaSetDevice(device));
    return CS_OK;
}

int cs_num_sms()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}
