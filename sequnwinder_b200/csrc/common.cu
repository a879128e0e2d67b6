// common.cu — error state and device checks for libsequnwinder_b200.so.
#include "common.cuh"

#include <cstring>

namespace sqw {

char *last_error_buf()
{
    static thread_local char buf[512] = {0};
    return buf;
}

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return set_error(SQW_ERR_NO_DEVICE,
                         "no CUDA device available (libsequnwinder_b200 has no CPU fallback): %s",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return SQW_OK;
}

}  // namespace sqw

extern "C" {

int sqw_abi_version(void) { return SQW_ABI_VERSION; }

const char *sqw_last_error(void) { return sqw::last_error_buf(); }

int sqw_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int sqw_set_device(int device)
{
    int rc = sqw::require_device();
    if (rc != SQW_OK)
        return rc;
    SQW_CUDA_CHECK(cudaSetDevice(device));
    return SQW_OK;
}

int sqw_get_device(void)
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return dev;
}

}  // extern "C"
