// capi.cu -- runtime half of the C ABI: error reporting, device selection, memory helpers.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cuda_runtime.h>

#include "common.h"
#include "iir_internal.h"

namespace sdsp_b200
{
static thread_local char g_error[512] = "";

int set_error(int status, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return status;
}

int cuda_fail(int cuda_error, const char *what, const char *file, int line)
{
    const cudaError_t e = (cudaError_t)cuda_error;
    cudaGetLastError(); // clear the sticky-free error state
    const char *slash = strrchr(file, '/');
    int status = SDSP_B200_ERR_CUDA;
    if (e == cudaErrorMemoryAllocation)
        status = SDSP_B200_ERR_OOM;
    else if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
        status = SDSP_B200_ERR_NO_DEVICE;
    return set_error(status, "CUDA error %d (%s) in %s at %s:%d", cuda_error, cudaGetErrorString(e), what, slash ? slash + 1 : file, line);
}

static std::mutex g_dev_mu;
static int g_dev_checked[64];
static int g_dev_sms[64];

int ensure_device(int device)
{
    if (device < 0 || device >= 64)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "bad device index %d", device);
    std::lock_guard<std::mutex> lock(g_dev_mu);
    if (!g_dev_checked[device]) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) {
            cudaGetLastError();
            return set_error(SDSP_B200_ERR_NO_DEVICE, "no CUDA device available (%s); libsdsp_b200 has no CPU fallback",
                             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        }
        if (device >= count)
            return set_error(SDSP_B200_ERR_NO_DEVICE, "device %d requested but only %d present", device, count);
        cudaDeviceProp prop;
        SDSP_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            return set_error(SDSP_B200_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this library carries sm_100a code only", device, prop.name,
                             prop.major, prop.minor);
        g_dev_sms[device] = prop.multiProcessorCount;
        g_dev_checked[device] = 1;
    }
    SDSP_CUDA(cudaSetDevice(device));
    return SDSP_B200_OK;
}

int device_sm_count(int device)
{
    return (device >= 0 && device < 64 && g_dev_sms[device] > 0) ? g_dev_sms[device] : 148;
}
} // namespace sdsp_b200

using namespace sdsp_b200;

extern "C" {

int sdsp_b200_version(void)
{
    return SDSP_B200_VERSION;
}

const char *sdsp_b200_last_error(void)
{
    return g_error;
}

int sdsp_b200_device_count(int *count)
{
    if (!count)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "device_count: null out pointer");
    *count = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(SDSP_B200_ERR_NO_DEVICE, "no CUDA device available (%s)", cudaGetErrorString(e));
    }
    *count = n;
    return SDSP_B200_OK;
}

int sdsp_b200_init(int device)
{
    int rc = ensure_device(device);
    if (rc)
        return rc;
    SDSP_CUDA(cudaFree(nullptr)); // force context creation
    return SDSP_B200_OK;
}

int sdsp_b200_shutdown(void)
{
    // plans and banks own their allocations and die with their handles; the one library-owned cache is the set of
    // one-channel banks behind sdsp_b200_iir_process_once
    iir_release_process_once_cache();
    fft_release_l2_persist(); // the persisting-L2 carve-out of the large-frame FFT kernels (device-wide state)
    return SDSP_B200_OK;
}

int sdsp_b200_shard_range(size_t total, int rank, int world, size_t *first, size_t *count)
{
    if (!first || !count || world < 1 || rank < 0 || rank >= world)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "shard_range: rank %d of %d", rank, world);
    const size_t base = total / (size_t)world, extra = total % (size_t)world, r = (size_t)rank;
    *first = r * base + (r < extra ? r : extra);
    *count = base + (r < extra ? 1 : 0);
    return SDSP_B200_OK;
}

int sdsp_b200_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "host_alloc: null out pointer");
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess)
        return cuda_fail((int)e, "cudaHostAlloc", __FILE__, __LINE__);
    return SDSP_B200_OK;
}

int sdsp_b200_host_free(void *ptr)
{
    if (ptr)
        SDSP_CUDA(cudaFreeHost(ptr));
    return SDSP_B200_OK;
}

int sdsp_b200_device_alloc(void **ptr, size_t bytes, int device)
{
    if (!ptr)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "device_alloc: null out pointer");
    *ptr = nullptr;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    cudaError_t e = cudaMalloc(ptr, bytes ? bytes : 1);
    if (e != cudaSuccess)
        return cuda_fail((int)e, "cudaMalloc", __FILE__, __LINE__);
    return SDSP_B200_OK;
}

int sdsp_b200_device_free(void *ptr, int device)
{
    if (!ptr)
        return SDSP_B200_OK;
    int rc = ensure_device(device);
    if (rc)
        return rc;
    SDSP_CUDA(cudaFree(ptr));
    return SDSP_B200_OK;
}

int sdsp_b200_memcpy(void *dst, const void *src, size_t bytes, int device)
{
    int rc = ensure_device(device);
    if (rc)
        return rc;
    SDSP_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return SDSP_B200_OK;
}

int sdsp_b200_device_synchronize(int device)
{
    int rc = ensure_device(device);
    if (rc)
        return rc;
    SDSP_CUDA(cudaDeviceSynchronize());
    return SDSP_B200_OK;
}
}
