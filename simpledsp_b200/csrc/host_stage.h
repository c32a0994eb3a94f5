// host_stage.h -- what a plan / bank keeps for calls on HOST pointers (internal).
//
// The reference's objects work on caller-owned host buffers in place (fft.h:290-291, casc_2o_iir.h:71); a host-pointer
// call of the C ABI therefore stages the data through device memory and is synchronous.  Two regimes:
//   small calls (the drop-in headers' one frame / one filter object, BASELINE config 1): the time is launch and copy
//     latency, so the data bounces through a pinned buffer owned by the handle and everything is queued asynchronously
//     on one persistent stream, with a single synchronisation at the end;
//   large calls: slabs (FFT: frames; IIR: stretches of time across all channels) alternate between two persistent
//     streams, so the H2D copy of one slab overlaps the kernel and the D2H copy of the one before; the KERNELS are
//     chained by events (slab k+1's kernel waits for slab k's), because they share the handle's scratch memory,
//     counters (FFT, n >= 32768) or filter history (IIR).
#pragma once
#include <cstddef>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime.h>

#include "common.h"

namespace sdsp_b200
{
#ifndef SDSP_HOST_SLAB_MB_DEFAULT
#define SDSP_HOST_SLAB_MB_DEFAULT 64
#endif
struct HostStage {
    static constexpr size_t BOUNCE_BYTES = 1u << 20; // calls up to this size take the pinned bounce buffer
    cudaStream_t stream[2] = { nullptr, nullptr };
    cudaEvent_t kernel_done[2] = { nullptr, nullptr };
    void *bounce = nullptr; // pinned host memory, BOUNCE_BYTES
    bool ready = false;

    int ensure()
    {
        if (ready)
            return SDSP_B200_OK;
        for (int i = 0; i < 2; i++) {
            SDSP_CUDA(cudaStreamCreateWithFlags(&stream[i], cudaStreamNonBlocking));
            SDSP_CUDA(cudaEventCreateWithFlags(&kernel_done[i], cudaEventDisableTiming));
        }
        SDSP_CUDA(cudaHostAlloc(&bounce, BOUNCE_BYTES, cudaHostAllocDefault));
        ready = true;
        return SDSP_B200_OK;
    }
    void release()
    {
        for (int i = 0; i < 2; i++) {
            if (kernel_done[i])
                cudaEventDestroy(kernel_done[i]);
            if (stream[i])
                cudaStreamDestroy(stream[i]);
            kernel_done[i] = nullptr;
            stream[i] = nullptr;
        }
        if (bounce)
            cudaFreeHost(bounce);
        bounce = nullptr;
        ready = false;
    }
    // both streams idle; returns the first error seen (and clears it)
    int drain(int rc)
    {
        for (int i = 0; i < 2; i++) {
            cudaError_t e = cudaStreamSynchronize(stream[i]);
            if (e != cudaSuccess && rc == SDSP_B200_OK)
                rc = cuda_fail((int)e, "cudaStreamSynchronize (host staging)", __FILE__, __LINE__);
        }
        return rc;
    }
};

// grow-only device staging buffer of a handle
// bytes of one staging slab of a host-pointer call (two are in flight).  A call's first H2D and last D2H are not overlapped with
// anything, so a slab costs its own transfer time once per call: 64 MB = 2.8 ms of a 45 ms call on 2 GiB (6 %), 16 MB = 0.7 ms.
// SDSP_B200_HOST_SLAB_MB overrides (tuning aid; profiles/r02_host_slab_sweep.txt).
inline size_t host_slab_bytes()
{
    static size_t v = 0;
    if (!v) {
        const char *e = getenv("SDSP_B200_HOST_SLAB_MB");
        const int mb = e ? atoi(e) : 0;
        v = (size_t)(mb >= 1 && mb <= 1024 ? mb : SDSP_HOST_SLAB_MB_DEFAULT) << 20;
    }
    return v;
}

inline int ensure_device_stage(void *&d_stage, size_t &stage_bytes, size_t need, const char *who)
{
    if (stage_bytes >= need)
        return SDSP_B200_OK;
    if (d_stage)
        cudaFree(d_stage);
    d_stage = nullptr;
    stage_bytes = 0;
    if (cudaMalloc(&d_stage, need) != cudaSuccess) {
        cudaGetLastError();
        return set_error(SDSP_B200_ERR_OOM, "%s: cannot allocate %zu bytes of staging memory", who, need);
    }
    stage_bytes = need;
    return SDSP_B200_OK;
}
} // namespace sdsp_b200
