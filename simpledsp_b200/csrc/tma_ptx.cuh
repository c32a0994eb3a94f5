// tma_ptx.cuh -- inline-PTX wrappers for mbarrier / TMA (cp.async.bulk.tensor) and the tensor-map encoder
// lookup, shared by the TMA-fed kernels (internal).
#pragma once
#include <cstdint>

#include <cuda.h>
#include <cuda_runtime.h>

namespace sdsp_b200
{
namespace
{
    // ---- PTX wrappers ---------------------------------------------------------------------------
    __device__ __forceinline__ uint32_t smem_u32(const void *p)
    {
        return (uint32_t)__cvta_generic_to_shared(p);
    }
    __device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    }
    __device__ __forceinline__ void fence_mbar_init()
    {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __device__ __forceinline__ void fence_proxy_async()
    {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) // one try_wait: may suspend up to the hardware time limit
    {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        return ok != 0;
    }
    __device__ __forceinline__ void mbar_expect_tx_only(uint64_t *bar, uint32_t bytes) // raises the byte count, no arrival
    {
        asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void mbar_arrive(uint64_t *bar) // release at CTA scope: orders this thread's earlier writes
    {
        asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    __device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
    {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "WAIT_LOOP:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra WAIT_DONE;\n\t"
            "bra WAIT_LOOP;\n\t"
            "WAIT_DONE:\n\t"
            "}" ::"r"(smem_u32(bar)),
            "r"(parity)
            : "memory");
    }
    __device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) // non-blocking
    {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        return ok != 0;
    }
    __device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
    {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_u32(smem_dst)),
                     "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                     : "memory");
    }
    __device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int x, int y, const void *smem_src)
    {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y),
                     "r"(smem_u32(smem_src))
                     : "memory");
    }
    __device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar)
    {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                         smem_u32(smem_dst)),
                     "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                     : "memory");
    }
    __device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int x, int y, int z, const void *smem_src)
    {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(x), "r"(y), "r"(z),
                     "r"(smem_u32(smem_src))
                     : "memory");
    }
    __device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, int x, int y, int z, int w, uint64_t *bar)
    {
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                         smem_u32(smem_dst)),
                     "l"(map), "r"(x), "r"(y), "r"(z), "r"(w), "r"(smem_u32(bar))
                     : "memory");
    }
    __device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, int x, int y, int z, int w, const void *smem_src)
    {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(map), "r"(x), "r"(y), "r"(z),
                     "r"(w), "r"(smem_u32(smem_src))
                     : "memory");
    }
    __device__ __forceinline__ void tma_commit()
    {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    template <int N>
    __device__ __forceinline__ void tma_wait_read()
    {
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
    }
    __device__ __forceinline__ void tma_wait_all()
    {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    template <typename T>
    struct Vec16; // the 16-byte vector a lane moves per LDS/STS
    template <>
    struct Vec16<float> {
        using type = float4;
        static constexpr int N = 4;
    };
    template <>
    struct Vec16<double> {
        using type = double2;
        static constexpr int N = 2;
    };
    __device__ __forceinline__ float vget(const float4 &v, int i)
    {
        return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
    }
    __device__ __forceinline__ double vget(const double2 &v, int i)
    {
        return i == 0 ? v.x : v.y;
    }
    __device__ __forceinline__ void vset(float4 &v, int i, float x)
    {
        if (i == 0)
            v.x = x;
        else if (i == 1)
            v.y = x;
        else if (i == 2)
            v.z = x;
        else
            v.w = x;
    }
    __device__ __forceinline__ void vset(double2 &v, int i, double x)
    {
        if (i == 0)
            v.x = x;
        else
            v.y = x;
    }
} // namespace


typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline encode_tiled_fn get_encode_fn()
{
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

} // namespace sdsp_b200
