// common.h -- shared declarations of libsdsp_b200 (internal).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/sdsp_b200.h"

#if defined(__CUDACC__)
#define SDSP_HD __host__ __device__ __forceinline__
#else
#define SDSP_HD inline
#endif

namespace sdsp_b200
{
// ---- error plumbing (capi.cu) -------------------------------------------------------------
int set_error(int status, const char *fmt, ...);
int cuda_fail(int cuda_error, const char *what, const char *file, int line);
#define SDSP_CUDA(expr)                                                       \
    do {                                                                      \
        cudaError_t e__ = (expr);                                             \
        if (e__ != cudaSuccess)                                               \
            return ::sdsp_b200::cuda_fail((int)e__, #expr, __FILE__, __LINE__); \
    } while (0)

int ensure_device(int device); // context + sm_100 check, sets error
int device_sm_count(int device);

// ---- complex value type ---------------------------------------------------------------------
// Same layout as std::complex<T> / float2 / double2: (re, im) interleaved, naturally aligned so a
// whole element moves with one LDG.64 / LDG.128.
template <typename T>
struct alignas(2 * sizeof(T)) cplx {
    T x, y;
};

template <typename T>
SDSP_HD cplx<T> operator+(cplx<T> a, cplx<T> b)
{
    return { a.x + b.x, a.y + b.y };
}
template <typename T>
SDSP_HD cplx<T> operator-(cplx<T> a, cplx<T> b)
{
    return { a.x - b.x, a.y - b.y };
}

SDSP_HD float fma_t(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
SDSP_HD double fma_t(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// products and sums that must round on their own (nvcc contracts a plain a*b + c into an fma, gcc may too): the
// IIR paths promise bit-identical output however a stream is cut, which only holds if every kernel and the host
// emulation round every operation alike
SDSP_HD float mul_t(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
SDSP_HD double mul_t(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
SDSP_HD float add_t(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
SDSP_HD double add_t(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// a * w with two multiplies and two fused multiply-adds
template <typename T>
SDSP_HD cplx<T> cmul(cplx<T> a, cplx<T> w)
{
    return { fma_t(a.x, w.x, -(a.y * w.y)), fma_t(a.x, w.y, a.y * w.x) };
}

constexpr bool is_pow2(uint32_t v)
{
    return v != 0 && (v & (v - 1)) == 0;
}
constexpr int ilog2(uint32_t v)
{
    int r = 0;
    while (v >>= 1)
        r++;
    return r;
}
} // namespace sdsp_b200
