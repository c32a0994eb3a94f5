// sdsp/fft.h -- drop-in replacement for the reference's header of the same name, backed by
// libsdsp_b200.so (CUDA, sm_100a).  Host-only C++17; include it exactly as the reference is included:
//
//     #include "sdsp/fft.h"
//     sdsp::complex_array<1024> frame{ ... };
//     sdsp::fft_radix4(frame);                       // forward
//     sdsp::fft_radix2<sdsp::reverse_fft>(frame);    // reverse, scaled by 1/N
//
// Names, template parameters and static_asserts follow reference include/sdsp/fft.h (cited per item).
// Nothing is computed on the CPU: each call hands the frame to sdsp_b200_fft_exec(); a failure (no
// device, library missing at link time) surfaces as std::runtime_error.  The batched overloads at the
// bottom are additions -- the reference transforms one frame per call.
#pragma once
#include <array>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../sdsp_b200.h"

namespace sdsp
{
using uint = unsigned int; // the reference relies on glibc's ::uint

// ---- integer helpers: reference fft.h:12-43 --------------------------------------------------
constexpr uint log2(uint num)
{
    uint r{ 0 };
    for (num >>= 1; num > 0u; num >>= 1)
        r++;
    return r;
}
constexpr uint log4(uint num)
{
    uint r{ 0 };
    for (num >>= 2; num > 0u; num >>= 2)
        r++;
    return r;
}
constexpr bool isPowerOf2(uint num)
{
    return num != 0 && (num & (num - 1)) == 0;
}
constexpr bool isPowerOf4(uint num)
{
    return isPowerOf2(num) && (log2(num) % 2 == 0);
}

// ---- types: reference fft.h:45-52 ------------------------------------------------------------
template <size_t N>
using trig_array = std::array<std::array<double, N>, log2(N)>;
template <size_t N>
using coeff_array = std::array<std::array<std::complex<double>, N>, log2(N)>;
template <size_t N>
using complex_array = std::array<std::complex<double>, N>;
// addition: single-precision frames
template <size_t N>
using complex_array_f = std::array<std::complex<float>, N>;

namespace detail
{
    inline void check(int status, const char *what)
    {
        if (status != SDSP_B200_OK)
            throw std::runtime_error(std::string(what) + ": " + sdsp_b200_last_error());
    }

    // one plan per (N, radix, precision, direction), created on first use: the counterpart of the
    // reference's function-local constexpr tables (fft.h:264-265, 307-309)
    struct plan_holder {
        sdsp_b200_fft_plan plan{ nullptr };
        plan_holder(uint32_t n, int radix, int precision, int direction)
        {
            check(sdsp_b200_fft_plan_create(&plan, n, radix, precision, direction, 0), "sdsp_b200_fft_plan_create");
        }
        ~plan_holder()
        {
            sdsp_b200_fft_plan_destroy(plan);
        }
        plan_holder(const plan_holder &) = delete;
        plan_holder &operator=(const plan_holder &) = delete;
    };

    template <typename S>
    constexpr int precision_of()
    {
        static_assert(std::is_same_v<S, float> || std::is_same_v<S, double>, "float or double");
        return std::is_same_v<S, float> ? SDSP_B200_F32 : SDSP_B200_F64;
    }

    template <class T, int RADIX, typename S>
    void run(std::complex<S> *frames, uint32_t n, size_t n_frames, int ptr_kind = SDSP_B200_PTR_HOST, void *stream = nullptr)
    {
        // plans are keyed by size at run time for the pointer overloads
        thread_local std::vector<std::pair<uint32_t, plan_holder *>> cache;
        plan_holder *h = nullptr;
        for (auto &e : cache)
            if (e.first == n)
                h = e.second;
        if (!h) {
            h = new plan_holder(n, RADIX, precision_of<S>(), T::Direction());
            cache.emplace_back(n, h);
        }
        check(sdsp_b200_fft_exec(h->plan, frames, n_frames, ptr_kind, stream), "sdsp_b200_fft_exec");
    }

    template <class T, int RADIX, typename S>
    void run_real(const S *real_frames, std::complex<S> *spectra, uint32_t n, size_t n_frames, int ptr_kind, void *stream)
    {
        thread_local std::vector<std::pair<uint32_t, plan_holder *>> cache;
        plan_holder *h = nullptr;
        for (auto &e : cache)
            if (e.first == n)
                h = e.second;
        if (!h) {
            h = new plan_holder(n, RADIX, precision_of<S>(), T::Direction());
            cache.emplace_back(n, h);
        }
        check(sdsp_b200_fft_exec_real(h->plan, real_frames, spectra, n_frames, ptr_kind, stream), "sdsp_b200_fft_exec_real");
    }
} // namespace detail

// ---- direction policies: reference fft.h:121-146 ---------------------------------------------
// Sign() keeps the reference's meaning (+1 forward: e^{-i theta}; -1 reverse).  ScaleValues is kept
// for source compatibility; on the device the 1/N factor is fused into the last pass.
class reverse_fft {
public:
    constexpr static double Sign()
    {
        return -1.0;
    }
    constexpr static int Direction()
    {
        return SDSP_B200_REVERSE;
    }
    template <size_t N>
    constexpr static void ScaleValues(complex_array<N> &data)
    {
        for (auto &v : data)
            v *= (1.0 / N);
    }
};

class forward_fft {
public:
    constexpr static double Sign()
    {
        return 1.0;
    }
    constexpr static int Direction()
    {
        return SDSP_B200_FORWARD;
    }
    template <size_t N>
    constexpr static void ScaleValues(complex_array<N> &)
    {
    }
};

// ---- tables: reference fft.h:148-256 ---------------------------------------------------------
// digit_reverse<N, base>: reverse the base-`base` digits of an index of log2(N) bits (fft.h:217-236)
template <size_t N, uint base>
constexpr uint digit_reverse(uint n)
{
    static_assert(base == 2 || base == 4, "base 2 or 4");
    constexpr uint bits{ log2(base) };
    constexpr uint digits{ log2(static_cast<uint>(N)) / bits };
    uint r{ 0 };
    for (uint d{ 0 }; d < digits; d++) {
        r = (r << bits) | (n & (base - 1));
        n >>= bits;
    }
    return r;
}

// calc_swap_lookup<N, base>: the table a linear sweep uses to swap every pair once (fft.h:238-256)
template <size_t N, uint base>
constexpr std::array<uint, N> calc_swap_lookup()
{
    std::array<uint, N> t{};
    for (size_t i{ 0 }; i < N; i++) {
        const uint r{ digit_reverse<N, base>(static_cast<uint>(i)) };
        t[i] = r < i ? static_cast<uint>(i) : r;
    }
    return t;
}

// calc_wCoeffs<N, T>: W[i][j] = exp(-i * Sign * 2 pi j / 2^(i+1)) (fft.h:197-214).  Produced by the
// library's table generator (the one that fills the device tables); not constexpr.
template <size_t N, class T>
coeff_array<N> calc_wCoeffs()
{
    coeff_array<N> w{};
    detail::check(sdsp_b200_twiddle_table(static_cast<uint32_t>(N), T::Direction(), reinterpret_cast<double *>(w.data())),
                  "sdsp_b200_twiddle_table");
    return w;
}

// ---- transforms: reference fft.h:258-299 and 301-360 -----------------------------------------
template <class T = forward_fft, size_t N>
void fft_radix2(complex_array<N> &data)
{
    static_assert(isPowerOf2(N), "FFT size must be a power of 2!");
    static detail::plan_holder holder(static_cast<uint32_t>(N), 2, SDSP_B200_F64, T::Direction());
    detail::check(sdsp_b200_fft_exec(holder.plan, data.data(), 1, SDSP_B200_PTR_HOST, nullptr), "sdsp_b200_fft_exec");
}

template <class T = forward_fft, size_t N>
void fft_radix4(complex_array<N> &data)
{
    static_assert(isPowerOf4(N), "FFT radix 4 size must be a power of 4!");
    static detail::plan_holder holder(static_cast<uint32_t>(N), 4, SDSP_B200_F64, T::Direction());
    detail::check(sdsp_b200_fft_exec(holder.plan, data.data(), 1, SDSP_B200_PTR_HOST, nullptr), "sdsp_b200_fft_exec");
}

// ---- additions: single precision and batches -------------------------------------------------
template <class T = forward_fft, size_t N>
void fft_radix2(complex_array_f<N> &data)
{
    static_assert(isPowerOf2(N), "FFT size must be a power of 2!");
    detail::run<T, 2, float>(data.data(), static_cast<uint32_t>(N), 1);
}
template <class T = forward_fft, size_t N>
void fft_radix4(complex_array_f<N> &data)
{
    static_assert(isPowerOf4(N), "FFT radix 4 size must be a power of 4!");
    detail::run<T, 4, float>(data.data(), static_cast<uint32_t>(N), 1);
}
// n_frames contiguous frames of n points (host memory), transformed in place
template <class T = forward_fft, typename S>
void fft_radix2(std::complex<S> *frames, size_t n, size_t n_frames)
{
    detail::run<T, 2, S>(frames, static_cast<uint32_t>(n), n_frames);
}
template <class T = forward_fft, typename S>
void fft_radix4(std::complex<S> *frames, size_t n, size_t n_frames)
{
    detail::run<T, 4, S>(frames, static_cast<uint32_t>(n), n_frames);
}
// the same on device-resident frames, asynchronously on `stream` (a cudaStream_t)
template <class T = forward_fft, typename S>
void fft_radix4_device(std::complex<S> *frames, size_t n, size_t n_frames, void *stream = nullptr)
{
    detail::run<T, 4, S>(frames, static_cast<uint32_t>(n), n_frames, SDSP_B200_PTR_DEVICE, stream);
}
template <class T = forward_fft, typename S>
void fft_radix2_device(std::complex<S> *frames, size_t n, size_t n_frames, void *stream = nullptr)
{
    detail::run<T, 2, S>(frames, static_cast<uint32_t>(n), n_frames, SDSP_B200_PTR_DEVICE, stream);
}
// real frames in, spectra out (out of place): what the reference's callers do by hand -- real part = signal, imaginary
// part = 0 (test/testFFT.cpp:24, :86) -- folded into the transform's first load.  Device-resident buffers.
template <class T = forward_fft, typename S>
void fft_radix4_real_device(const S *real_frames, std::complex<S> *spectra, size_t n, size_t n_frames, void *stream = nullptr)
{
    detail::run_real<T, 4, S>(real_frames, spectra, static_cast<uint32_t>(n), n_frames, SDSP_B200_PTR_DEVICE, stream);
}
template <class T = forward_fft, typename S>
void fft_radix2_real_device(const S *real_frames, std::complex<S> *spectra, size_t n, size_t n_frames, void *stream = nullptr)
{
    detail::run_real<T, 2, S>(real_frames, spectra, static_cast<uint32_t>(n), n_frames, SDSP_B200_PTR_DEVICE, stream);
}
// the same on host buffers (staged through the device)
template <class T = forward_fft, typename S>
void fft_radix4_real(const S *real_frames, std::complex<S> *spectra, size_t n, size_t n_frames)
{
    detail::run_real<T, 4, S>(real_frames, spectra, static_cast<uint32_t>(n), n_frames, SDSP_B200_PTR_HOST, nullptr);
}
} // namespace sdsp
