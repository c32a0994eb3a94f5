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
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <memory>
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
        // plans are keyed by size at run time for the pointer overloads; the cache owns them (released at thread exit)
        thread_local std::vector<std::pair<uint32_t, std::unique_ptr<plan_holder>>> cache;
        plan_holder *h = nullptr;
        for (auto &e : cache)
            if (e.first == n)
                h = e.second.get();
        if (!h) {
            cache.emplace_back(n, std::make_unique<plan_holder>(n, RADIX, precision_of<S>(), T::Direction()));
            h = cache.back().second.get();
        }
        check(sdsp_b200_fft_exec(h->plan, frames, n_frames, ptr_kind, stream), "sdsp_b200_fft_exec");
    }

    template <typename S>
    void run_c2r(const std::complex<S> *half_spectra, S *real_frames, uint32_t n, size_t n_frames, int ptr_kind, void *stream)
    {
        thread_local std::vector<std::pair<uint32_t, std::unique_ptr<plan_holder>>> cache;
        plan_holder *h = nullptr;
        for (auto &e : cache)
            if (e.first == n)
                h = e.second.get();
        if (!h) {
            cache.emplace_back(n, std::make_unique<plan_holder>(n, 2, precision_of<S>(), SDSP_B200_REVERSE));
            h = cache.back().second.get();
        }
        check(sdsp_b200_fft_exec_c2r(h->plan, half_spectra, real_frames, n_frames, ptr_kind, stream), "sdsp_b200_fft_exec_c2r");
    }

    template <class T, int RADIX, typename S, bool HALF = false>
    void run_real(const S *real_frames, std::complex<S> *spectra, uint32_t n, size_t n_frames, int ptr_kind, void *stream)
    {
        thread_local std::vector<std::pair<uint32_t, std::unique_ptr<plan_holder>>> cache;
        plan_holder *h = nullptr;
        for (auto &e : cache)
            if (e.first == n)
                h = e.second.get();
        if (!h) {
            cache.emplace_back(n, std::make_unique<plan_holder>(n, RADIX, precision_of<S>(), T::Direction()));
            h = cache.back().second.get();
        }
        if (HALF)
            check(sdsp_b200_fft_exec_r2c(h->plan, real_frames, spectra, n_frames, ptr_kind, stream), "sdsp_b200_fft_exec_r2c");
        else
            check(sdsp_b200_fft_exec_real(h->plan, real_frames, spectra, n_frames, ptr_kind, stream), "sdsp_b200_fft_exec_real");
    }
} // namespace detail

// ---- trigonometric tables: reference fft.h:54-119, 148-214 --------------------------------------
// The policy classes keep the reference's member names: the value at 0 and 90 degrees, the function itself, and the
// sign a value picks up when the table walk is reflected at 0 / at 90 degrees.
class sine_calculator {
public:
    constexpr static double Value0() { return 0.0; }
    constexpr static double Value90() { return 1.0; }
    constexpr static double Value(double rad) { return std::sin(rad); }
    constexpr static double Sym0() { return -1.0; }
    constexpr static double Sym90() { return 1.0; }
};
class cosine_calculator {
public:
    constexpr static double Value0() { return 1.0; }
    constexpr static double Value90() { return 0.0; }
    constexpr static double Value(double rad) { return std::cos(rad); }
    constexpr static double Sym0() { return 1.0; }
    constexpr static double Sym90() { return -1.0; }
};

// row i, column j: T(2 pi j / 2^(i+1)) straight from libm (fft.h:54-65; the reference keeps it as the unused alternative)
template <size_t N, class T>
constexpr trig_array<N> calc_trigs_naive()
{
    trig_array<N> t{};
    for (size_t i{ 0 }; i < t.size(); i++)
        for (size_t j{ 0 }; j < N; j++)
            t[i][j] = T::Value(2 * M_PI * static_cast<double>(j) / static_cast<double>(size_t{ 2 } << i));
    return t;
}

// The same table with libm evaluated on the first quarter wave only (fft.h:148-194): 0 and 90 degrees are exact, every
// other entry is a first-quadrant value times the product of the reflection signs collected on the way there -- so the
// table has the exact symmetries the butterflies rely on.  Written here in closed form per index (quadrant = j / quarter
// wave) instead of the reference's running walk; an entry that falls ON a reflection point carries the sign of the
// quadrant the walk arrives from, which reproduces the reference's table down to the sign of its zeros.
template <size_t N, class T>
constexpr trig_array<N> calc_trigs()
{
    trig_array<N> t{};
    for (size_t i{ 0 }; i < t.size(); i++) {
        const size_t period{ size_t{ 2 } << i }, quarter{ period / 4 };
        for (size_t j{ 0 }; j < N; j++) {
            if (i == 0) { // period 2: +v, -v, +v, ... (each entry the previous one negated)
                t[i][j] = (j == 0) ? T::Value0() : t[i][j - 1] * -1.0;
                continue;
            }
            // i == 1 has an empty first quadrant (quarter == 1): only the exact points exist
            const size_t quad{ j / quarter }, r{ j % quarter };
            const size_t from{ (r == 0 && quad > 0) ? quad - 1 : quad }; // reflection points belong to the quadrant before
            const size_t fold{ (quad % 2 == 0) ? r : quarter - r };       // index on the first quarter wave, 0 .. quarter
            const double v{ fold == 0 ? T::Value0() :
                            fold == quarter ? T::Value90() :
                                              T::Value(2 * M_PI * static_cast<double>(fold) / static_cast<double>(period)) };
            double sign{ 1.0 };
            for (size_t q{ 0 }; q < from % 4; q++) // quadrant signs: 1, Sym90, Sym90 Sym0, Sym90 Sym0 Sym90
                sign *= (q % 2 == 0) ? T::Sym90() : T::Sym0();
            t[i][j] = v * sign;
        }
    }
    return t;
}

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

// calc_wCoeffs<N, T>: W[i][j] = cos - i Sign sin of 2 pi j / 2^(i+1) (fft.h:197-214), from the two tables above -- constexpr
// like the reference's, for callers that build their own tables from it.  (The transforms below do not read it: the
// device tables come from the library's generator, whose host-side copy is calc_wCoeffs_library.)
template <size_t N, class T>
constexpr coeff_array<N> calc_wCoeffs()
{
    const trig_array<N> c{ calc_trigs<N, cosine_calculator>() };
    const trig_array<N> s{ calc_trigs<N, sine_calculator>() };
    coeff_array<N> w{};
    for (size_t i{ 0 }; i < w.size(); i++)
        for (size_t j{ 0 }; j < N; j++)
            w[i][j] = std::complex<double>(c[i][j], T::Sign() * -1.0 * s[i][j]);
    return w;
}
// addition: the same table as the library's own generator produces it (octant-symmetric, long double)
template <size_t N, class T>
coeff_array<N> calc_wCoeffs_library()
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
    static_assert(N <= SDSP_B200_FFT_MAX_N_F64, "libsdsp_b200 transforms double frames of up to 2^17 points");
    static detail::plan_holder holder(static_cast<uint32_t>(N), 2, SDSP_B200_F64, T::Direction());
    detail::check(sdsp_b200_fft_exec(holder.plan, data.data(), 1, SDSP_B200_PTR_HOST, nullptr), "sdsp_b200_fft_exec");
}

template <class T = forward_fft, size_t N>
void fft_radix4(complex_array<N> &data)
{
    static_assert(isPowerOf4(N), "FFT radix 4 size must be a power of 4!");
    static_assert(N <= SDSP_B200_FFT_MAX_N_F64, "libsdsp_b200 transforms double frames of up to 2^17 points");
    static detail::plan_holder holder(static_cast<uint32_t>(N), 4, SDSP_B200_F64, T::Direction());
    detail::check(sdsp_b200_fft_exec(holder.plan, data.data(), 1, SDSP_B200_PTR_HOST, nullptr), "sdsp_b200_fft_exec");
}

// ---- additions: single precision and batches -------------------------------------------------
template <class T = forward_fft, size_t N>
void fft_radix2(complex_array_f<N> &data)
{
    static_assert(isPowerOf2(N), "FFT size must be a power of 2!");
    static_assert(N <= SDSP_B200_FFT_MAX_N_F32, "libsdsp_b200 transforms float frames of up to 2^18 points");
    detail::run<T, 2, float>(data.data(), static_cast<uint32_t>(N), 1);
}
template <class T = forward_fft, size_t N>
void fft_radix4(complex_array_f<N> &data)
{
    static_assert(isPowerOf4(N), "FFT radix 4 size must be a power of 4!");
    static_assert(N <= SDSP_B200_FFT_MAX_N_F32, "libsdsp_b200 transforms float frames of up to 2^18 points");
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
// real frames in, HALF spectra out: bins 0 .. n/2 of each frame (n/2 + 1 values, frames n/2 + 1 values apart); the rest of the
// spectrum of a real signal is the conjugate mirror, X[n - k] = conj X[k].  Half the bytes of the calls above in either direction.
template <typename S>
void fft_half_spectrum(const S *real_frames, std::complex<S> *half_spectra, size_t n, size_t n_frames)
{
    detail::run_real<forward_fft, 2, S, true>(real_frames, half_spectra, static_cast<uint32_t>(n), n_frames, SDSP_B200_PTR_HOST, nullptr);
}
template <typename S>
void fft_half_spectrum_device(const S *real_frames, std::complex<S> *half_spectra, size_t n, size_t n_frames, void *stream = nullptr)
{
    detail::run_real<forward_fft, 2, S, true>(real_frames, half_spectra, static_cast<uint32_t>(n), n_frames, SDSP_B200_PTR_DEVICE, stream);
}
// one frame in the reference's container style (std::array, as complex_array<N> is)
template <size_t N, typename S>
void fft_half_spectrum(const std::array<S, N> &real_frame, std::array<std::complex<S>, N / 2 + 1> &half_spectrum)
{
    static_assert(isPowerOf2(N) && N >= 4, "FFT size must be a power of 2!");
    fft_half_spectrum(real_frame.data(), half_spectrum.data(), N, 1);
}
template <size_t N, typename S>
void fft_real_from_half_spectrum(const std::array<std::complex<S>, N / 2 + 1> &half_spectrum, std::array<S, N> &real_frame);
// and back: half spectra in, real frames out, 1/n included (reverse_fft's scaling, reference fft.h:128-132)
template <typename S>
void fft_real_from_half_spectrum(const std::complex<S> *half_spectra, S *real_frames, size_t n, size_t n_frames)
{
    detail::run_c2r<S>(half_spectra, real_frames, static_cast<uint32_t>(n), n_frames, SDSP_B200_PTR_HOST, nullptr);
}
template <typename S>
void fft_real_from_half_spectrum_device(const std::complex<S> *half_spectra, S *real_frames, size_t n, size_t n_frames, void *stream = nullptr)
{
    detail::run_c2r<S>(half_spectra, real_frames, static_cast<uint32_t>(n), n_frames, SDSP_B200_PTR_DEVICE, stream);
}
template <size_t N, typename S>
void fft_real_from_half_spectrum(const std::array<std::complex<S>, N / 2 + 1> &half_spectrum, std::array<S, N> &real_frame)
{
    static_assert(isPowerOf2(N) && N >= 4, "FFT size must be a power of 2!");
    fft_real_from_half_spectrum(half_spectrum.data(), real_frame.data(), N, 1);
}
} // namespace sdsp
