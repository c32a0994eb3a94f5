// additions_test.cpp -- exercises the ADDITIVE part of the drop-in headers (batches, fp32, real-input frames, device
// pointers staged by the library, sdsp::iir_bank and its time-parallel path) against the
// reference-shaped single-object API of the same headers, which the reference's own tests pin.  Built by
// `make -C oracle additions` against include/sdsp/*.h + libsdsp_b200.so; run by tests/test_gpu_reference_tests.py.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sdsp/casc_2o_iir.h"
#include "sdsp/fft.h"

static int failures = 0;
#define CHECK(cond, ...)                                 \
    do {                                                 \
        if (!(cond)) {                                   \
            failures++;                                  \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);                    \
            std::printf("\n");                           \
        }                                                \
    } while (0)

static double rel_l2(const std::complex<double> *a, const std::complex<double> *b, size_t n)
{
    double num = 0, den = 0;
    for (size_t i = 0; i < n; i++) {
        num += std::norm(a[i] - b[i]);
        den += std::norm(b[i]);
    }
    return std::sqrt(num / den);
}

int main()
{
    std::srand(7);
    auto rnd = [] { return (double)std::rand() / RAND_MAX - 0.5; };

    // ---- batched fp64 frames against the one-frame template call (reference signature, fft.h:301-302)
    {
        constexpr size_t N = 1024, frames = 9;
        std::vector<std::complex<double>> batch(N * frames), ref(N * frames);
        for (auto &v : batch)
            v = { rnd(), rnd() };
        ref = batch;
        for (size_t f = 0; f < frames; f++) {
            sdsp::complex_array<N> one;
            for (size_t i = 0; i < N; i++)
                one[i] = ref[f * N + i];
            sdsp::fft_radix4(one);
            for (size_t i = 0; i < N; i++)
                ref[f * N + i] = one[i];
        }
        sdsp::fft_radix4(batch.data(), N, frames);
        CHECK(rel_l2(batch.data(), ref.data(), N * frames) == 0.0, "batched fp64 transform differs from frame-by-frame calls");
        sdsp::fft_radix4<sdsp::reverse_fft>(batch.data(), N, frames); // back to the signal ...
        sdsp::fft_radix4(batch.data(), N, frames);                    // ... and forward again: the spectra once more
        CHECK(rel_l2(batch.data(), ref.data(), N * frames) < 1e-13, "forward(reverse(X)) != X in fp64");
    }
    // ---- fp32 batch and real-input frames
    {
        constexpr size_t N = 4096, frames = 5;
        std::vector<float> real(N * frames);
        for (auto &v : real)
            v = (float)rnd();
        std::vector<std::complex<float>> zc(N * frames), spec(N * frames);
        for (size_t i = 0; i < N * frames; i++)
            zc[i] = { real[i], 0.f };
        sdsp::fft_radix4(zc.data(), N, frames);
        sdsp::fft_radix4_real(real.data(), spec.data(), N, frames);
        bool same = true;
        for (size_t i = 0; i < N * frames; i++)
            same = same && zc[i] == spec[i];
        CHECK(same, "real-input transform differs from the complex entry point fed (x, 0)");
        // against fp64 on the same data
        std::vector<std::complex<double>> zd(N * frames), sd(N * frames);
        for (size_t i = 0; i < N * frames; i++) {
            zd[i] = { real[i], 0.0 };
            sd[i] = spec[i];
        }
        sdsp::fft_radix4(zd.data(), N, frames);
        CHECK(rel_l2(sd.data(), zd.data(), N * frames) < 1e-5, "fp32 real-input transform off by more than 1e-5 rel-L2");
        // half spectra: bins 0 .. N/2, frames N/2 + 1 bins apart, against the fp64 spectra of the same frames
        std::vector<std::complex<float>> half((N / 2 + 1) * frames);
        sdsp::fft_half_spectrum(real.data(), half.data(), N, frames);
        std::vector<std::complex<double>> hd((N / 2 + 1) * frames), hr((N / 2 + 1) * frames);
        for (size_t f = 0; f < frames; f++)
            for (size_t k = 0; k <= N / 2; k++) {
                hd[f * (N / 2 + 1) + k] = half[f * (N / 2 + 1) + k];
                hr[f * (N / 2 + 1) + k] = zd[f * N + k];
            }
        CHECK(rel_l2(hd.data(), hr.data(), hd.size()) < 1e-5, "fp32 half spectrum off by more than 1e-5 rel-L2");
        std::vector<double> reald(real.begin(), real.end());
        std::vector<std::complex<double>> halfd((N / 2 + 1) * frames);
        sdsp::fft_half_spectrum(reald.data(), halfd.data(), N, frames);
        CHECK(rel_l2(halfd.data(), hr.data(), hr.size()) < 1e-12, "fp64 half spectrum off by more than 1e-12 rel-L2");
        // and back: the real frames again, 1/N included
        std::vector<double> backd(N * frames);
        sdsp::fft_real_from_half_spectrum(halfd.data(), backd.data(), N, frames);
        double num = 0, den = 0;
        for (size_t i = 0; i < N * frames; i++) {
            num += (backd[i] - reald[i]) * (backd[i] - reald[i]);
            den += reald[i] * reald[i];
        }
        CHECK(std::sqrt(num / den) < 1e-12, "fp64 real_from_half_spectrum(half_spectrum(x)) != x");
        // one frame in std::array containers, the reference's style
        std::array<double, 1024> one{};
        for (size_t i = 0; i < one.size(); i++)
            one[i] = reald[i];
        std::array<std::complex<double>, 513> one_half{};
        sdsp::fft_half_spectrum<1024>(one, one_half);
        sdsp::complex_array<1024> full{};
        for (size_t i = 0; i < one.size(); i++)
            full[i] = { one[i], 0.0 };
        sdsp::fft_radix2(full);
        CHECK(rel_l2(one_half.data(), full.data(), 513) < 1e-12, "std::array half spectrum differs from fft_radix2 on (x, 0)");
        std::array<double, 1024> one_back{};
        sdsp::fft_real_from_half_spectrum<1024>(one_half, one_back);
        double e2 = 0, n2 = 0;
        for (size_t i = 0; i < one.size(); i++) {
            e2 += (one_back[i] - one[i]) * (one_back[i] - one[i]);
            n2 += one[i] * one[i];
        }
        CHECK(std::sqrt(e2 / n2) < 1e-12, "std::array round trip through the half spectrum");
    }
    // ---- a bank of channels against one filter object per channel (reference signature, casc_2o_iir.h:36-80)
    {
        constexpr size_t C = 37, n = 3000;
        sdsp::iir_bank<4, double> bank(C);
        std::vector<double> x(C * n), ref;
        for (auto &v : x)
            v = rnd();
        ref = x;
        for (size_t c = 0; c < C; c++) {
            sdsp::casc_2o_iir<4> f;
            if (c % 2)
                f.set_hp_coeff(1000.0 + 400.0 * c, 100e3);
            else
                f.set_lp_coeff(1000.0 + 400.0 * c, 100e3);
            bank.copy_coeff_from(c, f);
            f.process(ref.begin() + c * n, ref.begin() + (c + 1) * n);
        }
        bank.process(x.data(), n, n);
        bool same = true;
        for (size_t i = 0; i < C * n; i++)
            same = same && x[i] == ref[i];
        CHECK(same, "iir_bank (bit-exact streaming path) differs from one casc_2o_iir object per channel");
    }
    // ---- one long channel through the time-parallel path against the sequential object
    {
        constexpr size_t n = 400000;
        sdsp::casc_2o_iir<4> f;
        f.set_lp_coeff(10e3, 100e3);
        sdsp::iir_bank<4, double> bank(1);
        bank.copy_coeff_from(0, f);
        std::vector<double> x(n), ref;
        for (auto &v : x)
            v = rnd();
        ref = x;
        f.process(ref.begin(), ref.end());
        bank.process(x.data(), n, n, SDSP_B200_IIR_SCAN);
        double err = 0, peak = 0;
        for (size_t i = 0; i < n; i++) {
            err = std::fmax(err, std::fabs(x[i] - ref[i]));
            peak = std::fmax(peak, std::fabs(ref[i]));
        }
        CHECK(err / peak < 1e-10, "time-parallel path off by %.3e of peak", err / peak);
    }
    std::printf("%s (%d failures)\n", failures ? "FAILED" : "all additions ok", failures);
    return failures ? 1 : 0;
}
