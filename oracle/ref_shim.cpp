// ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" wrapper around the UNMODIFIED simpledsp headers.  It is compiled by oracle/Makefile
// with -I$(REFERENCE)/include (default /root/reference/include), straight from where the reference
// lies; no reference source is copied into this repository.  The result, oracle/_ref/libsdsp_ref.so,
// is git-ignored but travels to the GPU box, where it serves as the "reference" CPU baseline and as
// a second parity oracle.  Only tests/, __graft_entry__.smoke() and bench.py may load it.
//
// g++ only: nvcc's front end rejects the reference's constexpr std::sin/std::cos (fft.h:77-80,104-107).
#include "sdsp/casc_2o_iir.h"
#include "sdsp/fft.h"

#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace
{
using cd = std::complex<double>;

template <size_t N>
int run_fft(double *data, int radix, int inverse)
{
    // complex_array<N> is std::array<std::complex<double>, N>: layout-compatible with N (re,im) pairs
    auto &arr = *reinterpret_cast<sdsp::complex_array<N> *>(data);
    if (radix == 2) {
        if (inverse)
            sdsp::fft_radix2<sdsp::reverse_fft>(arr);
        else
            sdsp::fft_radix2(arr);
        return 0;
    }
    if constexpr (sdsp::isPowerOf4(N)) {
        if (radix == 4) {
            if (inverse)
                sdsp::fft_radix4<sdsp::reverse_fft>(arr);
            else
                sdsp::fft_radix4(arr);
            return 0;
        }
    }
    return -1;
}

int dispatch_fft(double *data, uint32_t n, int radix, int inverse)
{
    switch (n) {
    case 4: return run_fft<4>(data, radix, inverse);
    case 8: return run_fft<8>(data, radix, inverse);
    case 16: return run_fft<16>(data, radix, inverse);
    case 32: return run_fft<32>(data, radix, inverse);
    case 64: return run_fft<64>(data, radix, inverse);
    case 128: return run_fft<128>(data, radix, inverse);
    case 256: return run_fft<256>(data, radix, inverse);
    case 512: return run_fft<512>(data, radix, inverse);
    case 1024: return run_fft<1024>(data, radix, inverse);
    case 2048: return run_fft<2048>(data, radix, inverse);
    case 4096: return run_fft<4096>(data, radix, inverse);
#ifdef SDSP_REF_BIG
    case 16384: return run_fft<16384>(data, radix, inverse);
    case 65536: return run_fft<65536>(data, radix, inverse);
#endif
    default: return -1;
    }
}

bool fft_supported(uint32_t n, int radix)
{
#ifdef SDSP_REF_BIG
    const uint32_t max_n = 65536;
#else
    const uint32_t max_n = 4096;
#endif
    if (n < 4 || n > max_n || !sdsp::isPowerOf2(n) || (n > 4096 && n != 16384 && n != 65536))
        return false;
    return radix == 2 || (radix == 4 && sdsp::isPowerOf4(n));
}

template <typename F>
void parallel_for(size_t count, int threads, F &&body)
{
    if (threads <= 1 || count <= 1) {
        for (size_t i = 0; i < count; i++)
            body(i);
        return;
    }
    std::atomic<size_t> next{ 0 };
    std::vector<std::thread> pool;
    const size_t grain = std::max<size_t>(1, count / (static_cast<size_t>(threads) * 8));
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&] {
            for (;;) {
                const size_t lo = next.fetch_add(grain);
                if (lo >= count)
                    return;
                const size_t hi = std::min(count, lo + grain);
                for (size_t i = lo; i < hi; i++)
                    body(i);
            }
        });
    }
    for (auto &th : pool)
        th.join();
}

// One filter of any class behind a uniform face.
struct filter_iface {
    virtual ~filter_iface() = default;
    virtual int set_lp(double, double, double) { return -1; }
    virtual int set_hp(double, double, double) { return -1; }
    virtual int set_bp(double, double, double, double) { return -1; }
    virtual int preload(double) { return -1; }
    virtual void process(double *, size_t) = 0;
    virtual filter_iface *clone() const = 0;
};

template <size_t M>
struct generic_filter final : filter_iface {
    sdsp::casc_2o_iir<M> f;
    int set_lp(double f0, double fs, double g) override { f.set_lp_coeff(f0, fs, g); return 0; }
    int set_hp(double f0, double fs, double g) override { f.set_hp_coeff(f0, fs, g); return 0; }
    int set_bp(double f0, double fs, double q, double g) override { f.set_bp_coeff(f0, fs, q, g); return 0; }
    int preload(double v) override { f.preload_filter(v); return 0; }
    void process(double *d, size_t n) override { f.process(d, d + n); }
    filter_iface *clone() const override { return new generic_filter<M>(*this); }
};
template <size_t M>
struct lp_filter final : filter_iface {
    sdsp::casc_2o_iir_lp<M> f;
    int set_lp(double f0, double fs, double g) override { f.set_lp_coeff(f0, fs, g); return 0; }
    void process(double *d, size_t n) override { f.process(d, d + n); }
    filter_iface *clone() const override { return new lp_filter<M>(*this); }
};
template <size_t M>
struct hp_filter final : filter_iface {
    sdsp::casc_2o_iir_hp<M> f;
    int set_hp(double f0, double fs, double g) override { f.set_hp_coeff(f0, fs, g); return 0; }
    void process(double *d, size_t n) override { f.process(d, d + n); }
    filter_iface *clone() const override { return new hp_filter<M>(*this); }
};
template <size_t M>
struct bp_filter final : filter_iface {
    sdsp::casc_2o_iir_bp<M> f;
    int set_bp(double f0, double fs, double q, double g) override { f.set_bp_coeff(f0, fs, q, g); return 0; }
    void process(double *d, size_t n) override { f.process(d, d + n); }
    filter_iface *clone() const override { return new bp_filter<M>(*this); }
};

template <size_t M>
filter_iface *make_filter(int kind)
{
    switch (kind) {
    case 0: return new generic_filter<M>();
    case 1: return new lp_filter<M>();
    case 2: return new hp_filter<M>();
    case 3: return new bp_filter<M>();
    default: return nullptr;
    }
}
} // namespace

extern "C" {

const char *sdsp_ref_describe()
{
    return "simpledsp reference headers (include/sdsp/fft.h, casc_2o_iir.h) compiled unmodified by g++"
#ifdef SDSP_REF_BIG
           " [+16384, 65536]"
#endif
        ;
}

int sdsp_ref_hardware_threads()
{
    const unsigned n = std::thread::hardware_concurrency();
    return n ? static_cast<int>(n) : 1;
}

// n interleaved (re,im) doubles, in place.  radix 2 or 4.  -1 = size/radix not instantiated.
int sdsp_ref_fft(double *data, uint32_t n, int radix, int inverse)
{
    return dispatch_fft(data, n, radix, inverse);
}

int sdsp_ref_fft_batch(double *data, uint32_t n, size_t frames, int radix, int inverse, int threads)
{
    if (!fft_supported(n, radix))
        return -1;
    std::atomic<int> rc{ 0 };
    parallel_for(frames, threads, [&](size_t f) {
        if (dispatch_fft(data + f * 2 * static_cast<size_t>(n), n, radix, inverse) != 0)
            rc = -1;
    });
    return rc;
}

// the half swap table of fft.h:238-256 for the sizes instantiated above
int sdsp_ref_swap_lookup(uint32_t n, uint32_t base, uint32_t *out)
{
#define SDSP_REF_LUT(NN)                                               \
    case NN:                                                           \
        if (base == 2) {                                               \
            auto t = sdsp::calc_swap_lookup<NN, 2>();                  \
            std::copy(t.begin(), t.end(), out);                        \
            return 0;                                                  \
        }                                                              \
        if constexpr (sdsp::isPowerOf4(NN)) {                          \
            if (base == 4) {                                           \
                auto t = sdsp::calc_swap_lookup<NN, 4>();              \
                std::copy(t.begin(), t.end(), out);                    \
                return 0;                                              \
            }                                                          \
        }                                                              \
        return -1;
    switch (n) {
        SDSP_REF_LUT(4)
        SDSP_REF_LUT(8)
        SDSP_REF_LUT(16)
        SDSP_REF_LUT(32)
        SDSP_REF_LUT(64)
        SDSP_REF_LUT(128)
        SDSP_REF_LUT(256)
        SDSP_REF_LUT(512)
        SDSP_REF_LUT(1024)
        SDSP_REF_LUT(2048)
        SDSP_REF_LUT(4096)
        SDSP_REF_LUT(16384)
        SDSP_REF_LUT(65536)
    default: return -1;
    }
#undef SDSP_REF_LUT
}

// kind: 0 casc_2o_iir, 1 casc_2o_iir_lp, 2 _hp, 3 _bp.  sections in {2,4,6,8}.
void *sdsp_ref_iir_create(int sections, int kind)
{
    switch (sections) {
    case 2: return make_filter<2>(kind);
    case 4: return make_filter<4>(kind);
    case 6: return make_filter<6>(kind);
    case 8: return make_filter<8>(kind);
    default: return nullptr;
    }
}
void sdsp_ref_iir_destroy(void *h) { delete static_cast<filter_iface *>(h); }
void *sdsp_ref_iir_clone(const void *h) { return static_cast<const filter_iface *>(h)->clone(); }
int sdsp_ref_iir_set_lp(void *h, double f0, double fs, double gain) { return static_cast<filter_iface *>(h)->set_lp(f0, fs, gain); }
int sdsp_ref_iir_set_hp(void *h, double f0, double fs, double gain) { return static_cast<filter_iface *>(h)->set_hp(f0, fs, gain); }
int sdsp_ref_iir_set_bp(void *h, double f0, double fs, double q, double gain)
{
    return static_cast<filter_iface *>(h)->set_bp(f0, fs, q, gain);
}
int sdsp_ref_iir_preload(void *h, double value) { return static_cast<filter_iface *>(h)->preload(value); }
void sdsp_ref_iir_process(void *h, double *data, size_t n) { static_cast<filter_iface *>(h)->process(data, n); }

// Channel bank: data is planar [channels][stride]; channel c is designed as ftype[c] (1 lp, 2 hp, 3 bp)
// at f0[c]/fs with unit gain and zero state, then run over n samples.  One filter object per channel,
// channels spread over `threads` std::threads -- the all-cores CPU baseline of SURVEY 8(d).
int sdsp_ref_iir_bank_run(double *data, size_t channels, size_t n, size_t stride, int sections, int kind,
                          const int *ftype, const double *f0, double fs, double q, int threads)
{
    std::atomic<int> rc{ 0 };
    parallel_for(channels, threads, [&](size_t c) {
        filter_iface *f = static_cast<filter_iface *>(sdsp_ref_iir_create(sections, kind));
        if (!f) {
            rc = -1;
            return;
        }
        int r = -1;
        if (ftype[c] == 1)
            r = f->set_lp(f0[c], fs, 1.0);
        else if (ftype[c] == 2)
            r = f->set_hp(f0[c], fs, 1.0);
        else if (ftype[c] == 3)
            r = f->set_bp(f0[c], fs, q, 1.0);
        if (r != 0)
            rc = -1;
        else
            f->process(data + c * stride, n);
        delete f;
    });
    return rc;
}
}
