// Dumps the compile-time tables of sdsp/fft.h as raw bytes on stdout.  Compiled twice by tests/test_header_tables.py: once against
// this repo's drop-in header (include/), once -- where it is present -- against the reference's own header; the two outputs
// must be identical byte for byte (reference include/sdsp/fft.h:12-43, 54-119, 148-256).
#include <cstdio>
#include <cstdint>

#include "sdsp/fft.h"

template <typename A>
static void put(const A &a)
{
    fwrite(a.data(), sizeof(a[0]), a.size(), stdout);
}

template <size_t N>
static void dump()
{
    static constexpr auto cosines = sdsp::calc_trigs<N, sdsp::cosine_calculator>();
    static constexpr auto sines = sdsp::calc_trigs<N, sdsp::sine_calculator>();
    static constexpr auto naive = sdsp::calc_trigs_naive<N, sdsp::cosine_calculator>();
    static constexpr auto wf = sdsp::calc_wCoeffs<N, sdsp::forward_fft>();
    static constexpr auto wr = sdsp::calc_wCoeffs<N, sdsp::reverse_fft>();
    static constexpr auto s2 = sdsp::calc_swap_lookup<N, 2>();
    put(cosines);
    put(sines);
    put(naive);
    put(wf);
    put(wr);
    put(s2);
    if constexpr (sdsp::isPowerOf4(N)) {
        static constexpr auto s4 = sdsp::calc_swap_lookup<N, 4>();
        put(s4);
    }
    const uint32_t ints[4] = { sdsp::log2(N), sdsp::log4(N), sdsp::isPowerOf2(N), sdsp::isPowerOf4(N) };
    fwrite(ints, sizeof ints, 1, stdout);
    for (unsigned i = 0; i < N; i += 7) {
        const uint32_t r[2] = { sdsp::digit_reverse<N, 2>(i), sdsp::isPowerOf4(N) ? sdsp::digit_reverse<N, 4>(i) : 0u };
        fwrite(r, sizeof r, 1, stdout);
    }
}

int main()
{
    dump<2>();
    dump<4>();
    dump<8>();
    dump<64>();
    dump<128>();
    dump<256>();
    dump<1024>();
    return 0;
}
