// fft.cu -- batched complex FFT for sm_100a: kernels, plans and the C-ABI entry points.
//
// Replaces sdsp::fft_radix2 / sdsp::fft_radix4 (reference include/sdsp/fft.h:258-299, :301-360) and the
// compile-time tables behind them (calc_wCoeffs :197-214, calc_swap_lookup :238-256), for batches of
// frames resident in HBM.  See fft_core.cuh for the factorisation and DESIGN.md for the roofline.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <cooperative_groups.h>

#include <cuda_runtime.h>

#include "fft_core.cuh"
#include "host_stage.h"
#include "tma_ptx.cuh"

namespace sdsp_b200
{
// =================================================================================================
// twiddles.  exp(-2*pi*i*num/den) for den a power of two, evaluated in long double on the first
// octant only and mirrored, so the table has the exact symmetries (and exact 0 / +-1 / equal
// cos = sin at 45 degrees) that the reference's quarter-wave construction has (fft.h:148-194).
static void unit_root(uint64_t num, uint64_t den, long double &re, long double &im)
{
    num %= den;
    if (den < 8) { // den in {1,2,4}: all roots are exact
        const int q = (int)(num * 4 / den);
        static const int cr[4] = { 1, 0, -1, 0 }, ci[4] = { 0, -1, 0, 1 };
        re = cr[q];
        im = ci[q];
        return;
    }
    const uint64_t oct = den / 8;
    const uint64_t o = num / oct; // octant 0..7
    const uint64_t r = num % oct;
    // angle inside the octant, folded to [0, pi/4]
    const bool fold = (o & 1) != 0;
    const uint64_t rr = fold ? (oct - r) : r;
    const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)rr / (long double)den;
    long double c = cosl(ang), s = sinl(ang);
    if (rr == 0) {
        c = 1.0L;
        s = 0.0L;
    } else if (rr == oct) {
        c = s = 0.70710678118654752440084436210484903928L;
    }
    if (fold) {
        const long double t = c;
        c = s;
        s = t;
    }
    // now (c, s) = (cos, sin) of the angle reduced into the first quadrant position within quadrant o/2
    long double cq, sq; // cos/sin of full angle theta = 2*pi*num/den
    switch (o / 2) {
    case 0: cq = c; sq = s; break;
    case 1: cq = -s; sq = c; break;
    case 2: cq = -c; sq = -s; break;
    default: cq = s; sq = -c; break;
    }
    re = cq;
    im = -sq;
}

template <typename T>
static void build_twiddles(int n, const int *radix, int npass, std::vector<cplx<T>> &out)
{
    out.clear();
    int pp = 1;
    for (int p = 0; p + 1 < npass; p++) {
        const int r = radix[p];
        const int np = n / pp;
        const int m_range = np / r;
        for (int k = 1; k < r; k++)
            for (int m = 0; m < m_range; m++) {
                long double re, im;
                unit_root((uint64_t)m * (uint64_t)k, (uint64_t)np, re, im);
                out.push_back(cplx<T>{ (T)re, (T)im });
            }
        pp *= r;
    }
}

// =================================================================================================
// the single-CTA kernel: FPC frames per CTA, TPF threads per frame, E points per thread
template <typename T>
__device__ __forceinline__ cplx<T> ld_stream(const cplx<T> *p)
{
    return *p;
}
template <typename T>
__device__ __forceinline__ void st_stream(cplx<T> *p, cplx<T> v)
{
    *p = v;
}

__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes) // bytes: multiple of 16
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// DB = false: one exchange buffer, a barrier after every write and after every read that is followed by a
//   write (four per 3-pass frame).
// DB = true: two buffers used alternately; a pass reads one and writes the other, so only the barrier after
//   each write remains (two per 3-pass frame).  fs1 = second buffer.
// BAR = 0: the barriers are __syncthreads(); BAR > 0: named barrier BAR over THREADS threads (the CTA holds other warps too).
template <int BAR, int THREADS>
__device__ __forceinline__ void cta_sync()
{
    if constexpr (BAR == 0)
        __syncthreads();
    else
        asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(THREADS) : "memory");
}
template <class Cfg, typename T, int THREADS, int MINB, int P, bool DB = false, int BAR = 0>
__device__ __forceinline__ void fft_kernel_passes(cplx<T> (&v)[Cfg::E], cplx<T> *fs, const cplx<T> *__restrict__ tw, int t,
                                                  cplx<T> *fs1 = nullptr)
{
    if constexpr (P < Cfg::NPASS) {
        cplx<T> *rd = (DB && (P % 2 == 0)) ? fs1 : fs; // pass P reads what pass P-1 wrote
        cplx<T> *wr = (DB && (P % 2 == 1)) ? fs1 : fs;
        if constexpr (P > 0) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = rd[fft_read_phys<Cfg>(t, e)];
            if constexpr (P + 1 < Cfg::NPASS && !DB)
                cta_sync<BAR, THREADS>(); // everyone has read before this pass overwrites the exchange buffer
        }
        fft_pass<Cfg, P, T>(v, t, tw);
        if constexpr (P + 1 < Cfg::NPASS) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                wr[fft_out_phys<Cfg, P>(t, e)] = v[e];
            cta_sync<BAR, THREADS>();
            fft_kernel_passes<Cfg, T, THREADS, MINB, P + 1, DB, BAR>(v, fs, tw, t, fs1);
        }
    }
}

// ALIAS variant: the exchange buffer doubles as the staging area of the NEXT group.  Once the last exchange has been read
// (and a barrier has made that true for every thread) the buffer is idle for the rest of the group -- last pass, stores -- so
// each thread starts cp.async copies of exactly the points it will take next (thread-private slots, natural layout) and the
// next iteration begins with shared-memory reads instead of an L2 / HBM round trip.  Same number of barriers per group as the
// plain variant (the barrier that used to end a group now follows the read of the staged points instead).
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
template <class Cfg, typename T, int P, typename Issue>
__device__ __forceinline__ void fft_kernel_passes_alias(cplx<T> (&v)[Cfg::E], cplx<T> *fs, const cplx<T> *__restrict__ tw, int t, Issue &&issue_next)
{
    if constexpr (P < Cfg::NPASS) {
        if constexpr (P > 0) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = fs[fft_read_phys<Cfg>(t, e)];
            __syncthreads(); // everyone has read: the buffer may be overwritten (next exchange, or the staging copies)
        }
        if constexpr (P + 1 == Cfg::NPASS)
            issue_next();
        fft_pass<Cfg, P, T>(v, t, tw);
        if constexpr (P + 1 < Cfg::NPASS) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                fs[fft_out_phys<Cfg, P>(t, e)] = v[e];
            __syncthreads();
            fft_kernel_passes_alias<Cfg, T, P + 1>(v, fs, tw, t, issue_next);
        }
    }
}

template <class Cfg, typename T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    fft_cta_alias_kernel(cplx<T> *__restrict__ data, const cplx<T> *__restrict__ tw, size_t n_frames, int inverse, T scale)
{
    constexpr int FPC = THREADS / Cfg::TPF;
    static_assert(THREADS % Cfg::TPF == 0 && FPC >= 1 && Cfg::NPASS > 1, "block must hold whole frames; needs an exchange buffer");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *smem = reinterpret_cast<cplx<T> *>(smem_raw);
    const int fl = threadIdx.x / Cfg::TPF;
    const int t = threadIdx.x % Cfg::TPF;
    cplx<T> *fs = smem + (size_t)fl * Cfg::PADDED_N;
    const size_t groups = (n_frames + FPC - 1) / FPC;
    auto stage = [&](size_t grp) { // this thread's own points of its frame in group `grp`, natural positions t + S e
        const size_t frame = grp * FPC + fl;
        if (grp < groups && frame < n_frames) {
            const cplx<T> *gp = data + frame * (size_t)Cfg::N + t;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++) {
                if constexpr (sizeof(cplx<T>) == 16)
                    cp_async16(fs + t + Cfg::S * e, gp + Cfg::S * e);
                else
                    cp_async8(fs + t + Cfg::S * e, gp + Cfg::S * e);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(blockIdx.x);
    for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const size_t frame = g * FPC + fl;
        const bool active = frame < n_frames;
        cplx<T> *gp = data + frame * (size_t)Cfg::N + t;
        cplx<T> v[Cfg::E];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int e = 0; e < Cfg::E; e++)
            v[e] = active ? fs[t + Cfg::S * e] : cplx<T>{ 0, 0 };
        __syncthreads(); // everyone holds its points: the first exchange may overwrite the staging area
        if (inverse) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = cplx<T>{ v[e].y, v[e].x };
        }
        fft_kernel_passes_alias<Cfg, T, 0>(v, fs, tw, t, [&] { stage(g + gridDim.x); });
        if (inverse) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = cplx<T>{ v[e].y * scale, v[e].x * scale };
        }
        if (active) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                st_stream(gp + Cfg::S * e, v[e]);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <class Cfg, typename T, int THREADS, int MINB, bool DB = false>
__global__ void __launch_bounds__(THREADS, MINB)
    fft_cta_kernel(cplx<T> *__restrict__ data, const T *__restrict__ real_in, const cplx<T> *__restrict__ tw, size_t n_frames, int inverse,
                   T scale, int prefetch)
{
    constexpr int FPC = THREADS / Cfg::TPF;
    static_assert(THREADS % Cfg::TPF == 0 && FPC >= 1, "block must hold whole frames");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *smem = reinterpret_cast<cplx<T> *>(smem_raw);
    const int fl = threadIdx.x / Cfg::TPF;
    const int t = threadIdx.x % Cfg::TPF;
    cplx<T> *fs = smem + (size_t)fl * Cfg::PADDED_N;
    cplx<T> *fs1 = fs + (size_t)FPC * Cfg::PADDED_N; // second exchange buffer (DB only)
    const size_t groups = (n_frames + FPC - 1) / FPC;

    for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const size_t frame = g * FPC + fl;
        const bool active = frame < n_frames;
        // the frames this CTA takes next: pulled into L2 while this group is transformed (one bulk prefetch, no registers)
        if (prefetch && threadIdx.x == 0 && g + gridDim.x < groups) {
            const size_t nf = (g + gridDim.x) * FPC;
            const size_t cnt = (n_frames - nf) < (size_t)FPC ? (n_frames - nf) : (size_t)FPC;
            if (real_in)
                prefetch_l2_bulk(real_in + nf * (size_t)Cfg::N, (unsigned)(cnt * Cfg::N * sizeof(T)));
            else
                prefetch_l2_bulk(data + nf * (size_t)Cfg::N, (unsigned)(cnt * Cfg::N * sizeof(cplx<T>)));
        }
        cplx<T> *gp = data + frame * (size_t)Cfg::N + t;
        cplx<T> v[Cfg::E];
        if (active && real_in) { // real samples (imaginary part zero, as the reference's callers fill their arrays), out of place
            const T *rp = real_in + frame * (size_t)Cfg::N + t;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = cplx<T>{ rp[Cfg::S * e], (T)0 };
        } else if (active) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = ld_stream(gp + Cfg::S * e);
        } else {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = cplx<T>{ 0, 0 };
        }
        if (inverse) { // IDFT(x) = swap(DFT(swap(x))) / N
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = cplx<T>{ v[e].y, v[e].x };
        }
        fft_kernel_passes<Cfg, T, THREADS, MINB, 0, DB>(v, fs, tw, t, fs1);
        if (inverse) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = cplx<T>{ v[e].y * scale, v[e].x * scale };
        }
        if (active) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                st_stream(gp + Cfg::S * e, v[e]);
        }
        // single buffer: the last pass has read the exchange buffer before the next group writes it.  With two
        // buffers the barriers inside the next frame already order every reuse (a buffer is rewritten only after a
        // barrier that all readers of its previous contents have passed), except when one buffer serves both the
        // first write and the last read (even number of exchanges)
        if constexpr (Cfg::NPASS > 1 && (!DB || (Cfg::NPASS - 1) % 2 == 1))
            __syncthreads();
    }
}


// =================================================================================================
// Real frames in, HALF spectra out (sdsp_b200_fft_exec_r2c): a real frame of N = 2M samples is read as M complex numbers
// z[n] = x[2n] + i x[2n + 1] -- the frame as it lies in memory -- transformed by the M-point kernel code above, and the bins
// k = 0 .. M of the N-point spectrum come out of one more exchange (every bin needs its mirror Z[M - k]):
//   X[k] = (Z[k] + conj Z[M - k]) / 2  +  W_N^k (-i) (Z[k] - conj Z[M - k]) / 2,      X[M] = Re Z[0] - Im Z[0].
// 4 bytes in and 4 bytes out per real sample against 8 + 8 for the reference's calling convention (real part filled, imaginary
// part zero: test/testFFT.cpp:24, :86) and 4 + 8 for sdsp_b200_fft_exec_real.  W_N^k = W_N^t W_N^(S e) for k = t + S e: the first
// factor is a per-thread constant (one table look-up per frame), the second is W_(2E)^e, a compile-time index into a 64th-root
// table in constant memory.  The bins above M are the conjugates of those below (not written).
__constant__ float2 c_w64_f32[32];
__constant__ double2 c_w64_f64[32];
template <typename T>
__device__ __forceinline__ cplx<T> w64(int j)
{
    if constexpr (sizeof(T) == 4)
        return cplx<T>{ c_w64_f32[j].x, c_w64_f32[j].y };
    else
        return cplx<T>{ c_w64_f64[j].x, c_w64_f64[j].y };
}

template <class Cfg, typename T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    fft_r2c_kernel(const cplx<T> *__restrict__ in, cplx<T> *__restrict__ out, const cplx<T> *__restrict__ tw, const cplx<T> *__restrict__ tw_n,
                   size_t n_frames, int prefetch)
{
    constexpr int FPC = THREADS / Cfg::TPF, M = Cfg::N;
    static_assert(THREADS % Cfg::TPF == 0 && FPC >= 1, "block must hold whole frames");
    static_assert(32 % Cfg::E == 0, "the 64th-root table covers 2, 4, 8, 16 or 32 points per thread");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *smem = reinterpret_cast<cplx<T> *>(smem_raw);
    const int fl = threadIdx.x / Cfg::TPF;
    const int t = threadIdx.x % Cfg::TPF;
    cplx<T> *fs = smem + (size_t)fl * Cfg::PADDED_N;
    const size_t groups = (n_frames + FPC - 1) / FPC;

    for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const size_t frame = g * FPC + fl;
        const bool active = frame < n_frames;
        if (prefetch && threadIdx.x == 0 && g + gridDim.x < groups) {
            const size_t nf = (g + gridDim.x) * FPC;
            const size_t cnt = (n_frames - nf) < (size_t)FPC ? (n_frames - nf) : (size_t)FPC;
            prefetch_l2_bulk(in + nf * (size_t)M, (unsigned)(cnt * M * sizeof(cplx<T>)));
        }
        cplx<T> v[Cfg::E];
        const cplx<T> *gp = in + frame * (size_t)M + t;
#pragma unroll
        for (int e = 0; e < Cfg::E; e++)
            v[e] = active ? ld_stream(gp + Cfg::S * e) : cplx<T>{ 0, 0 };
        fft_kernel_passes<Cfg, T, THREADS, MINB, 0, false>(v, fs, tw, t, nullptr);
        // v[e] = Z[t + S e].  The mirror terms through the exchange buffer in natural order (reads run downwards: no padding needed)
        if constexpr (Cfg::NPASS > 1)
            __syncthreads(); // everyone has read the last exchange
#pragma unroll
        for (int e = 0; e < Cfg::E; e++)
            fs[t + Cfg::S * e] = v[e];
        __syncthreads();
        if (active) {
            cplx<T> *op = out + frame * (size_t)(M + 1);
            const cplx<T> wt = tw_n[t]; // W_N^t (re-read per frame: two or four registers less across the passes)
#pragma unroll
            for (int e = 0; e < Cfg::E; e++) {
                const int k = t + Cfg::S * e;
                const cplx<T> zm = fs[(M - k) & (M - 1)]; // k = 0 mirrors into itself
                const cplx<T> w = e == 0 ? wt : cmul(wt, w64<T>(e * (32 / Cfg::E)));
                st_stream(op + k, r2c_bin(v[e], zm, w));
            }
            if (t == 0)
                st_stream(op + M, cplx<T>{ v[0].x - v[0].y, (T)0 });
        }
        __syncthreads(); // the mirror terms have been read before the next group's first exchange
    }
}


// The way back (sdsp_b200_fft_exec_c2r, reverse plans): bins 0 .. M of a real frame's spectrum in, the n = 2M real samples out, 1/n
// included (reverse_fft::ScaleValues, reference fft.h:128-132).  The separation runs backwards at the first load -- both terms come
// from global memory, so it needs no exchange:
//   Z[k] = 1/2 [ (X[k] + conj X[M - k]) + i conj(W_n^k) (X[k] - conj X[M - k]) ],   z = IDFT_M(Z),   x[2j] + i x[2j + 1] = z[j]
// with the inverse transform done by the forward passes on swapped components (IDFT(x) = swap(DFT(swap x)) / M, exact).
template <class Cfg, typename T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    fft_c2r_kernel(const cplx<T> *__restrict__ in, cplx<T> *__restrict__ out, const cplx<T> *__restrict__ tw, const cplx<T> *__restrict__ tw_n,
                   size_t n_frames, int prefetch)
{
    constexpr int FPC = THREADS / Cfg::TPF, M = Cfg::N;
    static_assert(THREADS % Cfg::TPF == 0 && FPC >= 1, "block must hold whole frames");
    static_assert(32 % Cfg::E == 0, "the 64th-root table covers 2, 4, 8, 16 or 32 points per thread");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *smem = reinterpret_cast<cplx<T> *>(smem_raw);
    const int fl = threadIdx.x / Cfg::TPF;
    const int t = threadIdx.x % Cfg::TPF;
    cplx<T> *fs = smem + (size_t)fl * Cfg::PADDED_N;
    const size_t groups = (n_frames + FPC - 1) / FPC;
    const T half_over_m = (T)(0.5 / (double)M);

    for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const size_t frame = g * FPC + fl;
        const bool active = frame < n_frames;
        // the half spectra this CTA takes next, pulled into L2 meanwhile (frames M + 1 bins apart: the 16-byte-aligned inside of the range)
        if (prefetch && threadIdx.x == 0 && g + gridDim.x < groups) {
            const size_t nf = (g + gridDim.x) * FPC;
            const size_t cnt = (n_frames - nf) < (size_t)FPC ? (n_frames - nf) : (size_t)FPC;
            size_t lo = nf * (size_t)(M + 1) * sizeof(cplx<T>), hi = (nf + cnt) * (size_t)(M + 1) * sizeof(cplx<T>);
            lo = (lo + 15) & ~(size_t)15;
            hi &= ~(size_t)15;
            if (hi > lo)
                prefetch_l2_bulk(reinterpret_cast<const char *>(in) + lo, (unsigned)(hi - lo));
        }
        cplx<T> v[Cfg::E];
        if (active) {
            const cplx<T> *ip = in + frame * (size_t)(M + 1);
            const cplx<T> wt = tw_n[t];
#pragma unroll
            for (int e = 0; e < Cfg::E; e++) {
                const int k = t + Cfg::S * e;
                const cplx<T> a = ld_stream(ip + k), xm = ld_stream(ip + (M - k));
                const cplx<T> w = e == 0 ? wt : cmul(wt, w64<T>(e * (32 / Cfg::E)));
                const cplx<T> z = c2r_bin(a, xm, w);
                v[e] = cplx<T>{ z.y, z.x }; // swapped for the inverse
            }
        } else {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = cplx<T>{ 0, 0 };
        }
        fft_kernel_passes<Cfg, T, THREADS, MINB, 0, false>(v, fs, tw, t, nullptr);
        if (active) {
            cplx<T> *op = out + frame * (size_t)M + t;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                st_stream(op + Cfg::S * e, cplx<T>{ v[e].y * half_over_m, v[e].x * half_over_m });
        }
        if constexpr (Cfg::NPASS > 1)
            __syncthreads(); // the last pass has read the exchange buffer before the next group writes it
    }
}

// =================================================================================================
// digit reversal on the device (reference fft.h:217-236).  Base 2: bit reversal of the log2(n) low
// bits; base 4: the same with the two bits of every digit kept in order.
__device__ __forceinline__ uint32_t digit_reverse_dev(uint32_t i, int log2n, uint32_t base)
{
    uint32_t r = __brev(i) >> (32 - log2n);
    if (base == 4)
        r = ((r & 0xAAAAAAAAu) >> 1) | ((r & 0x55555555u) << 1);
    return r;
}

__global__ void digit_reverse_table_kernel(uint32_t *out, uint32_t n, int log2n, uint32_t base, int half_table)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    uint32_t r = digit_reverse_dev(i, log2n, base);
    // calc_swap_lookup (fft.h:249-254): of every pair only the lower index keeps its partner, so a
    // linear sweep swaps once.  Entries 0 and n-1 are fixed points anyway.
    if (half_table && r < i)
        r = i;
    out[i] = r;
}

template <typename T>
__global__ void digit_reverse_permute_kernel(cplx<T> *data, uint32_t n, int log2n, uint32_t base, size_t n_frames)
{
    // one thread per pair (i, rev(i)) with i < rev(i): an in-place swap touches every element once
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = n_frames * (size_t)n;
    if (gid >= total)
        return;
    const size_t frame = gid / n;
    const uint32_t i = (uint32_t)(gid % n);
    const uint32_t r = digit_reverse_dev(i, log2n, base);
    if (r > i) {
        cplx<T> *f = data + frame * n;
        const cplx<T> a = f[i], b = f[r];
        f[i] = b;
        f[r] = a;
    }
}

// =================================================================================================
// plans
struct FftPlan;
typedef int (*fft_launch_fn)(const FftPlan &, void *data, const void *real_in, size_t n_frames, cudaStream_t stream);
typedef void (*fft_emulate_fn)(void *frame, const void *tw, bool inverse);
typedef int (*fft_r2c_launch_fn)(const FftPlan &, const void *real_in, void *half_out, size_t n_frames, cudaStream_t stream);

struct FftPlan {
    uint32_t n = 0;
    int radix = 2, precision = 0, direction = 0, device = 0;
    int npass = 0, radices[4] = { 1, 1, 1, 1 }, e = 0, threads = 0, frames_per_cta = 0, min_blocks = 0;
    size_t smem_bytes = 0;
    int ctas_per_sm = 0, sm_count = 0;
    bool double_buffered = false;
    bool staged = false; // next group staged into the exchange buffer by cp.async (fft_cta_alias_kernel)
    bool two_slot = false; // data-mover kernel with two tile slots that double as exchange buffers (fft_fused_tma2_kernel)
    int real64k_ctas = 0;  // > 0: forward real-input frames of 65536 points take fft_real64k_kernel (resident CTAs per SM)
    size_t real64k_smem = 0;
    // half-spectrum path (sdsp_b200_fft_exec_r2c): an n/2-point complex transform plus a separation step; set up at the first call
    bool r2c_ready = false, r2c_through_full = false;
    void *d_half_tmp = nullptr; // full-spectrum temporary of the sizes without a direct half-spectrum kernel
    size_t half_tmp_bytes = 0;
    fft_r2c_launch_fn r2c_launch = nullptr;
    void *d_r2c_tw = nullptr, *d_r2c_twn = nullptr;
    size_t r2c_smem = 0;
    int r2c_ctas = 0, r2c_threads = 0, r2c_fpc = 0, r2c_e = 0, r2c_npass = 0;
    void *d_tw = nullptr;
    size_t tw_bytes = 0;
    fft_launch_fn launch = nullptr;
    // multi-pass ("four-step") path for frames larger than one CTA can hold: n = n1 * 256
    bool large = false;
    bool cluster = false; // n = 65536: one frame per 8-CTA cluster, single pass over HBM
    bool fused = false;   // n = 65536: column and row tiles in one persistent kernel, intermediate in an L2-resident ring
    void *d_fused_counters = nullptr;
    size_t fused_counter_bytes = 0;
    int cluster_slots = 0, cluster_ctas = 0;
    int n1 = 0;
    void *d_tw_cols = nullptr, *d_tw_rows = nullptr, *d_tw_hi = nullptr, *d_tw_lo = nullptr;
    void *d_scratch = nullptr;
    size_t scratch_frames = 0;
    int large_threads_a = 0;
    size_t large_smem_a = 0, large_smem_b = 0;
    // host staging (ptr_kind == HOST)
    void *d_stage = nullptr;
    size_t stage_bytes = 0;
    HostStage host;
    std::mutex mu;
};

static bool fft_prefetch_enabled() // SDSP_B200_FFT_PREFETCH=0 switches the L2 prefetch of the next frames off (comparison aid)
{
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("SDSP_B200_FFT_PREFETCH");
        on = e ? atoi(e) : 1;
    }
    return on != 0;
}

template <class Cfg, typename T, int THREADS, int MINB, bool DB>
static int launch_cta(const FftPlan &p, void *data, const void *real_in, size_t n_frames, cudaStream_t stream)
{
    constexpr int FPC = THREADS / Cfg::TPF;
    const size_t groups = (n_frames + FPC - 1) / FPC;
    if (groups == 0)
        return SDSP_B200_OK;
    const size_t resident = (size_t)p.sm_count * (size_t)p.ctas_per_sm;
    // persistent-style grid: a whole number of waves, each CTA strides over frame groups
    size_t grid = groups < resident * 4 ? groups : resident * 4;
    const T scale = (T)(1.0 / (double)Cfg::N);
    fft_cta_kernel<Cfg, T, THREADS, MINB, DB><<<(unsigned)grid, THREADS, p.smem_bytes, stream>>>(
        reinterpret_cast<cplx<T> *>(data), static_cast<const T *>(real_in), reinterpret_cast<const cplx<T> *>(p.d_tw), n_frames,
        p.direction == SDSP_B200_REVERSE ? 1 : 0, scale, fft_prefetch_enabled() && (Cfg::N * sizeof(T)) % 16 == 0 && reinterpret_cast<uintptr_t>(data) % 16 == 0 &&
                reinterpret_cast<uintptr_t>(real_in) % 16 == 0 ?
            1 :
            0);
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

template <class Cfg, typename T, int THREADS, int MINB>
static int launch_cta_alias(const FftPlan &p, void *data, const void *real_in, size_t n_frames, cudaStream_t stream)
{
    constexpr int FPC = THREADS / Cfg::TPF;
    // real input and buffers that are not 16-byte aligned take the plain kernel (same arithmetic, same bits)
    if (real_in || reinterpret_cast<uintptr_t>(data) % 16 != 0)
        return launch_cta<Cfg, T, THREADS, MINB, false>(p, data, real_in, n_frames, stream);
    const size_t groups = (n_frames + FPC - 1) / FPC;
    if (groups == 0)
        return SDSP_B200_OK;
    const size_t resident = (size_t)p.sm_count * (size_t)p.ctas_per_sm;
    const size_t grid = groups < resident * 4 ? groups : resident * 4;
    fft_cta_alias_kernel<Cfg, T, THREADS, MINB><<<(unsigned)grid, THREADS, p.smem_bytes, stream>>>(
        reinterpret_cast<cplx<T> *>(data), reinterpret_cast<const cplx<T> *>(p.d_tw), n_frames, p.direction == SDSP_B200_REVERSE ? 1 : 0,
        (T)(1.0 / (double)Cfg::N));
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

template <class Cfg, typename T, int THREADS, int MINB, bool DB = false>
static int setup_cta(FftPlan &p)
{
    constexpr int FPC = THREADS / Cfg::TPF;
    p.npass = Cfg::NPASS;
    p.radices[0] = Cfg::R0;
    p.radices[1] = Cfg::R1;
    p.radices[2] = Cfg::R2;
    p.radices[3] = Cfg::R3;
    p.e = Cfg::E;
    p.threads = THREADS;
    p.frames_per_cta = FPC;
    p.min_blocks = MINB;
    p.smem_bytes = Cfg::NPASS > 1 ? (size_t)FPC * Cfg::PADDED_N * sizeof(cplx<T>) * (DB ? 2 : 1) : 0;
    p.double_buffered = DB;
    auto kern = fft_cta_kernel<Cfg, T, THREADS, MINB, DB>;
    if (p.smem_bytes > 48 * 1024)
        SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    int occ = 0;
    SDSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, p.smem_bytes));
    if (occ < 1)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "fft kernel for n=%u does not fit on an SM", p.n);
    p.ctas_per_sm = occ;
    p.launch = &launch_cta<Cfg, T, THREADS, MINB, DB>;
    std::vector<cplx<T>> tw;
    build_twiddles<T>(Cfg::N, p.radices, Cfg::NPASS, tw);
    p.tw_bytes = tw.size() * sizeof(cplx<T>);
    if (p.tw_bytes) {
        SDSP_CUDA(cudaMalloc(&p.d_tw, p.tw_bytes));
        SDSP_CUDA(cudaMemcpy(p.d_tw, tw.data(), p.tw_bytes, cudaMemcpyHostToDevice));
    }
    return SDSP_B200_OK;
}

// switch a plan set up by setup_cta<..., DB = false> to the alias-staging kernel (same shared memory, same occupancy)
template <class Cfg, typename T, int THREADS, int MINB>
static int enable_alias(FftPlan &p)
{
    auto kern = fft_cta_alias_kernel<Cfg, T, THREADS, MINB>;
    if (p.smem_bytes > 48 * 1024)
        SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    p.launch = &launch_cta_alias<Cfg, T, THREADS, MINB>;
    p.staged = true;
    return SDSP_B200_OK;
}

template <class Cfg, typename T>
static void emulate_cfg(void *frame, bool inverse)
{
    int radices[4] = { Cfg::R0, Cfg::R1, Cfg::R2, Cfg::R3 };
    std::vector<cplx<T>> tw;
    build_twiddles<T>(Cfg::N, radices, Cfg::NPASS, tw);
    tw.push_back(cplx<T>{ 1, 0 });
    fft_emulate_frame<Cfg, T>(reinterpret_cast<cplx<T> *>(frame), tw.data(), inverse, (T)(1.0 / (double)Cfg::N));
}

// =================================================================================================
// frames larger than one CTA can hold: n = n1 * 256, Cooley-Tukey with n = 256*a + b, k = k1 + n1*k2:
//   X[k1 + n1 k2] = sum_b W_256^(b k2) [ W_n^(b k1) sum_a x[256 a + b] W_n1^(a k1) ]
// pass A ("columns"): for 16 adjacent b at a time the n1-point transforms over a (stride 256), times the
//   twiddle W_n^(b k1), written to a scratch frame at [k1][b] -- reads and writes are 128-byte runs;
// pass B ("rows"): for 16 adjacent k1 at a time the 256-point transforms over b (contiguous), the results
//   transposed through shared memory so that the 16 outputs X[k1.. + n1 k2] leave as one 128-byte run.
// The scratch holds only a slab of frames small enough to stay in the 126 MB L2 between the two
// passes, so HBM sees each frame once in and once out.  W_n^(b k1) is the product of two 256-/n1-entry
// tables (W_n^x = W_n1^(x >> 8) * W_n^(x & 255)).
constexpr int LARGE_COLS = 16; // columns / rows per CTA

// The 16 inter-transform twiddles a thread needs, W_n^(b*k1) with k1 = t + S*e, form a geometric sequence in e:
// W_n^(b t) * (W_n^(b S))^e.  Looking each one up costs two table loads whose addresses diverge across the warp
// (b varies by lane) -- 32 divergent loads per thread.  Instead seven terms are looked up (each exact to one
// rounding) and the rest are one product away:  w_e = qh[e / 4] * ql[e % 4],
//   qh[i] = W^(b (t + 4 i S)),  ql[j] = W^(j b S).
template <typename T>
struct TwiddleSeq {
    cplx<T> qh[4], ql[3];
    __device__ __forceinline__ TwiddleSeq(unsigned x0, unsigned step, const cplx<T> *__restrict__ hi, const cplx<T> *__restrict__ lo)
    {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned x = x0 + 4u * (unsigned)i * step;
            qh[i] = cmul(hi[x >> 8], lo[x & 255u]);
        }
#pragma unroll
        for (int j = 1; j < 4; j++) {
            const unsigned x = (unsigned)j * step;
            ql[j - 1] = cmul(hi[x >> 8], lo[x & 255u]);
        }
    }
    __device__ __forceinline__ cplx<T> get(int e) const // e is a compile-time constant after unrolling
    {
        return (e & 3) == 0 ? qh[e >> 2] : cmul(qh[e >> 2], ql[(e & 3) - 1]);
    }
};

// The same sixteen factors w_e = w0 * r^e from their first term and ratio alone (no table look-ups: the look-ups of TwiddleSeq
// are lane-divergent, up to 32 shared-memory wavefronts each -- a third of all the shared-memory wavefronts of the data-mover
// kernel, profiles/r01_ncu_fft65536_f32_fused_tma_v2.txt).  Eight more complex products per item; the deepest product chain
// is w0 * r^12 * r^3: w_e is good to ~8 ulp, two orders of magnitude inside the fp32 parity bound.
template <typename T>
struct TwiddleGeo {
    cplx<T> qh[4], ql[3];
    __device__ __forceinline__ TwiddleGeo(cplx<T> w0, cplx<T> r)
    {
        ql[0] = r;
        ql[1] = cmul(r, r);
        ql[2] = cmul(ql[1], r);
        const cplx<T> r4 = cmul(ql[1], ql[1]), r8 = cmul(r4, r4);
        qh[0] = w0;
        qh[1] = cmul(w0, r4);
        qh[2] = cmul(w0, r8);
        qh[3] = cmul(qh[2], r4);
    }
    __device__ __forceinline__ cplx<T> get(int e) const
    {
        return (e & 3) == 0 ? qh[e >> 2] : cmul(qh[e >> 2], ql[(e & 3) - 1]);
    }
};

template <class Cfg>
struct LargeStride { // odd frame pitch: the 16 frames a warp touches together land in different banks
    static constexpr int value = Cfg::PADDED_N + 1;
};

template <class Cfg, typename T, int THREADS>
__global__ void __launch_bounds__(THREADS)
    fft_large_cols_kernel(const cplx<T> *__restrict__ in, cplx<T> *__restrict__ out, const cplx<T> *__restrict__ tw,
                          const cplx<T> *__restrict__ tw_hi, const cplx<T> *__restrict__ tw_lo, size_t n_frames, int inverse)
{
    constexpr int N1 = Cfg::N, N2 = 256;
    static_assert(THREADS == LARGE_COLS * Cfg::TPF, "16 columns per CTA");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *smem = reinterpret_cast<cplx<T> *>(smem_raw);
    const int col = threadIdx.x & (LARGE_COLS - 1);
    const int t = threadIdx.x / LARGE_COLS;
    cplx<T> *fs = smem + (size_t)col * LargeStride<Cfg>::value;
    constexpr int TILES = N2 / LARGE_COLS; // column tiles per frame
    const size_t work = n_frames * TILES;
    for (size_t w = blockIdx.x; w < work; w += gridDim.x) {
        const size_t frame = w / TILES;
        const int c0 = (int)(w % TILES) * LARGE_COLS;
        const cplx<T> *gp = in + frame * ((size_t)N1 * N2) + c0 + col;
        cplx<T> v[Cfg::E];
#pragma unroll
        for (int e = 0; e < Cfg::E; e++)
            v[e] = gp[(size_t)(t + Cfg::S * e) * N2];
        if (inverse) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = cplx<T>{ v[e].y, v[e].x };
        }
        fft_kernel_passes<Cfg, T, THREADS, 1, 0>(v, fs, tw, t);
        cplx<T> *op = out + frame * ((size_t)N1 * N2) + c0 + col;
        const unsigned b = (unsigned)(c0 + col);
        static_assert(Cfg::E == 16, "TwiddleSeq covers 16 points per thread");
        const TwiddleSeq<T> wseq(b * (unsigned)t, b * (unsigned)Cfg::S, tw_hi, tw_lo);
#pragma unroll
        for (int e = 0; e < Cfg::E; e++) {
            const unsigned k1 = (unsigned)(t + Cfg::S * e);
            op[(size_t)k1 * N2] = cmul(v[e], wseq.get(e));
        }
        if constexpr (Cfg::NPASS > 1)
            __syncthreads();
    }
}

template <class Cfg, typename T, int THREADS>
__global__ void __launch_bounds__(THREADS)
    fft_large_rows_kernel(const cplx<T> *__restrict__ in, cplx<T> *__restrict__ out, const cplx<T> *__restrict__ tw, size_t n_frames,
                          int n1, int inverse, T scale)
{
    static_assert(Cfg::N == 256 && Cfg::TPF == 16 && THREADS == 256, "16 rows of 256 points per CTA");
    constexpr int N2 = 256;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *smem = reinterpret_cast<cplx<T> *>(smem_raw);
    const int fl = threadIdx.x >> 4, t = threadIdx.x & 15;
    cplx<T> *fs = smem + (size_t)fl * LargeStride<Cfg>::value;
    const int tiles = n1 / LARGE_COLS;
    const size_t work = n_frames * (size_t)tiles;
    for (size_t w = blockIdx.x; w < work; w += gridDim.x) {
        const size_t frame = w / tiles;
        const int r0 = (int)(w % tiles) * LARGE_COLS;
        const cplx<T> *gp = in + frame * ((size_t)n1 * N2) + (size_t)(r0 + fl) * N2 + t;
        cplx<T> v[Cfg::E];
#pragma unroll
        for (int e = 0; e < Cfg::E; e++)
            v[e] = gp[Cfg::S * e];
        fft_kernel_passes<Cfg, T, THREADS, 1, 0>(v, fs, tw, t);
        if (inverse) {
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = cplx<T>{ v[e].y * scale, v[e].x * scale };
        }
        __syncthreads(); // last pass has read the exchange buffer
#pragma unroll
        for (int e = 0; e < Cfg::E; e++)
            fs[Cfg::pad(t + Cfg::S * e)] = v[e]; // natural order: slot e of thread t is k2 = t + 16 e
        __syncthreads();
        // transposed read: 16 consecutive lanes take the same k2 of 16 consecutive rows
        const int rr = threadIdx.x & 15, q = threadIdx.x >> 4;
        const cplx<T> *rs = smem + (size_t)rr * LargeStride<Cfg>::value;
        cplx<T> *op = out + frame * ((size_t)n1 * N2) + r0 + rr;
#pragma unroll
        for (int e = 0; e < Cfg::E; e++) {
            const int k2 = q + 16 * e;
            op[(size_t)k2 * n1] = rs[Cfg::pad(k2)];
        }
        __syncthreads();
    }
}

template <class CfgA, typename T>
static int launch_large(const FftPlan &p, void *data, const void *real_in, size_t n_frames, cudaStream_t stream)
{
    if (real_in)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "fft: real-input frames are not built for the two-kernel path (n=%u)", p.n);
    using CfgB = FftCfg<256, 16, 16, 16>;
    constexpr int THREADS_A = LARGE_COLS * CfgA::TPF;
    cplx<T> *d = reinterpret_cast<cplx<T> *>(data);
    cplx<T> *scratch = reinterpret_cast<cplx<T> *>(p.d_scratch);
    const int inverse = p.direction == SDSP_B200_REVERSE ? 1 : 0;
    const T scale = (T)(1.0 / (double)p.n);
    const size_t grid_cap = (size_t)p.sm_count * 8;
    for (size_t f0 = 0; f0 < n_frames; f0 += p.scratch_frames) {
        const size_t cnt = (n_frames - f0) < p.scratch_frames ? (n_frames - f0) : p.scratch_frames;
        cplx<T> *slab = d + f0 * (size_t)p.n;
        const size_t work_a = cnt * (256 / LARGE_COLS), work_b = cnt * (size_t)(p.n1 / LARGE_COLS);
        fft_large_cols_kernel<CfgA, T, THREADS_A><<<(unsigned)(work_a < grid_cap ? work_a : grid_cap), THREADS_A, p.large_smem_a, stream>>>(
            slab, scratch, reinterpret_cast<const cplx<T> *>(p.d_tw_cols), reinterpret_cast<const cplx<T> *>(p.d_tw_hi),
            reinterpret_cast<const cplx<T> *>(p.d_tw_lo), cnt, inverse);
        fft_large_rows_kernel<CfgB, T, 256><<<(unsigned)(work_b < grid_cap ? work_b : grid_cap), 256, p.large_smem_b, stream>>>(
            scratch, slab, reinterpret_cast<const cplx<T> *>(p.d_tw_rows), cnt, p.n1, inverse, scale);
    }
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

template <typename T>
static int upload_table(void **dst, const std::vector<cplx<T>> &v)
{
    SDSP_CUDA(cudaMalloc(dst, v.size() * sizeof(cplx<T>)));
    SDSP_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice));
    return SDSP_B200_OK;
}

template <class CfgA, typename T>
static int setup_large(FftPlan &p)
{
    using CfgB = FftCfg<256, 16, 16, 16>;
    constexpr int THREADS_A = LARGE_COLS * CfgA::TPF;
    p.large = true;
    p.n1 = CfgA::N;
    p.npass = CfgA::NPASS + CfgB::NPASS;
    p.e = 16;
    p.threads = THREADS_A;
    p.large_threads_a = THREADS_A;
    p.large_smem_a = (size_t)LARGE_COLS * LargeStride<CfgA>::value * sizeof(cplx<T>);
    p.large_smem_b = (size_t)LARGE_COLS * LargeStride<CfgB>::value * sizeof(cplx<T>);
    auto ka = fft_large_cols_kernel<CfgA, T, THREADS_A>;
    auto kb = fft_large_rows_kernel<CfgB, T, 256>;
    if (p.large_smem_a > 48 * 1024)
        SDSP_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.large_smem_a));
    if (p.large_smem_b > 48 * 1024)
        SDSP_CUDA(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.large_smem_b));
    std::vector<cplx<T>> tw;
    int ra[4] = { CfgA::R0, CfgA::R1, CfgA::R2, CfgA::R3 }, rb[4] = { 16, 16, 1, 1 };
    build_twiddles<T>(CfgA::N, ra, CfgA::NPASS, tw);
    tw.push_back(cplx<T>{ 1, 0 });
    int rc = upload_table<T>(&p.d_tw_cols, tw);
    build_twiddles<T>(256, rb, 2, tw);
    if (!rc)
        rc = upload_table<T>(&p.d_tw_rows, tw);
    std::vector<cplx<T>> hi(CfgA::N), lo(256);
    for (int i = 0; i < CfgA::N; i++) {
        long double re, im;
        unit_root((uint64_t)i, (uint64_t)CfgA::N, re, im);
        hi[i] = cplx<T>{ (T)re, (T)im };
    }
    for (int i = 0; i < 256; i++) {
        long double re, im;
        unit_root((uint64_t)i, (uint64_t)p.n, re, im);
        lo[i] = cplx<T>{ (T)re, (T)im };
    }
    if (!rc)
        rc = upload_table<T>(&p.d_tw_hi, hi);
    if (!rc)
        rc = upload_table<T>(&p.d_tw_lo, lo);
    if (rc)
        return rc;
    // slab of frames that stays L2-resident between the two passes (about a quarter of the 126 MB L2)
    const size_t frame_bytes = (size_t)p.n * sizeof(cplx<T>);
    size_t slab = (32u << 20) / frame_bytes;
    if (slab < 1)
        slab = 1;
    p.scratch_frames = slab;
    if (cudaMalloc(&p.d_scratch, slab * frame_bytes) != cudaSuccess) {
        cudaGetLastError();
        return set_error(SDSP_B200_ERR_OOM, "fft: cannot allocate %zu bytes of scratch", slab * frame_bytes);
    }
    p.launch = &launch_large<CfgA, T>;
    p.tw_bytes = (tw.size() + hi.size() + lo.size()) * sizeof(cplx<T>);
    return SDSP_B200_OK;
}

// =================================================================================================
// 65536-point frames in ONE pass over HBM: a thread-block cluster of 8 CTAs holds the frame in its
// distributed shared memory (the same n = 256 a + b, k = k1 + 256 k2 split as above, but the [k1][b]
// intermediate never leaves the chip).
//   phase 1: CTA r owns columns b in [32r, 32r+32): 256-point transforms over a, read straight from HBM
//            (128-byte runs), times W_N^(b k1); result (k1, b) is written into the shared memory of CTA
//            k1 / 32 (st through the cluster's shared window), at row k1 % 32, position b;
//   phase 2: CTA r owns rows k1 in [32r, 32r+32): 256-point transforms over b out of its own shared
//            memory, stored to X[k1 + 256 k2] as 128-byte runs over k1.
// Each phase runs as two halves of 16 columns / rows (256 threads x 16 points), so a CTA needs one
// 16-transform exchange buffer plus the 32-row receive buffer (105 KB fp32: two clusters per SM overlap
// one another's load, exchange and store phases).  Two cluster barriers per frame, split into
// arrive/wait so that they cost nothing when CTAs are in step: "receive buffers are free" (arrive after
// the last read of phase 2, wait before the first remote write) and "all rows have arrived".
// Arithmetic is that of the two-kernel path above, operation for operation.
namespace cg = cooperative_groups;

__device__ __forceinline__ void cluster_arrive_release()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire()
{
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive without a fence: used where the only thing to order is this thread's own completed shared-memory READS (their
// values have been consumed by arithmetic before the call) against peers' later writes -- a release here would also wait
// for the global stores of the previous batch to drain (11 % of all stall samples in the first profile)
__device__ __forceinline__ void cluster_arrive_relaxed()
{
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t cluster_map_shared(uint32_t local_addr, uint32_t rank) // same offset in CTA `rank` of the cluster
{
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(local_addr), "r"(rank));
    return a;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, cplx<float> v)
{
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster(uint32_t addr, cplx<double> v)
{
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

// CL = CTAs per cluster: 8 (portable; two halves of 16 columns / rows per phase) or 16 (opt-in size; one
// transform batch per phase, 70 KB of shared memory in fp32, so three CTAs share an SM like the 4096-point kernel)
template <typename T, int CL, int MINB>
__global__ void __launch_bounds__(256, MINB)
    fft_cluster64k_kernel(cplx<T> *__restrict__ data, const T *__restrict__ real_in, const cplx<T> *__restrict__ tw,
                          const cplx<T> *__restrict__ tw_hi, const cplx<T> *__restrict__ tw_lo, size_t n_frames, int inverse, T scale)
{
    using Cfg = FftCfg<256, 16, 16, 16>;
    constexpr int PITCH = LargeStride<Cfg>::value;
    constexpr int N2 = 256;
    constexpr int OWN = N2 / CL;     // columns (phase 1) and rows (phase 2) this CTA owns
    constexpr int HALVES = OWN / 16; // batches of 16 transforms
    static_assert(CL == 8 || CL == 16, "cluster of 8 or 16 CTAs");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *xbuf = reinterpret_cast<cplx<T> *>(smem_raw); // exchange buffer of the 16 transforms in flight
    cplx<T> *rbuf = xbuf + 16 * PITCH;                     // OWN rows received from the whole cluster
    cplx<T> *s_hi = rbuf + OWN * PITCH, *s_lo = s_hi + 256; // the two 256-entry factors of W_N^x, read with lane-divergent indices
    s_hi[threadIdx.x] = tw_hi[threadIdx.x];
    s_lo[threadIdx.x] = tw_lo[threadIdx.x];
    __syncthreads();
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned r = cluster.block_rank();
    const size_t n_clusters = gridDim.x / CL, cid = blockIdx.x / CL;
    const int lo16 = threadIdx.x & 15, t = threadIdx.x >> 4; // column (phase 1) / row (phase 2) within the batch; thread in transform
    cplx<T> *fs = xbuf + (size_t)lo16 * PITCH;

    cluster_arrive_release(); // "my receive buffer is free" for the first frame
    for (size_t f = cid; f < n_frames; f += n_clusters) {
        cplx<T> *frame = data + f * ((size_t)N2 * N2);
        if (f + n_clusters < n_frames && reinterpret_cast<uintptr_t>(data) % 16 == 0 && reinterpret_cast<uintptr_t>(real_in) % 16 == 0) {
            // the columns this CTA reads in the next frame: one row piece per thread, into L2
            const size_t nf = f + n_clusters;
            if (real_in)
                prefetch_l2_bulk(real_in + nf * ((size_t)N2 * N2) + (size_t)threadIdx.x * N2 + OWN * r, OWN * (unsigned)sizeof(T));
            else
                prefetch_l2_bulk(data + nf * ((size_t)N2 * N2) + (size_t)threadIdx.x * N2 + OWN * r, OWN * (unsigned)sizeof(cplx<T>));
        }
        // ---- phase 1: columns
#pragma unroll 1
        for (int h = 0; h < HALVES; h++) {
            const unsigned b = (unsigned)OWN * r + 16u * (unsigned)h + (unsigned)lo16;
            const cplx<T> *gp = frame + b;
            cplx<T> v[Cfg::E];
            if (real_in) { // real samples, imaginary part zero; the spectrum goes to `data`
                const T *rp = real_in + f * ((size_t)N2 * N2) + b;
#pragma unroll
                for (int e = 0; e < Cfg::E; e++)
                    v[e] = cplx<T>{ rp[(size_t)(t + Cfg::S * e) * N2], (T)0 };
            } else {
#pragma unroll
                for (int e = 0; e < Cfg::E; e++)
                    v[e] = ld_stream(gp + (size_t)(t + Cfg::S * e) * N2);
            }
            if (inverse) {
#pragma unroll
                for (int e = 0; e < Cfg::E; e++)
                    v[e] = cplx<T>{ v[e].y, v[e].x };
            }
            __syncthreads(); // the previous transforms have read the exchange buffer
            fft_kernel_passes<Cfg, T, 256, MINB, 0>(v, fs, tw, t);
            if (h == 0)
                cluster_wait_acquire(); // every CTA has finished reading its receive buffer (previous frame)
            const uint32_t slot = (uint32_t)__cvta_generic_to_shared(rbuf) + (uint32_t)((t * PITCH + (int)(b + (b >> 4))) * (int)sizeof(cplx<T>));
            const TwiddleSeq<T> wseq(b * (unsigned)t, b * (unsigned)Cfg::S, s_hi, s_lo); // W_N^(b k1), k1 = t + 16 e
#pragma unroll
            for (int e = 0; e < Cfg::E; e++) { // k1 = t + 16 e lives in CTA k1 / OWN, row k1 % OWN
                constexpr int EPC = OWN / 16; // values of e per destination CTA
                const cplx<T> w = cmul(v[e], wseq.get(e));
                st_cluster(cluster_map_shared(slot + (uint32_t)(16 * (e % EPC) * PITCH * (int)sizeof(cplx<T>)), (unsigned)(e / EPC)), w);
            }
        }
        cluster_arrive_release(); // my share of every row has been written
        cluster_wait_acquire();   // all 256 columns of my rows are here
        // ---- phase 2: rows
#pragma unroll 1
        for (int g = 0; g < HALVES; g++) {
            const int row = 16 * g + lo16;
            const cplx<T> *rs = rbuf + (size_t)row * PITCH;
            cplx<T> v[Cfg::E];
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = rs[Cfg::pad(t + Cfg::S * e)];
            __syncthreads();
            if (g == HALVES - 1) {
                // nothing more to read from my receive buffer: free for the next frame.  The barrier above sits behind the
                // shared-memory loads of every thread of the CTA, so they have all returned.
                cluster_arrive_relaxed();
            }
            fft_kernel_passes<Cfg, T, 256, MINB, 0>(v, fs, tw, t);
            if (inverse) {
#pragma unroll
                for (int e = 0; e < Cfg::E; e++)
                    v[e] = cplx<T>{ v[e].y * scale, v[e].x * scale };
            }
            cplx<T> *op = frame + (unsigned)OWN * r + (unsigned)row;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                st_stream(op + (size_t)(t + Cfg::S * e) * N2, v[e]);
        }
    }
    cluster_wait_acquire(); // pairs with the last arrive; nobody leaves while a peer may still write to it
}

template <typename T, int CL, int MINB>
static int launch_cluster64k(const FftPlan &p, void *data, const void *real_in, size_t n_frames, cudaStream_t stream)
{
    if (n_frames == 0)
        return SDSP_B200_OK;
    const size_t clusters = (size_t)p.cluster_slots < n_frames ? (size_t)p.cluster_slots : n_frames;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CL), 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = CL;
    attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    SDSP_CUDA(cudaLaunchKernelEx(&cfg, fft_cluster64k_kernel<T, CL, MINB>, reinterpret_cast<cplx<T> *>(data), static_cast<const T *>(real_in),
                                 reinterpret_cast<const cplx<T> *>(p.d_tw_rows), reinterpret_cast<const cplx<T> *>(p.d_tw_hi),
                                 reinterpret_cast<const cplx<T> *>(p.d_tw_lo), n_frames, p.direction == SDSP_B200_REVERSE ? 1 : 0,
                                 (T)(1.0 / 65536.0)));
    return SDSP_B200_OK;
}

template <typename T, int CL, int MINB>
static int setup_cluster64k(FftPlan &p)
{
    using Cfg = FftCfg<256, 16, 16, 16>;
    p.smem_bytes = ((size_t)(16 + 256 / CL) * LargeStride<Cfg>::value + 512) * sizeof(cplx<T>);
    auto kern = fft_cluster64k_kernel<T, CL, MINB>;
    SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    if (CL > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return SDSP_B200_ERR_UNSUPPORTED;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.sm_count * MINB / CL * CL), 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = CL;
    attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int slots = 0;
    if (cudaOccupancyMaxActiveClusters(&slots, kern, &cfg) != cudaSuccess || slots < 1) {
        cudaGetLastError();
        return SDSP_B200_ERR_UNSUPPORTED;
    }
    p.cluster = true;
    p.cluster_ctas = CL;
    p.cluster_slots = slots;
    p.n1 = 256;
    p.npass = 4;
    p.e = 16;
    p.threads = 256;
    if (!p.d_tw_rows) {
        std::vector<cplx<T>> tw;
        int rb[4] = { 16, 16, 1, 1 };
        build_twiddles<T>(256, rb, 2, tw);
        int rc = upload_table<T>(&p.d_tw_rows, tw);
        std::vector<cplx<T>> hi(256), lo(256);
        for (int i = 0; i < 256; i++) {
            long double re, im;
            unit_root((uint64_t)i, (uint64_t)256, re, im);
            hi[i] = cplx<T>{ (T)re, (T)im };
            unit_root((uint64_t)i, (uint64_t)65536, re, im);
            lo[i] = cplx<T>{ (T)re, (T)im };
        }
        if (!rc)
            rc = upload_table<T>(&p.d_tw_hi, hi);
        if (!rc)
            rc = upload_table<T>(&p.d_tw_lo, lo);
        if (rc)
            return rc;
        p.tw_bytes = (tw.size() + hi.size() + lo.size()) * sizeof(cplx<T>);
    }
    p.launch = &launch_cluster64k<T, CL, MINB>;
    return SDSP_B200_OK;
}

// =================================================================================================
// 65536-point frames, third design: ONE persistent kernel runs the column tiles and the row tiles of the two-kernel
// path as a dependency-ordered work queue, so the [k1][b] intermediate lives in a small ring of scratch frames that
// stays in the 126 MB L2 and HBM sees each frame once in and once out -- with the occupancy of the 4096-point kernel
// (256 threads, 35 KB of shared memory, three CTAs per SM) instead of a cluster's.
//   work item q -> slot q / 16, tile q % 16.  Slots: the column tiles of frames 0 .. LAG-1, then alternately the
//   column tiles of frame f + LAG and the row tiles of frame f.  A row tile waits until the 16 column tiles of its
//   frame have been counted in col_done[f]; a column tile of frame f waits until the row tiles of frame f - RING have
//   released their scratch slot (row_done).  Items are handed out in order by an atomic ticket, and every wait points
//   at items with smaller tickets, which are held by CTAs that are already running: no deadlock, whatever the residency.
// geometry of the fused kernel for frames of N1 x 256 points: the column transforms have N1 points (N1 / 16 threads each, so a
// 256-thread CTA takes 256 / (N1/16) columns per tile), the row transforms 256; both phases have N1 / 16 tiles per frame
#ifndef SDSP_FUSED_LEAD_F32
#define SDSP_FUSED_LEAD_F32 512 // tiles of lead between a frame's column tiles and its row tiles (fp32)
#endif
#ifndef SDSP_FUSED_LEAD_F32_TMA
#define SDSP_FUSED_LEAD_F32_TMA 768 // (1024 with part of the ring pinned in L2 is 2 % faster but lets a quarter of the ring spill to HBM: profiles/r02_fft_lag_traffic.txt)
#endif
#ifndef SDSP_FUSED_POLL
#define SDSP_FUSED_POLL mbar_test // or mbar_try: one try_wait (suspends up to the hardware time limit) per round
#endif
#ifndef SDSP_FUSED_TMA_NST
#define SDSP_FUSED_TMA_NST 1
#endif
#ifndef SDSP_FUSED_TMA_DEFAULT
#define SDSP_FUSED_TMA_DEFAULT 1
#endif
#ifndef SDSP_FFT_L2_PERSIST_DEFAULT
#define SDSP_FFT_L2_PERSIST_DEFAULT 1
#endif
#ifndef SDSP_FFT_L2_PERSIST_MB_DEFAULT
#define SDSP_FFT_L2_PERSIST_MB_DEFAULT 16 // persisting carve-out of L2 in MB (sweep: profiles/r02_fft_l2_persist_sweep.txt)
#endif
#ifndef SDSP_FUSED_TMA_MINB
#define SDSP_FUSED_TMA_MINB 3
#endif
template <int N1>
struct FusedCols;
template <>
struct FusedCols<32> {
    using Cfg = FftCfg<32, 16, 16, 2>;
};
template <>
struct FusedCols<64> {
    using Cfg = FftCfg<64, 16, 16, 4>;
};
template <>
struct FusedCols<128> {
    using Cfg = FftCfg<128, 16, 16, 8>;
};
template <>
struct FusedCols<256> {
    using Cfg = FftCfg<256, 16, 16, 16>;
};
template <>
struct FusedCols<512> {
    using Cfg = FftCfg<512, 16, 16, 16, 2>;
};
template <>
struct FusedCols<1024> {
    using Cfg = FftCfg<1024, 16, 16, 16, 4>;
};
template <typename T, int N1>
struct FusedRing {
    static constexpr int TILES = N1 / 16;                     // tiles per frame, either phase
    static constexpr int COLS = 256 / (N1 / 16);              // columns per column tile
    // frames between a frame's column tiles and its row tiles: 512 (fp32) / 256 (fp64) tiles of lead, more than the CTAs in flight
    // (fp32 runs the data-mover kernel, which discards consumed ring lines: 768 tiles of lead, 48 MB with 16 MB of it pinned in L2, measured best up to
    // N1 = 512; 512 tiles for N1 = 1024)
    static constexpr int LAG = (sizeof(T) == 4 ? (N1 <= 512 ? SDSP_FUSED_LEAD_F32_TMA : SDSP_FUSED_LEAD_F32) : 256) / TILES;
    static constexpr int RING = 2 * LAG;                      // scratch frames: 32 MB (48 MB) whatever the frame size
};

__device__ __forceinline__ cplx<float> ld_l2(const cplx<float> *p)
{
    const float2 v = __ldcg(reinterpret_cast<const float2 *>(p));
    return { v.x, v.y };
}
__device__ __forceinline__ cplx<double> ld_l2(const cplx<double> *p)
{
    const double2 v = __ldcg(reinterpret_cast<const double2 *>(p));
    return { v.x, v.y };
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until(const unsigned *ctr, unsigned target, unsigned seen) // one thread
{
    while (seen < target) {
        __nanosleep(64);
        seen = ld_acquire_gpu(ctr);
    }
}

// decode work item q -> (column tile?, frame, tile); frame >= n_frames marks an empty slot
template <int LAG, int TILES>
__host__ __device__ __forceinline__ void fused_decode(size_t q, bool &cols, size_t &f, int &tile)
{
    const size_t slot = q / TILES;
    tile = (int)(q % TILES);
    if (slot < (size_t)LAG) {
        cols = true;
        f = slot;
    } else {
        const size_t u = slot - LAG;
        cols = (u & 1) == 0;
        f = cols ? LAG + u / 2 : u / 2;
    }
}

// Variants measured and dropped (profiles/r01_fft65536_variants.txt): L2 prefetch of the column tile one round ahead, cp.async
// staging of the next column tile, a separate transposing step in the row tiles, a single barrier per item.
template <typename T, int N1>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 3 : 1)
    fft_fused_kernel(cplx<T> *__restrict__ data, const T *__restrict__ real_in, cplx<T> *__restrict__ scratch,
                     const cplx<T> *__restrict__ tw_cols, const cplx<T> *__restrict__ tw, const cplx<T> *__restrict__ tw_hi,
                     const cplx<T> *__restrict__ tw_lo,
                        unsigned *__restrict__ ticket, unsigned *__restrict__ col_done, unsigned *__restrict__ row_done, size_t n_frames,
                        int inverse, T scale)
{
    using Cfg = FftCfg<256, 16, 16, 16>;          // rows
    using CCfg = typename FusedCols<N1>::Cfg;     // columns
    constexpr int PITCH = LargeStride<Cfg>::value, CPITCH = LargeStride<CCfg>::value;
    constexpr int N2 = 256, TILES = FusedRing<T, N1>::TILES, COLS = FusedRing<T, N1>::COLS;
    constexpr int LAG = FusedRing<T, N1>::LAG, RING = FusedRing<T, N1>::RING;
    constexpr int XBUF = 16 * PITCH > COLS * CPITCH ? 16 * PITCH : COLS * CPITCH;
    constexpr size_t FRAME = (size_t)N1 * N2;
    constexpr int MINB = sizeof(T) == 4 ? 3 : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T> *xbuf = reinterpret_cast<cplx<T> *>(smem_raw);
    cplx<T> *s_hi = xbuf + XBUF, *s_lo = s_hi + N1; // W_N1^i (N1 entries) and W_N^i (256 entries): W_N^x = hi[x >> 8] * lo[x & 255]
    __shared__ unsigned s_ticket, s_ready;
    for (int i = threadIdx.x; i < N1; i += 256)
        s_hi[i] = tw_hi[i];
    s_lo[threadIdx.x] = tw_lo[threadIdx.x];
    const int lo16 = threadIdx.x & 15, hi16 = threadIdx.x >> 4;
    const int ccol = threadIdx.x % COLS, ct = threadIdx.x / COLS; // column tiles: column within the tile, thread within the transform
    const size_t total = ((size_t)LAG + 2 * n_frames) * TILES;

    // Latency of the queue itself is kept off the critical path: thread 0 draws tickets ahead of their use (the atomic's
    // round trip overlaps an item), looks at the counter the coming item depends on as soon as its ticket is back, and
    // publishes both with the barrier that ends the item.
    size_t q = 0; // item in hand
    if (threadIdx.x == 0) {
        const unsigned q0 = atomicAdd(ticket, 1u);
        s_ticket = q0;
        bool c0;
        size_t f0;
        int t0;
        fused_decode<LAG, TILES>(q0, c0, f0, t0);
        if (q0 < total && f0 < n_frames && !c0)
            spin_until(col_done + f0, TILES, 0);
    }
    __syncthreads();
    q = s_ticket;
    // this thread's share of the inter-transform factors (see fft_fused_tma_kernel: same arithmetic, same bits)
    const cplx<T> w_a = s_lo[(unsigned)ccol * (unsigned)ct], w_b = s_lo[(unsigned)ccol * (unsigned)CCfg::S];
    for (;;) {
        if (q >= total)
            break;
        unsigned drawn = 0;
        if (threadIdx.x == 0)
            drawn = atomicAdd(ticket, 1u);
        bool cols;
        size_t f;
        int tile;
        fused_decode<LAG, TILES>(q, cols, f, tile);
        cplx<T> *sc = scratch + (f % RING) * (FRAME);
        cplx<T> v[Cfg::E];
        unsigned *done = nullptr;
        if (f >= n_frames) {
            __syncthreads(); // empty slot (column tiles past the last frame): everyone has read the mailbox before it is rewritten
        } else if (cols) {
            // ---- COLS columns b = COLS tile + ccol: N1-point transforms over a, times W_N^(b k1), to scratch [k1][b]
            const int t = ct;
            const unsigned b = (unsigned)COLS * (unsigned)tile + (unsigned)ccol;
            unsigned seen = TILES;
            if (threadIdx.x == 0 && f >= (size_t)RING)
                seen = ld_acquire_gpu(row_done + (f - RING)); // has the scratch slot's previous tenant been read out?
            if (real_in) {
                const T *rp = real_in + f * FRAME + b;
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = cplx<T>{ __ldcs(rp + (size_t)(t + CCfg::S * e) * N2), (T)0 };
            } else {
                const cplx<T> *gp = data + f * FRAME + b;
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = ld_stream(gp + (size_t)(t + CCfg::S * e) * N2);
            }
            if (inverse) {
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = cplx<T>{ v[e].y, v[e].x };
            }
            if (threadIdx.x == 0 && f >= (size_t)RING)
                spin_until(row_done + (f - RING), TILES, seen);
            // (the barriers inside the passes sit between thread 0's check above and every thread's stores below)
            fft_kernel_passes<CCfg, T, 256, MINB, 0>(v, xbuf + (size_t)ccol * CPITCH, tw_cols, t);
            const unsigned xt = (unsigned)COLS * (unsigned)tile * (unsigned)t;
            const TwiddleGeo<T> wseq(cmul(w_a, cmul(s_hi[xt >> 8], s_lo[xt & 255u])), cmul(w_b, s_hi[tile]));
            cplx<T> *op = sc + b;
#pragma unroll
            for (int e = 0; e < CCfg::E; e++)
                op[(size_t)(t + CCfg::S * e) * N2] = cmul(v[e], wseq.get(e));
            done = col_done + f;
        } else {
            // ---- 16 rows k1 = 16 tile + r: 256-point transforms over b out of scratch, stored to X[k1 + N1 k2]
            // (col_done[f] was checked by thread 0 before the barrier that published this ticket)
            // The first pass runs with lanes along b (thread = (t, row): coalesced 128-byte reads of the ring); the exchange
            // between the passes also re-maps the threads, so the second pass runs with lanes along the rows
            // (thread = (row, t)) and its natural-order outputs X[k1 + N1 (t + 16 e)] leave as 128-byte runs over k1 --
            // no separate transposing step.
            const int t = lo16, row = hi16;
            const cplx<T> *gp = sc + (size_t)(16 * tile + row) * N2 + t;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = ld_l2(gp + Cfg::S * e); // written by other SMs: read at L2, never from this SM's L1
            fft_pass<Cfg, 0, T>(v, t, tw);
            cplx<T> *fs = xbuf + (size_t)row * PITCH;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                fs[fft_out_phys<Cfg, 0>(t, e)] = v[e];
            __syncthreads();
            const int row2 = lo16, t2 = hi16;
            const cplx<T> *rs = xbuf + (size_t)row2 * PITCH;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = rs[fft_read_phys<Cfg>(t2, e)];
            fft_pass<Cfg, 1, T>(v, t2, tw);
            if (inverse) {
#pragma unroll
                for (int e = 0; e < Cfg::E; e++)
                    v[e] = cplx<T>{ v[e].y * scale, v[e].x * scale };
            }
            cplx<T> *op = data + f * (FRAME) + 16 * tile + row2;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                st_stream(op + (size_t)(t2 + Cfg::S * e) * N1, v[e]); // natural order: slot e of thread t2 is k2 = t2 + 16 e, X[k1 + N1 k2]
            done = row_done + f;
        }
        // ---- end of item: publish the ticket drawn during it, and whether the item that comes next has its dependency met,
        // with the barrier that also orders this item's stores before the completion count
        const size_t coming = drawn; // (thread 0's view; the other threads learn it from the mailbox)
        const unsigned *dep = nullptr;
        if (threadIdx.x == 0) {
            s_ticket = drawn;
            bool nc;
            size_t nf;
            int nt;
            fused_decode<LAG, TILES>(coming, nc, nf, nt);
            if (coming < total && nf < n_frames && !nc)
                dep = col_done + nf;
            s_ready = dep == nullptr || ld_acquire_gpu(dep) >= TILES; // (never spin here: this item is not counted yet)
        }
        __syncthreads();
        if (threadIdx.x == 0 && done) {
            __threadfence();
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(done) : "memory");
        }
        if (!s_ready) { // rare: the coming item's column tiles are still running somewhere
            if (threadIdx.x == 0)
                spin_until(dep, TILES, 0);
            __syncthreads();
        }
        q = s_ticket;
    }
}

template <typename T, int N1>
static int launch_fused(const FftPlan &p, void *data, const void *real_in, size_t n_frames, cudaStream_t stream)
{
    if (n_frames == 0)
        return SDSP_B200_OK;
    // counters: [0] ticket, then col_done[n], row_done[n]; zeroed for every launch
    const size_t need = (1 + 2 * n_frames) * sizeof(unsigned);
    FftPlan &mp = const_cast<FftPlan &>(p);
    if (mp.fused_counter_bytes < need) {
        if (mp.d_fused_counters)
            cudaFree(mp.d_fused_counters);
        mp.d_fused_counters = nullptr;
        mp.fused_counter_bytes = 0;
        if (cudaMalloc(&mp.d_fused_counters, need) != cudaSuccess) {
            cudaGetLastError();
            return set_error(SDSP_B200_ERR_OOM, "fft: cannot allocate %zu bytes of work-queue counters", need);
        }
        mp.fused_counter_bytes = need;
    }
    SDSP_CUDA(cudaMemsetAsync(mp.d_fused_counters, 0, need, stream));
    unsigned *ctr = static_cast<unsigned *>(mp.d_fused_counters);
    const size_t items = ((size_t)FusedRing<T, N1>::LAG + 2 * n_frames) * FusedRing<T, N1>::TILES;
    size_t grid = (size_t)p.sm_count * (size_t)p.ctas_per_sm;
    if (grid > items)
        grid = items;
    fft_fused_kernel<T, N1><<<(unsigned)grid, 256, p.smem_bytes, stream>>>(
        reinterpret_cast<cplx<T> *>(data), static_cast<const T *>(real_in), reinterpret_cast<cplx<T> *>(p.d_scratch),
        reinterpret_cast<const cplx<T> *>(p.d_tw_cols), reinterpret_cast<const cplx<T> *>(p.d_tw_rows),
        reinterpret_cast<const cplx<T> *>(p.d_tw_hi), reinterpret_cast<const cplx<T> *>(p.d_tw_lo), ctr, ctr + 1, ctr + 1 + n_frames, n_frames,
        p.direction == SDSP_B200_REVERSE ? 1 : 0, (T)(1.0 / ((double)N1 * 256.0)));
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

// -------------------------------------------------------------------------------------------------
// The fused queue again, fed by a data-mover warp (fp32: the default there; SDSP_B200_FFT_FUSED_TMA=0 disables).  Warp 8 draws the tickets, waits
// for each item's dependency and brings its 32 KB tile into a shared-memory stage -- a 3-D TMA box for a column tile (COLS columns
// x N1 rows out of the frame in HBM), one bulk copy for a row tile (16 contiguous rows of the L2-resident ring) -- while the 256
// compute threads are still busy with the item before: their items start with shared-memory reads, the queue's latency (ticket,
// dependency look-up) leaves the compute warps' path altogether, and the only CTA-wide waits left are the exchange barriers.
//   full[s]  : the producer's arrive (+ the tile's bytes)  -> the compute threads may read stage s and mailbox s_item[s]
//   empty[s] : 256 arrivals, each thread after it has copied its 16 points into registers -> the producer may refill stage s
// Ordering of the scratch ring across CTAs: column tiles store with ordinary st.global, then (barrier) thread 0 counts the tile
// with red.release.gpu; the producer that wants the frame's rows reads the counter with ld.acquire.gpu and issues
// fence.proxy.async before its bulk copy (generic-proxy writes -> async-proxy read).
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void wait_counter(const unsigned *ctr, unsigned target) // one thread
{
    while (ld_acquire_gpu(ctr) < target)
        __nanosleep(64);
}

template <typename T, int N1, int NST, int MINB>
__global__ void __launch_bounds__(288, MINB)
    fft_fused_tma_kernel(const __grid_constant__ CUtensorMap in_map, cplx<T> *__restrict__ data, int real_in, cplx<T> *__restrict__ scratch,
                         const cplx<T> *__restrict__ tw_cols, const cplx<T> *__restrict__ tw, const cplx<T> *__restrict__ tw_hi,
                         const cplx<T> *__restrict__ tw_lo, unsigned *__restrict__ ticket, unsigned *__restrict__ col_done,
                         unsigned *__restrict__ row_done, size_t n_frames, int inverse, T scale)
{
    using Cfg = FftCfg<256, 16, 16, 16>;          // rows
    using CCfg = typename FusedCols<N1>::Cfg;     // columns
    constexpr int BOXES = N1 <= 256 ? 1 : N1 / 256; // a TMA box holds at most 256 rows
    constexpr int PITCH = LargeStride<Cfg>::value, CPITCH = LargeStride<CCfg>::value;
    constexpr int N2 = 256, TILES = FusedRing<T, N1>::TILES, COLS = FusedRing<T, N1>::COLS;
    constexpr int LAG = FusedRing<T, N1>::LAG, RING = FusedRing<T, N1>::RING;
    constexpr int XBUF = 16 * PITCH > COLS * CPITCH ? 16 * PITCH : COLS * CPITCH;
    constexpr size_t FRAME = (size_t)N1 * N2;
    constexpr uint32_t TILE_BYTES = 4096 * sizeof(cplx<T>);
    constexpr int ND = NST + 1; // completion barriers in rotation
    extern __shared__ __align__(128) unsigned char smem_raw128[];
    cplx<T> *stage0 = reinterpret_cast<cplx<T> *>(smem_raw128);
    cplx<T> *xbuf = stage0 + (size_t)NST * 4096;
    cplx<T> *s_hi = xbuf + XBUF, *s_lo = s_hi + N1;
    // (no static shared memory in this kernel: the dynamic area then starts at the window's aligned base, as the TMA boxes need)
    uint64_t *full = reinterpret_cast<uint64_t *>(s_lo + 256), *empty = full + NST, *done_bar = empty + NST, *xfree = done_bar + ND;
    unsigned *s_item = reinterpret_cast<unsigned *>(xfree + 1);
    for (int i = threadIdx.x; i < N1; i += 288)
        s_hi[i] = tw_hi[i];
    if (threadIdx.x < 256)
        s_lo[threadIdx.x] = tw_lo[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 256);
        }
        for (int s = 0; s < ND; s++)
            mbar_init(&done_bar[s], 256);
        mbar_init(xfree, 256);
        fence_mbar_init();
    }
    __syncthreads();
    const size_t total = ((size_t)LAG + 2 * n_frames) * TILES;

    if (threadIdx.x >= 256) { // ---- the data mover; it also counts finished tiles, so no compute warp ever waits on a fence
        if (threadIdx.x != 256)
            return;
        // Finished tiles are counted as soon as their barrier completes -- in particular from inside every wait below: a tile that is
        // done but not yet counted may be exactly what another CTA's (or this CTA's next) item is waiting for.
        unsigned *pend[ND];
        for (int i = 0; i < ND; i++)
            pend[i] = nullptr;
        unsigned it = 0, next_pub = 0; // items issued so far = it; items counted so far = next_pub
        auto count_tile = [&](unsigned j) {
            unsigned *d = pend[j % ND];
            if (d)
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(d) : "memory");
        };
        auto try_publish = [&]() {
            while (next_pub < it && mbar_test(&done_bar[next_pub % ND], (next_pub / ND) & 1)) {
                count_tile(next_pub);
                next_pub++;
            }
        };
        auto wait_dep = [&](const unsigned *ctr) {
            while (ld_acquire_gpu(ctr) < (unsigned)TILES) {
                try_publish();
                __nanosleep(32);
            }
        };
        // When to draw the next ticket: as late as possible -- once the stage is free.  A ticket held without being worked on delays
        // every item that depends on it: drawing a whole item early, or when the compute threads pass the exchange of the item
        // before, were both measured slower although they take the ticket's round trip off the path (profiles/r01_fft65536_variants.txt).
        for (;; it++) {
            const int s = it % NST;
            if (it >= (unsigned)NST) {
                // (counting here, promptly, matters: leaving it until the next copy is out was measured 3 % slower -- other CTAs'
                // row tiles follow their frame's column tiles by barely more than an item's latency)
                while (!SDSP_FUSED_POLL(&empty[s], ((it / NST) - 1) & 1)) // until item it - NST has been taken into registers
                    try_publish();
                while (next_pub + ND <= it) { // this item's completion barrier is free again once item it - ND is counted
                    mbar_wait(&done_bar[next_pub % ND], (next_pub / ND) & 1);
                    count_tile(next_pub);
                    next_pub++;
                }
            }
            const size_t q = atomicAdd(ticket, 1u);
            bool cols = false;
            size_t f = n_frames;
            int tile = 0;
            if (q < total)
                fused_decode<LAG, TILES>(q, cols, f, tile);
            const bool real = q < total && f < n_frames;
            if (real && !cols)
                wait_dep(col_done + f); // the frame's column tiles are all in the ring
            s_item[s] = q < total ? (unsigned)q : 0xffffffffu;
            if (q >= total) {
                mbar_arrive(&full[s]);
                break;
            }
            cplx<T> *dst = stage0 + (size_t)s * 4096;
            unsigned *d = nullptr;
            if (!real) {
                mbar_arrive(&full[s]); // empty slot
            } else if (cols) {
                // the dependency of a column tile guards the compute threads' stores into the ring, not this copy: start the copy,
                // then look at the counter, and only then hand the stage over
                mbar_expect_tx_only(&full[s], real_in ? TILE_BYTES / 2 : TILE_BYTES);
#pragma unroll
                for (int bx = 0; bx < BOXES; bx++) {
                    void *bd = real_in ? static_cast<void *>(reinterpret_cast<T *>(dst) + (size_t)bx * 256 * COLS)
                                       : static_cast<void *>(dst + (size_t)bx * 256 * COLS);
                    tma_load_3d(bd, &in_map, real_in ? COLS * tile : 2 * COLS * tile, 256 * bx, (int)f, &full[s]);
                }
                if (f >= (size_t)RING)
                    wait_dep(row_done + (f - RING)); // the ring slot's previous tenant has been read out
                mbar_arrive(&full[s]);
                d = col_done + f;
            } else {
                asm volatile("fence.proxy.async.global;" ::: "memory");
                mbar_expect_tx(&full[s], TILE_BYTES);
                bulk_load_1d(dst, scratch + (f % RING) * FRAME + (size_t)(16 * tile) * N2, TILE_BYTES, &full[s]);
                d = row_done + f;
            }
            pend[it % ND] = d;
        }
        for (; next_pub < it; next_pub++) { // the tail: the last items are still to be counted
            mbar_wait(&done_bar[next_pub % ND], (next_pub / ND) & 1);
            count_tile(next_pub);
        }
        return;
    }

    // ---- the 256 compute threads: one CTA-wide barrier per item (the exchange); reuse of the exchange buffer by the next item is
    // ordered by a split barrier (xfree: arrive after the exchange reads, wait before the next item's exchange writes)
    const int lo16 = threadIdx.x & 15, hi16 = threadIdx.x >> 4;
    const int ccol = threadIdx.x % COLS, ct = threadIdx.x / COLS;
    // inter-transform factor of a column tile: W_N^(b (ct + S e)), b = COLS tile + ccol, = w0 r^e with
    //   w0 = W^(ccol ct) W^(COLS tile ct)   and   r = W^(S ccol) W^(S COLS tile) = W^(S ccol) W_N1^tile     (S COLS = 256)
    // the first factor of each is this thread's for the whole launch (indices below COLS S = 256: the W_N^i table)
    const cplx<T> w_a = s_lo[(unsigned)ccol * (unsigned)ct], w_b = s_lo[(unsigned)ccol * (unsigned)CCfg::S];
    unsigned n_real = 0;
    for (unsigned it = 0;; it++) {
        const int s = it % NST;
        mbar_wait(&full[s], (it / NST) & 1);
        const unsigned q = s_item[s];
        if (q == 0xffffffffu)
            break;
        bool cols;
        size_t f;
        int tile;
        fused_decode<LAG, TILES>(q, cols, f, tile);
        const cplx<T> *st = stage0 + (size_t)s * 4096;
        cplx<T> *sc = scratch + (f % RING) * (FRAME);
        cplx<T> v[Cfg::E];
        if (f >= n_frames) {
            mbar_arrive(&empty[s]);
            mbar_arrive(&done_bar[it % ND]);
            continue;
        }
        if (cols) {
            const int t = ct;
            const unsigned b = (unsigned)COLS * (unsigned)tile + (unsigned)ccol;
            if (real_in) {
                const T *rs = reinterpret_cast<const T *>(st);
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = cplx<T>{ rs[(t + CCfg::S * e) * COLS + ccol], (T)0 };
            } else {
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = st[(t + CCfg::S * e) * COLS + ccol];
            }
            mbar_arrive(&empty[s]);
            if (inverse) {
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = cplx<T>{ v[e].y, v[e].x };
            }
            cplx<T> *fs = xbuf + (size_t)ccol * CPITCH;
            if constexpr (CCfg::NPASS == 2) {
                fft_pass<CCfg, 0, T>(v, t, tw_cols);
                if (n_real > 0)
                    mbar_wait(xfree, (n_real - 1) & 1);
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    fs[fft_out_phys<CCfg, 0>(t, e)] = v[e];
                cta_sync<1, 256>();
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = fs[fft_read_phys<CCfg>(t, e)];
                mbar_arrive(xfree);
                fft_pass<CCfg, 1, T>(v, t, tw_cols);
            } else { // three passes: two exchanges with the usual barriers between them, inside the split barrier
                if (n_real > 0)
                    mbar_wait(xfree, (n_real - 1) & 1);
                fft_kernel_passes<CCfg, T, 256, MINB, 0, false, 1>(v, fs, tw_cols, t);
                mbar_arrive(xfree);
            }
            // (the per-tile factors are looked up with at most two distinct addresses per warp: broadcasts, no bank conflicts)
            const unsigned xt = (unsigned)COLS * (unsigned)tile * (unsigned)t;
            const TwiddleGeo<T> wseq(cmul(w_a, cmul(s_hi[xt >> 8], s_lo[xt & 255u])), cmul(w_b, s_hi[tile]));
            cplx<T> *op = sc + b;
#pragma unroll
            for (int e = 0; e < CCfg::E; e++)
                op[(size_t)(t + CCfg::S * e) * N2] = cmul(v[e], wseq.get(e));
        } else {
            const int t = lo16, row = hi16;
            const cplx<T> *gp = st + row * N2 + t;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = gp[Cfg::S * e];
            mbar_arrive(&empty[s]);
            // the tile's 32 KB of the ring are dead now (256 lines of 128 bytes, one per thread): drop them from L2 instead of
            // letting them be written back to HBM when they are evicted -- that is what makes a 48 MB ring affordable
            // (profiles/r01_fft65536_variants.txt)
            if constexpr (sizeof(cplx<T>) == 8)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(sc + (size_t)(16 * tile) * N2 + (size_t)threadIdx.x * 16) : "memory");
            fft_pass<Cfg, 0, T>(v, t, tw);
            cplx<T> *fs = xbuf + (size_t)row * PITCH;
            if (n_real > 0)
                mbar_wait(xfree, (n_real - 1) & 1);
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                fs[fft_out_phys<Cfg, 0>(t, e)] = v[e];
            cta_sync<1, 256>();
            const int row2 = lo16, t2 = hi16;
            const cplx<T> *rs = xbuf + (size_t)row2 * PITCH;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = rs[fft_read_phys<Cfg>(t2, e)];
            mbar_arrive(xfree);
            fft_pass<Cfg, 1, T>(v, t2, tw);
            if (inverse) {
#pragma unroll
                for (int e = 0; e < Cfg::E; e++)
                    v[e] = cplx<T>{ v[e].y * scale, v[e].x * scale };
            }
            cplx<T> *op = data + f * (FRAME) + 16 * tile + row2;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                st_stream(op + (size_t)(t2 + Cfg::S * e) * N1, v[e]);
        }
        mbar_arrive(&done_bar[it % ND]); // (release: this thread's stores are ordered before the data mover's count)
        n_real++;
    }
}

// -------------------------------------------------------------------------------------------------
// The data-mover kernel with TWO tile slots and no separate exchange buffer: the compute threads exchange IN the slot whose tile
// they have just taken into registers (each slot is as large as the padded exchange layout), so a second slot costs 3 KB instead
// of 32 and the kernel keeps three CTAs per SM.  The data mover is then a whole item ahead: ticket, dependency look-up and copy of
// item i + 1 run while item i is transformed, and the compute threads' wait for their next tile (17 % of the stall samples of the
// one-slot kernel, profiles/r01_ncu_fft65536_f32_fused_tma_v2.txt) disappears.
//   full[s]  : the producer's arrive (+ the tile's bytes)            -> the compute threads may read slot s and mailbox s_item[s]
//   empty[s] : 256 arrivals, each thread after its last exchange read -> the producer may overwrite slot s (generic-proxy accesses
//              are ordered before the TMA's async-proxy write by fence.proxy.async in every arriving thread)
// Between a thread's register loads from the slot and the first exchange write into it sits one CTA barrier (all 256 have loaded).
template <typename T, int N1, int MINB>
__global__ void __launch_bounds__(288, MINB)
    fft_fused_tma2_kernel(const __grid_constant__ CUtensorMap in_map, cplx<T> *__restrict__ data, int real_in, cplx<T> *__restrict__ scratch,
                         const cplx<T> *__restrict__ tw_cols, const cplx<T> *__restrict__ tw, const cplx<T> *__restrict__ tw_hi,
                         const cplx<T> *__restrict__ tw_lo, unsigned *__restrict__ ticket, unsigned *__restrict__ col_done,
                         unsigned *__restrict__ row_done, size_t n_frames, int inverse, T scale)
{
    using Cfg = FftCfg<256, 16, 16, 16>;          // rows
    using CCfg = typename FusedCols<N1>::Cfg;     // columns
    constexpr int BOXES = N1 <= 256 ? 1 : N1 / 256; // a TMA box holds at most 256 rows
    constexpr int PITCH = LargeStride<Cfg>::value, CPITCH = LargeStride<CCfg>::value;
    constexpr int N2 = 256, TILES = FusedRing<T, N1>::TILES, COLS = FusedRing<T, N1>::COLS;
    constexpr int LAG = FusedRing<T, N1>::LAG, RING = FusedRing<T, N1>::RING;
    constexpr int XBUF = 16 * PITCH > COLS * CPITCH ? 16 * PITCH : COLS * CPITCH;
    constexpr size_t FRAME = (size_t)N1 * N2;
    constexpr uint32_t TILE_BYTES = 4096 * sizeof(cplx<T>);
    constexpr int NST = 2, ND = NST + 1; // completion barriers in rotation
    constexpr int SLOT = (XBUF + 15) / 16 * 16; // complex elements per slot (128-byte multiple: the second slot stays TMA-aligned)
    extern __shared__ __align__(128) unsigned char smem_raw128[];
    cplx<T> *stage0 = reinterpret_cast<cplx<T> *>(smem_raw128);
    cplx<T> *s_hi = stage0 + (size_t)NST * SLOT, *s_lo = s_hi + N1;
    // (no static shared memory in this kernel: the dynamic area then starts at the window's aligned base, as the TMA boxes need)
    uint64_t *full = reinterpret_cast<uint64_t *>(s_lo + 256), *empty = full + NST, *done_bar = empty + NST;
    unsigned *s_item = reinterpret_cast<unsigned *>(done_bar + ND);
    for (int i = threadIdx.x; i < N1; i += 288)
        s_hi[i] = tw_hi[i];
    if (threadIdx.x < 256)
        s_lo[threadIdx.x] = tw_lo[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 256);
        }
        for (int s = 0; s < ND; s++)
            mbar_init(&done_bar[s], 256);
        fence_mbar_init();
    }
    __syncthreads();
    const size_t total = ((size_t)LAG + 2 * n_frames) * TILES;

    if (threadIdx.x >= 256) { // ---- the data mover; it also counts finished tiles, so no compute warp ever waits on a fence
        if (threadIdx.x != 256)
            return;
        // Finished tiles are counted as soon as their barrier completes -- in particular from inside every wait below: a tile that is
        // done but not yet counted may be exactly what another CTA's (or this CTA's next) item is waiting for.
        unsigned *pend[ND];
        for (int i = 0; i < ND; i++)
            pend[i] = nullptr;
        unsigned it = 0, next_pub = 0; // items issued so far = it; items counted so far = next_pub
        auto count_tile = [&](unsigned j) {
            unsigned *d = pend[j % ND];
            if (d)
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(d) : "memory");
        };
        auto try_publish = [&]() {
            while (next_pub < it && mbar_test(&done_bar[next_pub % ND], (next_pub / ND) & 1)) {
                count_tile(next_pub);
                next_pub++;
            }
        };
        auto wait_dep = [&](const unsigned *ctr) {
            while (ld_acquire_gpu(ctr) < (unsigned)TILES) {
                try_publish();
                __nanosleep(32);
            }
        };
        // When to draw the next ticket: as late as possible -- once the stage is free.  A ticket held without being worked on delays
        // every item that depends on it: drawing a whole item early, or when the compute threads pass the exchange of the item
        // before, were both measured slower although they take the ticket's round trip off the path (profiles/r01_fft65536_variants.txt).
        for (;; it++) {
            const int s = it % NST;
            if (it >= (unsigned)NST) {
                // (counting here, promptly, matters: leaving it until the next copy is out was measured 3 % slower -- other CTAs'
                // row tiles follow their frame's column tiles by barely more than an item's latency)
                while (!SDSP_FUSED_POLL(&empty[s], ((it / NST) - 1) & 1)) // until item it - NST has finished exchanging in the slot
                    try_publish();
                while (next_pub + ND <= it) { // this item's completion barrier is free again once item it - ND is counted
                    mbar_wait(&done_bar[next_pub % ND], (next_pub / ND) & 1);
                    count_tile(next_pub);
                    next_pub++;
                }
            }
            const size_t q = atomicAdd(ticket, 1u);
            bool cols = false;
            size_t f = n_frames;
            int tile = 0;
            if (q < total)
                fused_decode<LAG, TILES>(q, cols, f, tile);
            const bool real = q < total && f < n_frames;
            if (real && !cols)
                wait_dep(col_done + f); // the frame's column tiles are all in the ring
            s_item[s] = q < total ? (unsigned)q : 0xffffffffu;
            if (q >= total) {
                mbar_arrive(&full[s]);
                break;
            }
            cplx<T> *dst = stage0 + (size_t)s * SLOT;
            unsigned *d = nullptr;
            if (!real) {
                mbar_arrive(&full[s]); // empty slot
            } else if (cols) {
                // the dependency of a column tile guards the compute threads' stores into the ring, not this copy: start the copy,
                // then look at the counter, and only then hand the stage over
                mbar_expect_tx_only(&full[s], real_in ? TILE_BYTES / 2 : TILE_BYTES);
#pragma unroll
                for (int bx = 0; bx < BOXES; bx++) {
                    void *bd = real_in ? static_cast<void *>(reinterpret_cast<T *>(dst) + (size_t)bx * 256 * COLS)
                                       : static_cast<void *>(dst + (size_t)bx * 256 * COLS);
                    tma_load_3d(bd, &in_map, real_in ? COLS * tile : 2 * COLS * tile, 256 * bx, (int)f, &full[s]);
                }
                if (f >= (size_t)RING)
                    wait_dep(row_done + (f - RING)); // the ring slot's previous tenant has been read out
                mbar_arrive(&full[s]);
                d = col_done + f;
            } else {
                asm volatile("fence.proxy.async.global;" ::: "memory");
                mbar_expect_tx(&full[s], TILE_BYTES);
                bulk_load_1d(dst, scratch + (f % RING) * FRAME + (size_t)(16 * tile) * N2, TILE_BYTES, &full[s]);
                d = row_done + f;
            }
            pend[it % ND] = d;
        }
        for (; next_pub < it; next_pub++) { // the tail: the last items are still to be counted
            mbar_wait(&done_bar[next_pub % ND], (next_pub / ND) & 1);
            count_tile(next_pub);
        }
        return;
    }

    // ---- the 256 compute threads: two CTA-wide barriers per item (slot taken into registers; the exchange)
    const int lo16 = threadIdx.x & 15, hi16 = threadIdx.x >> 4;
    const int ccol = threadIdx.x % COLS, ct = threadIdx.x / COLS;
    // inter-transform factor of a column tile: W_N^(b (ct + S e)), b = COLS tile + ccol, = w0 r^e with
    //   w0 = W^(ccol ct) W^(COLS tile ct)   and   r = W^(S ccol) W^(S COLS tile) = W^(S ccol) W_N1^tile     (S COLS = 256)
    // the first factor of each is this thread's for the whole launch (indices below COLS S = 256: the W_N^i table)
    const cplx<T> w_a = s_lo[(unsigned)ccol * (unsigned)ct], w_b = s_lo[(unsigned)ccol * (unsigned)CCfg::S];
    for (unsigned it = 0;; it++) {
        const int s = it % NST;
        mbar_wait(&full[s], (it / NST) & 1);
        const unsigned q = s_item[s];
        if (q == 0xffffffffu)
            break;
        bool cols;
        size_t f;
        int tile;
        fused_decode<LAG, TILES>(q, cols, f, tile);
        cplx<T> *st = stage0 + (size_t)s * SLOT, *xbuf = st;
        cplx<T> *sc = scratch + (f % RING) * (FRAME);
        cplx<T> v[Cfg::E];
        if (f >= n_frames) {
            mbar_arrive(&empty[s]);
            mbar_arrive(&done_bar[it % ND]);
            continue;
        }
        if (cols) {
            const int t = ct;
            const unsigned b = (unsigned)COLS * (unsigned)tile + (unsigned)ccol;
            if (real_in) {
                const T *rs = reinterpret_cast<const T *>(st);
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = cplx<T>{ rs[(t + CCfg::S * e) * COLS + ccol], (T)0 };
            } else {
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = st[(t + CCfg::S * e) * COLS + ccol];
            }
            cta_sync<1, 256>(); // every thread has its points: the slot becomes the exchange buffer
            if (inverse) {
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = cplx<T>{ v[e].y, v[e].x };
            }
            cplx<T> *fs = xbuf + (size_t)ccol * CPITCH;
            if constexpr (CCfg::NPASS == 2) {
                fft_pass<CCfg, 0, T>(v, t, tw_cols);
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    fs[fft_out_phys<CCfg, 0>(t, e)] = v[e];
                cta_sync<1, 256>();
#pragma unroll
                for (int e = 0; e < CCfg::E; e++)
                    v[e] = fs[fft_read_phys<CCfg>(t, e)];
                fence_proxy_async();
                mbar_arrive(&empty[s]);
                fft_pass<CCfg, 1, T>(v, t, tw_cols);
            } else { // three passes: two exchanges with the usual barriers between them
                fft_kernel_passes<CCfg, T, 256, MINB, 0, false, 1>(v, fs, tw_cols, t);
                fence_proxy_async();
                mbar_arrive(&empty[s]);
            }
            // (the per-tile factors are looked up with at most two distinct addresses per warp: broadcasts, no bank conflicts)
            const unsigned xt = (unsigned)COLS * (unsigned)tile * (unsigned)t;
            const TwiddleGeo<T> wseq(cmul(w_a, cmul(s_hi[xt >> 8], s_lo[xt & 255u])), cmul(w_b, s_hi[tile]));
            cplx<T> *op = sc + b;
#pragma unroll
            for (int e = 0; e < CCfg::E; e++)
                op[(size_t)(t + CCfg::S * e) * N2] = cmul(v[e], wseq.get(e));
        } else {
            const int t = lo16, row = hi16;
            const cplx<T> *gp = st + row * N2 + t;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = gp[Cfg::S * e];
            cta_sync<1, 256>(); // every thread has its points: the slot becomes the exchange buffer
            // the tile's 32 KB of the ring are dead now (256 lines of 128 bytes, one per thread): drop them from L2 instead of
            // letting them be written back to HBM when they are evicted -- that is what makes a 48 MB ring affordable
            // (profiles/r01_fft65536_variants.txt)
            if constexpr (sizeof(cplx<T>) == 8)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(sc + (size_t)(16 * tile) * N2 + (size_t)threadIdx.x * 16) : "memory");
            fft_pass<Cfg, 0, T>(v, t, tw);
            cplx<T> *fs = xbuf + (size_t)row * PITCH;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                fs[fft_out_phys<Cfg, 0>(t, e)] = v[e];
            cta_sync<1, 256>();
            const int row2 = lo16, t2 = hi16;
            const cplx<T> *rs = xbuf + (size_t)row2 * PITCH;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = rs[fft_read_phys<Cfg>(t2, e)];
            fence_proxy_async();
            mbar_arrive(&empty[s]);
            fft_pass<Cfg, 1, T>(v, t2, tw);
            if (inverse) {
#pragma unroll
                for (int e = 0; e < Cfg::E; e++)
                    v[e] = cplx<T>{ v[e].y * scale, v[e].x * scale };
            }
            cplx<T> *op = data + f * (FRAME) + 16 * tile + row2;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                st_stream(op + (size_t)(t2 + Cfg::S * e) * N1, v[e]);
        }
        mbar_arrive(&done_bar[it % ND]); // (release: this thread's stores are ordered before the data mover's count)
    }
}


// -------------------------------------------------------------------------------------------------
// REAL frames of 65536 points (sdsp_b200_fft_exec_real, forward): the two-slot work queue again, doing half the work.
// The reference's callers put a real signal into the real part of a complex_array and leave the imaginary part zero
// (test/testFFT.cpp:24, :86); the spectrum of such a frame is conjugate-symmetric, X[N - k] = conj(X[k]), and so is every
// column transform Y_b[k1] over a.  Hence
//   column tile (8 per frame, 32 real columns): columns b, b + 1 ride one complex 256-point transform as real and imaginary
//     part, z = x_b + i x_(b+1); Y_b[k1] = (Z[k1] + conj Z[256 - k1]) / 2, Y_(b+1)[k1] = -i (Z[k1] - conj Z[256 - k1]) / 2 -- the
//     mirror term comes from the thread that holds it through one more exchange in the slot -- and only k1 = 0 .. 128 go to the
//     ring (times W_N^(b k1)), two neighbouring columns per 16-byte store;
//   row tile (9 per frame: rows 0 .. 127 in eight tiles, row 128 in a ninth): the usual 256-point transform over b gives
//     X[k1 + 256 k2]; rows 1 .. 127 also store conj to the mirror bin (256 - k1) + 256 (255 - k2), rows 0 and 128 mirror into
//     themselves.  Rows 129 .. 255 are never computed.
// 17 items per frame instead of 32, the output is conjugate-symmetric to the bit.  Queue order: column tiles of frame f + LAG, then
// row tiles of frame f; a row tile waits for its frame's 8 column tiles, a column tile for the 9 row tiles of the ring slot's
// previous tenant -- both hold smaller tickets.
#ifndef SDSP_REAL_NST
#define SDSP_REAL_NST 2 // tile slots per CTA
#endif
#ifndef SDSP_REAL_MINB
#define SDSP_REAL_MINB SDSP_FUSED_TMA_MINB
#endif
// Frames between a frame's column tiles and its row tiles (x 17 items; the ring holds twice that, 51 MB), with 16 MB of the ring pinned
// in L2 (profiles/r02_fft_l2_persist_sweep.txt, r02_fft_lag_traffic.txt).  96 frames for both output forms.  With full spectra out
// a fifth of the ring spills to HBM at that lag (DRAM traffic 1.19 x the algorithmic bytes against 1.01 x at 64 frames) -- HBM is not
// what bounds this kernel, and config 5 runs 4 % faster for it (78.8 against 82.0 ms), so the spill is accepted; with half spectra out
// (half the output, more room in L2) the traffic stays at 1.0 x and 96 frames are 6 % faster than 64.
#ifndef SDSP_REAL_LAG
#define SDSP_REAL_LAG 96
#endif
#ifndef SDSP_REAL_LAG_HALF
#define SDSP_REAL_LAG_HALF 96
#endif
constexpr int REAL_CT = 8, REAL_RT = 9, REAL_ROWS = 129;
constexpr int REAL_RING_MAX = 2 * (SDSP_REAL_LAG > SDSP_REAL_LAG_HALF ? SDSP_REAL_LAG : SDSP_REAL_LAG_HALF); // what the scratch is sized for
template <int REAL_LAG>
__host__ __device__ __forceinline__ void real_decode(size_t q, bool &cols, size_t &f, int &tile)
{
    if (q < (size_t)REAL_LAG * REAL_CT) {
        cols = true;
        f = q / REAL_CT;
        tile = (int)(q % REAL_CT);
        return;
    }
    const size_t r = q - (size_t)REAL_LAG * REAL_CT, u = r / (REAL_CT + REAL_RT);
    const int w = (int)(r % (REAL_CT + REAL_RT));
    cols = w < REAL_CT;
    f = cols ? REAL_LAG + u : u;
    tile = cols ? w : w - REAL_CT;
}

template <typename T, int MINB, int REAL_LAG>
__global__ void __launch_bounds__(288, MINB)
    fft_real64k_kernel(const __grid_constant__ CUtensorMap in_map, cplx<T> *__restrict__ data, cplx<T> *__restrict__ scratch,
                       const cplx<T> *__restrict__ tw, const cplx<T> *__restrict__ tw_hi, const cplx<T> *__restrict__ tw_lo,
                       unsigned *__restrict__ ticket, unsigned *__restrict__ col_done, unsigned *__restrict__ row_done, size_t n_frames, int half)
{
    // half != 0 (sdsp_b200_fft_exec_r2c): only the bins 0 .. 32768 are written, frames 32769 bins apart -- of every row the lower
    // half of k2 directly and the upper half through its mirror bin, which lies in the lower half of the spectrum
    using Cfg = FftCfg<256, 16, 16, 16>; // rows and (packed) columns alike
    constexpr int N1 = 256, N2 = 256, PITCH = LargeStride<Cfg>::value, REAL_RING = 2 * REAL_LAG;
    constexpr int XBUF = 16 * PITCH;
    constexpr size_t FRAME = (size_t)N1 * N2, RFRAME = (size_t)REAL_ROWS * N2; // output frame; ring frame (rows 0 .. 128)
    constexpr uint32_t TILE_BYTES = 4096 * sizeof(cplx<T>);
    constexpr int NST = SDSP_REAL_NST, ND = NST + 1;
    constexpr int SLOT = (XBUF + 15) / 16 * 16;
    extern __shared__ __align__(128) unsigned char smem_raw128[];
    cplx<T> *stage0 = reinterpret_cast<cplx<T> *>(smem_raw128);
    cplx<T> *s_hi = stage0 + (size_t)NST * SLOT, *s_lo = s_hi + N1;
    uint64_t *full = reinterpret_cast<uint64_t *>(s_lo + 256), *empty = full + NST, *done_bar = empty + NST;
    unsigned *s_item = reinterpret_cast<unsigned *>(done_bar + ND);
    if (threadIdx.x < 256) {
        s_hi[threadIdx.x] = tw_hi[threadIdx.x];
        s_lo[threadIdx.x] = tw_lo[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 256);
        }
        for (int s = 0; s < ND; s++)
            mbar_init(&done_bar[s], 256);
        fence_mbar_init();
    }
    __syncthreads();
    const size_t total = (size_t)REAL_LAG * REAL_CT + n_frames * (REAL_CT + REAL_RT);

    if (threadIdx.x >= 256) { // ---- the data mover (as in fft_fused_tma2_kernel)
        if (threadIdx.x != 256)
            return;
        unsigned *pend[ND];
        for (int i = 0; i < ND; i++)
            pend[i] = nullptr;
        unsigned it = 0, next_pub = 0;
        auto count_tile = [&](unsigned j) {
            unsigned *d = pend[j % ND];
            if (d)
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(d) : "memory");
        };
        auto try_publish = [&]() {
            while (next_pub < it && mbar_test(&done_bar[next_pub % ND], (next_pub / ND) & 1)) {
                count_tile(next_pub);
                next_pub++;
            }
        };
        auto wait_dep = [&](const unsigned *ctr, unsigned target) {
            while (ld_acquire_gpu(ctr) < target) {
                try_publish();
                __nanosleep(32);
            }
        };
        for (;; it++) {
            const int s = it % NST;
            if (it >= (unsigned)NST) {
                while (!SDSP_FUSED_POLL(&empty[s], ((it / NST) - 1) & 1))
                    try_publish();
                while (next_pub + ND <= it) {
                    mbar_wait(&done_bar[next_pub % ND], (next_pub / ND) & 1);
                    count_tile(next_pub);
                    next_pub++;
                }
            }
            const size_t q = atomicAdd(ticket, 1u); // one at a time: drawing two or four per atomic is 13 % / 47 % slower (a held ticket delays its dependants)
            bool cols = false;
            size_t f = n_frames;
            int tile = 0;
            if (q < total)
                real_decode<REAL_LAG>(q, cols, f, tile);
            const bool real = q < total && f < n_frames;
            if (real && !cols)
                wait_dep(col_done + f, REAL_CT);
            s_item[s] = q < total ? (unsigned)q : 0xffffffffu;
            if (q >= total) {
                mbar_arrive(&full[s]);
                break;
            }
            cplx<T> *dst = stage0 + (size_t)s * SLOT;
            unsigned *d = nullptr;
            if (!real) {
                mbar_arrive(&full[s]);
            } else if (cols) {
                mbar_expect_tx_only(&full[s], TILE_BYTES);
                tma_load_3d(dst, &in_map, 32 * tile, 0, (int)f, &full[s]); // 256 rows x 32 real columns
                if (f >= (size_t)REAL_RING)
                    wait_dep(row_done + (f - REAL_RING), REAL_RT);
                mbar_arrive(&full[s]);
                d = col_done + f;
            } else {
                const uint32_t bytes = tile == REAL_RT - 1 ? (uint32_t)(N2 * sizeof(cplx<T>)) : TILE_BYTES; // the ninth tile is row 128 alone
                asm volatile("fence.proxy.async.global;" ::: "memory");
                mbar_expect_tx(&full[s], bytes);
                bulk_load_1d(dst, scratch + (f % REAL_RING) * RFRAME + (size_t)(16 * tile) * N2, bytes, &full[s]);
                d = row_done + f;
            }
            pend[it % ND] = d;
        }
        for (; next_pub < it; next_pub++) {
            mbar_wait(&done_bar[next_pub % ND], (next_pub / ND) & 1);
            count_tile(next_pub);
        }
        return;
    }

    // ---- the 256 compute threads
    const int lo16 = threadIdx.x & 15, hi16 = threadIdx.x >> 4;
    // column tiles: packed pair p = lo16 (columns 32 c + 2 p, + 1), thread ct = hi16 of its transform.  Launch-long share of the
    // factors W_N^(b k1), k1 = ct + 16 e, b = 32 c + 2 p:  W^(b k1) = [W^(2p ct) W^(32c ct)] [W^(32 p) W_256^(2c)]^e
    const unsigned x_a = 2u * (unsigned)lo16 * (unsigned)hi16, x_b = 32u * (unsigned)lo16; // < 512
    const cplx<T> w_a = cmul(s_hi[x_a >> 8], s_lo[x_a & 255u]), w_b = cmul(s_hi[x_b >> 8], s_lo[x_b & 255u]);
    for (unsigned it = 0;; it++) {
        const int s = it % NST;
        mbar_wait(&full[s], (it / NST) & 1);
        const unsigned q = s_item[s];
        if (q == 0xffffffffu)
            break;
        bool cols;
        size_t f;
        int tile;
        real_decode<REAL_LAG>(q, cols, f, tile);
        cplx<T> *st = stage0 + (size_t)s * SLOT, *xbuf = st;
        cplx<T> *sc = scratch + (f % REAL_RING) * RFRAME;
        cplx<T> v[Cfg::E];
        if (f >= n_frames) {
            mbar_arrive(&empty[s]);
            mbar_arrive(&done_bar[it % ND]);
            continue;
        }
        if (cols) {
            const int t = hi16, p = lo16;
            // the tile is [256 rows][32 floats]: columns 2p, 2p + 1 of row a are one 8-byte element
            const cplx<T> *rs = st + p;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = rs[(t + Cfg::S * e) * 16];
            cta_sync<1, 256>(); // every thread has its points: the slot becomes the exchange buffer
            cplx<T> *fs = xbuf + (size_t)p * PITCH;
            fft_pass<Cfg, 0, T>(v, t, tw);
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                fs[fft_out_phys<Cfg, 0>(t, e)] = v[e];
            cta_sync<1, 256>();
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = fs[fft_read_phys<Cfg>(t, e)];
            fft_pass<Cfg, 1, T>(v, t, tw); // v[e] = Z[k1], k1 = t + 16 e
            // the mirror terms Z[256 - k1] sit in other threads of the transform: one more exchange, natural order
            // (only the upper half is ever asked for: the mirrors of k1 = 1 .. 128 are 255 .. 128; k1 = 0 mirrors into itself)
            cta_sync<1, 256>(); // (everyone has read the previous exchange)
#pragma unroll
            for (int e = 8; e < Cfg::E; e++)
                fs[Cfg::pad(t + 16 * e)] = v[e];
            cta_sync<1, 256>();
            // factors: first term and ratio for the even column; the odd column's are those times W_N^(k1) (a table look-up with
            // two distinct addresses per warp)
            const unsigned xt = 32u * (unsigned)tile * (unsigned)t;
            cplx<T> w0 = cmul(w_a, cmul(s_hi[xt >> 8], s_lo[xt & 255u]));
            w0 = cplx<T>{ w0.x * (T)0.5, w0.y * (T)0.5 }; // the 1/2 of the separation (exact)
            const cplx<T> r1 = cmul(w_b, s_hi[2 * tile]);
            const cplx<T> r2 = cmul(r1, r1), r3 = cmul(r2, r1), r4 = cmul(r2, r2);
            const cplx<T> q1 = cmul(w0, r4);
            cplx<T> *op = sc + 32 * tile + 2 * p;
#pragma unroll
            for (int e = 0; e <= 8; e++) {
                if (e == 8 && t != 0)
                    break; // k1 = 128 exists for t = 0 only
                const int k1 = t + 16 * e;
                const cplx<T> zm = k1 == 0 ? v[0] : fs[Cfg::pad(256 - k1)];
                const cplx<T> a = v[e], b = cplx<T>{ zm.x, -zm.y };
                const cplx<T> ye = a + b, d = a - b;
                const cplx<T> yo = cplx<T>{ d.y, -d.x }; // -i (a - b)
                const cplx<T> base = e < 4 ? w0 : (e < 8 ? q1 : cmul(q1, r4));
                const cplx<T> we = (e & 3) == 0 ? base : cmul(base, (e & 3) == 1 ? r1 : (e & 3) == 2 ? r2 : r3);
                const cplx<T> wo = cmul(we, s_lo[k1]);
                const cplx<T> o0 = cmul(ye, we), o1 = cmul(yo, wo);
                if constexpr (sizeof(T) == 4)
                    *reinterpret_cast<float4 *>(op + (size_t)k1 * N2) = make_float4(o0.x, o0.y, o1.x, o1.y);
                else {
                    op[(size_t)k1 * N2] = o0;
                    op[(size_t)k1 * N2 + 1] = o1;
                }
            }
            fence_proxy_async();
            mbar_arrive(&empty[s]);
        } else {
            const int t = lo16, row = hi16;
            const bool last_tile = tile == REAL_RT - 1; // row 128 alone: only the first 2 KB of the slot were filled
            const cplx<T> *gp = st + row * N2 + t;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = (last_tile && row != 0) ? cplx<T>{ 0, 0 } : gp[Cfg::S * e];
            cta_sync<1, 256>();
            if constexpr (sizeof(cplx<T>) == 8) { // the tile's lines of the ring are dead: drop them from L2
                if (!last_tile || threadIdx.x < 16)
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(sc + (size_t)(16 * tile) * N2 + (size_t)threadIdx.x * 16) : "memory");
            }
            fft_pass<Cfg, 0, T>(v, t, tw);
            cplx<T> *fs = xbuf + (size_t)row * PITCH;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                fs[fft_out_phys<Cfg, 0>(t, e)] = v[e];
            cta_sync<1, 256>();
            const int row2 = lo16, t2 = hi16;
            const cplx<T> *rs = xbuf + (size_t)row2 * PITCH;
#pragma unroll
            for (int e = 0; e < Cfg::E; e++)
                v[e] = rs[fft_read_phys<Cfg>(t2, e)];
            fence_proxy_async();
            mbar_arrive(&empty[s]);
            fft_pass<Cfg, 1, T>(v, t2, tw);
            const int k1 = 16 * tile + row2;
            if (!last_tile || row2 == 0) {
                const size_t pitch = half ? FRAME / 2 + 1 : FRAME;
                cplx<T> *op = data + f * pitch + k1;
#pragma unroll
                for (int e = 0; e < Cfg::E; e++)
                    if (!half || e < 8 || (e == 8 && t2 == 0 && k1 == 0)) // k2 = t2 + 16 e < 128, and the bin N / 2
                        st_stream(op + (size_t)(t2 + Cfg::S * e) * N1, v[e]);
                if (k1 != 0 && k1 != 128) { // X[N - k] = conj X[k]
                    cplx<T> *mp = data + f * pitch + (256 - k1);
#pragma unroll
                    for (int e = 0; e < Cfg::E; e++)
                        if (!half || e >= 8)
                            st_stream(mp + (size_t)(255 - t2 - Cfg::S * e) * N1, cplx<T>{ v[e].x, -v[e].y });
                }
            }
        }
        mbar_arrive(&done_bar[it % ND]);
    }
}

template <typename T, int N1>
struct FusedTmaCfg {
    static constexpr int NST = SDSP_FUSED_TMA_NST, MINB = SDSP_FUSED_TMA_MINB;
};


// The scratch ring of the queue kernels as a PERSISTING window of L2 (per launch, cudaLaunchAttributeAccessPolicyWindow): part of the
// ring's lines are pinned, so the streaming input and output of the same kernel do not push them out, which is what limited the lead of the
// column phase over the row phase.  The carve-out (cudaLimitPersistingL2CacheSize, device-wide state, set at the first such launch and
// given back by sdsp_b200_shutdown) is 16 MB: the gain peaks there and a carve-out of the ring's size is a LOSS -- what is left of L2
// must still buffer the streams (profiles/r02_fft_l2_persist_sweep.txt).  SDSP_B200_FFT_L2_PERSIST=0 switches the window off,
// SDSP_B200_FFT_L2_PERSIST_MB sizes the carve-out (tuning aids).
static std::mutex g_persist_mu;
static int g_persist_bytes[64], g_persist_window[64];
static bool g_persist_seen[64];

static bool ring_persist_window(int device, void *base, size_t bytes, cudaLaunchAttribute &attr)
{
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("SDSP_B200_FFT_L2_PERSIST");
        mode = e ? atoi(e) : SDSP_FFT_L2_PERSIST_DEFAULT;
    }
    if (mode == 0 || device < 0 || device >= 64)
        return false;
    std::lock_guard<std::mutex> lock(g_persist_mu);
    if (!g_persist_seen[device]) {
        g_persist_seen[device] = true;
        int max_persist = 0;
        g_persist_window[device] = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&g_persist_window[device], cudaDevAttrMaxAccessPolicyWindowSize, device);
        size_t want = (size_t)SDSP_FFT_L2_PERSIST_MB_DEFAULT << 20;
        if (getenv("SDSP_B200_FFT_L2_PERSIST_MB"))
            want = (size_t)atoi(getenv("SDSP_B200_FFT_L2_PERSIST_MB")) << 20;
        if (want < (size_t)max_persist)
            max_persist = (int)want;
        if (max_persist > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) != cudaSuccess) {
            cudaGetLastError();
            max_persist = 0;
        }
        g_persist_bytes[device] = max_persist;
    }
    if (g_persist_bytes[device] <= 0 || g_persist_window[device] <= 0)
        return false;
    const size_t win = bytes < (size_t)g_persist_window[device] ? bytes : (size_t)g_persist_window[device];
    attr.id = cudaLaunchAttributeAccessPolicyWindow;
    attr.val.accessPolicyWindow.base_ptr = base;
    attr.val.accessPolicyWindow.num_bytes = win;
    attr.val.accessPolicyWindow.hitRatio = win <= (size_t)g_persist_bytes[device] ? 1.0f : (float)g_persist_bytes[device] / (float)win;
    attr.val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    return true;
}

// sdsp_b200_shutdown: give the persisting carve-out back on every device that got one
void fft_release_l2_persist()
{
    std::lock_guard<std::mutex> lock(g_persist_mu);
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    for (int d = 0; d < 64; d++) {
        if (g_persist_seen[d] && g_persist_bytes[d] > 0 && cudaSetDevice(d) == cudaSuccess) {
            cudaCtxResetPersistingL2Cache();
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        }
        g_persist_seen[d] = false;
        g_persist_bytes[d] = 0;
    }
    if (have_cur)
        cudaSetDevice(cur);
    cudaGetLastError();
}

// forward real-input frames of 65536 points (fp32): the half-work queue, fft_real64k_kernel
static int launch_real64k(const FftPlan &p, void *data, const void *real_in, size_t n_frames, cudaStream_t stream, bool half = false)
{
    using T = float;
    const size_t need = (1 + 2 * n_frames) * sizeof(unsigned);
    FftPlan &mp = const_cast<FftPlan &>(p);
    if (mp.fused_counter_bytes < need) {
        if (mp.d_fused_counters)
            cudaFree(mp.d_fused_counters);
        mp.d_fused_counters = nullptr;
        mp.fused_counter_bytes = 0;
        if (cudaMalloc(&mp.d_fused_counters, need) != cudaSuccess) {
            cudaGetLastError();
            return set_error(SDSP_B200_ERR_OOM, "fft: cannot allocate %zu bytes of work-queue counters", need);
        }
        mp.fused_counter_bytes = need;
    }
    SDSP_CUDA(cudaMemsetAsync(mp.d_fused_counters, 0, need, stream));
    unsigned *ctr = static_cast<unsigned *>(mp.d_fused_counters);
    const cuuint64_t gdim[3] = { 256, 256, (cuuint64_t)n_frames };
    const cuuint64_t gstride[2] = { 256 * sizeof(T), 256 * 256 * sizeof(T) };
    const cuuint32_t box[3] = { 32, 256, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    CUtensorMap map;
    CUresult r = get_encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(real_in), gdim, gstride, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(SDSP_B200_ERR_CUDA, "fft: cuTensorMapEncodeTiled failed with %d (real input, n=%u frames=%zu)", (int)r, p.n, n_frames);
    const int lag = half ? SDSP_REAL_LAG_HALF : SDSP_REAL_LAG;
    const size_t items = (size_t)lag * REAL_CT + n_frames * (REAL_CT + REAL_RT);
    size_t grid = (size_t)p.sm_count * (size_t)p.real64k_ctas;
    if (grid > items)
        grid = items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(288, 1, 1);
    cfg.dynamicSmemBytes = p.real64k_smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    const size_t ring_bytes = (size_t)2 * lag * REAL_ROWS * 256 * sizeof(cplx<T>);
    if (ring_persist_window(p.device, p.d_scratch, ring_bytes, attr[0])) {
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    auto kern = half ? fft_real64k_kernel<T, SDSP_REAL_MINB, SDSP_REAL_LAG_HALF> : fft_real64k_kernel<T, SDSP_REAL_MINB, SDSP_REAL_LAG>;
    SDSP_CUDA(cudaLaunchKernelEx(&cfg, kern, map, reinterpret_cast<cplx<T> *>(data), reinterpret_cast<cplx<T> *>(p.d_scratch),
                                 reinterpret_cast<const cplx<T> *>(p.d_tw_rows), reinterpret_cast<const cplx<T> *>(p.d_tw_hi),
                                 reinterpret_cast<const cplx<T> *>(p.d_tw_lo), ctr, ctr + 1, ctr + 1 + n_frames, n_frames, half ? 1 : 0));
    return SDSP_B200_OK;
}

template <typename T, int N1>
static int launch_fused_tma(const FftPlan &p, void *data, const void *real_in, size_t n_frames, cudaStream_t stream)
{
    if (n_frames == 0)
        return SDSP_B200_OK;
    constexpr int NST = FusedTmaCfg<T, N1>::NST, MINB = FusedTmaCfg<T, N1>::MINB, COLS = FusedRing<T, N1>::COLS;
    const void *src = real_in ? real_in : data;
    if (reinterpret_cast<uintptr_t>(src) % 16 != 0 || n_frames > 0x7fffffffu) // the tensor map needs a 16-byte-aligned base
        return launch_fused<T, N1>(p, data, real_in, n_frames, stream);
    if constexpr (N1 == 256 && sizeof(T) == 4) {
        if (real_in && p.real64k_ctas > 0 && p.direction == SDSP_B200_FORWARD && reinterpret_cast<uintptr_t>(data) % 16 == 0)
            return launch_real64k(p, data, real_in, n_frames, stream);
    }
    const size_t need = (1 + 2 * n_frames) * sizeof(unsigned);
    FftPlan &mp = const_cast<FftPlan &>(p);
    if (mp.fused_counter_bytes < need) {
        if (mp.d_fused_counters)
            cudaFree(mp.d_fused_counters);
        mp.d_fused_counters = nullptr;
        mp.fused_counter_bytes = 0;
        if (cudaMalloc(&mp.d_fused_counters, need) != cudaSuccess) {
            cudaGetLastError();
            return set_error(SDSP_B200_ERR_OOM, "fft: cannot allocate %zu bytes of work-queue counters", need);
        }
        mp.fused_counter_bytes = need;
    }
    SDSP_CUDA(cudaMemsetAsync(mp.d_fused_counters, 0, need, stream));
    unsigned *ctr = static_cast<unsigned *>(mp.d_fused_counters);
    // the frames as a [frame][a][b] tensor of T: complex input is 2 x 256 values per row, real input 256
    const cuuint64_t inner = real_in ? 256 : 512;
    const cuuint64_t gdim[3] = { inner, (cuuint64_t)N1, (cuuint64_t)n_frames };
    const cuuint64_t gstride[2] = { inner * sizeof(T), inner * sizeof(T) * N1 };
    const cuuint32_t box[3] = { (cuuint32_t)(real_in ? COLS : 2 * COLS), (cuuint32_t)(N1 < 256 ? N1 : 256), 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    CUtensorMap map;
    CUresult r = get_encode_fn()(&map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(src),
                                 gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(SDSP_B200_ERR_CUDA, "fft: cuTensorMapEncodeTiled failed with %d (n=%u frames=%zu)", (int)r, p.n, n_frames);
    const size_t items = ((size_t)FusedRing<T, N1>::LAG + 2 * n_frames) * FusedRing<T, N1>::TILES;
    size_t grid = (size_t)p.sm_count * (size_t)p.ctas_per_sm;
    if (grid > items)
        grid = items;
    cplx<T> *a_data = reinterpret_cast<cplx<T> *>(data), *a_scratch = reinterpret_cast<cplx<T> *>(p.d_scratch);
    const cplx<T> *a_twc = reinterpret_cast<const cplx<T> *>(p.d_tw_cols), *a_twr = reinterpret_cast<const cplx<T> *>(p.d_tw_rows);
    const cplx<T> *a_hi = reinterpret_cast<const cplx<T> *>(p.d_tw_hi), *a_lo = reinterpret_cast<const cplx<T> *>(p.d_tw_lo);
    unsigned *a_col = ctr + 1, *a_row = ctr + 1 + n_frames;
    const int a_real = real_in ? 1 : 0, a_inv = p.direction == SDSP_B200_REVERSE ? 1 : 0;
    const T a_scale = (T)(1.0 / ((double)N1 * 256.0));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(288, 1, 1);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (ring_persist_window(p.device, p.d_scratch, (size_t)FusedRing<T, N1>::RING * N1 * 256 * sizeof(cplx<T>), attr[0])) {
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    if (p.two_slot)
        SDSP_CUDA(cudaLaunchKernelEx(&cfg, fft_fused_tma2_kernel<T, N1, MINB>, map, a_data, a_real, a_scratch, a_twc, a_twr, a_hi, a_lo, ctr, a_col, a_row,
                                     n_frames, a_inv, a_scale));
    else
        SDSP_CUDA(cudaLaunchKernelEx(&cfg, fft_fused_tma_kernel<T, N1, NST, MINB>, map, a_data, a_real, a_scratch, a_twc, a_twr, a_hi, a_lo, ctr, a_col,
                                     a_row, n_frames, a_inv, a_scale));
    return SDSP_B200_OK;
}

static int fused_tma_mode() // SDSP_B200_FFT_FUSED_TMA = 0: compute threads load their own tiles; 1: data mover, one slot; 2: two slots
{
    static int w = -1;
    if (w < 0) {
        const char *e = getenv("SDSP_B200_FFT_FUSED_TMA");
        w = !get_encode_fn() ? 0 : e ? atoi(e) : SDSP_FUSED_TMA_DEFAULT;
    }
    return w;
}
static bool fused_tma_wanted()
{
    return fused_tma_mode() > 0;
}

template <typename T, int N1>
static int setup_fused(FftPlan &p)
{
    using Cfg = FftCfg<256, 16, 16, 16>;
    using CCfg = typename FusedCols<N1>::Cfg;
    constexpr int COLS = FusedRing<T, N1>::COLS;
    constexpr size_t XBUF = 16 * LargeStride<Cfg>::value > COLS * LargeStride<CCfg>::value ? 16 * LargeStride<Cfg>::value : COLS * LargeStride<CCfg>::value;
    p.smem_bytes = (XBUF + N1 + 256) * sizeof(cplx<T>);
    auto kern = fft_fused_kernel<T, N1>;
    if (p.smem_bytes > 48 * 1024)
        SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    int occ = 0;
    SDSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, p.smem_bytes));
    if (occ < 1)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "fft: the fused kernel for n=%u does not fit on an SM", p.n);
    p.ctas_per_sm = occ;
    p.fused = true;
    p.n1 = N1;
    p.npass = CCfg::NPASS + 2;
    p.e = 16;
    p.threads = 256;
    p.scratch_frames = FusedRing<T, N1>::RING;
    if (N1 == 256 && sizeof(T) == 4) { // the real-input kernel's ring (frames of 129 rows) may be the larger one
        const size_t need = ((size_t)REAL_RING_MAX * REAL_ROWS * 256 + 65535) / 65536;
        if (need > p.scratch_frames)
            p.scratch_frames = need;
    }
    const size_t frame_bytes = (size_t)N1 * 256 * sizeof(cplx<T>);
    if (cudaMalloc(&p.d_scratch, p.scratch_frames * frame_bytes) != cudaSuccess) {
        cudaGetLastError();
        return set_error(SDSP_B200_ERR_OOM, "fft: cannot allocate %zu bytes of scratch", p.scratch_frames * frame_bytes);
    }
    std::vector<cplx<T>> tw;
    int ra[4] = { CCfg::R0, CCfg::R1, CCfg::R2, CCfg::R3 }, rb[4] = { 16, 16, 1, 1 };
    build_twiddles<T>(N1, ra, CCfg::NPASS, tw);
    tw.push_back(cplx<T>{ 1, 0 });
    int rc = upload_table<T>(&p.d_tw_cols, tw);
    size_t tw_total = tw.size();
    build_twiddles<T>(256, rb, 2, tw);
    if (!rc)
        rc = upload_table<T>(&p.d_tw_rows, tw);
    tw_total += tw.size();
    std::vector<cplx<T>> hi(N1), lo(256);
    for (int i = 0; i < N1; i++) {
        long double re, im;
        unit_root((uint64_t)i, (uint64_t)N1, re, im);
        hi[i] = cplx<T>{ (T)re, (T)im };
    }
    for (int i = 0; i < 256; i++) {
        long double re, im;
        unit_root((uint64_t)i, (uint64_t)N1 * 256, re, im);
        lo[i] = cplx<T>{ (T)re, (T)im };
    }
    if (!rc)
        rc = upload_table<T>(&p.d_tw_hi, hi);
    if (!rc)
        rc = upload_table<T>(&p.d_tw_lo, lo);
    if (rc)
        return rc;
    p.tw_bytes = (tw_total + hi.size() + lo.size()) * sizeof(cplx<T>);
    p.launch = &launch_fused<T, N1>;
    if constexpr (sizeof(T) == 4) {
        if (fused_tma_wanted()) { // same queue, tiles brought in by a data-mover warp
            constexpr int NST = FusedTmaCfg<T, N1>::NST, MINB = FusedTmaCfg<T, N1>::MINB;
            auto tk = fft_fused_tma_kernel<T, N1, NST, MINB>;
            const size_t smem = p.smem_bytes + (size_t)NST * 4096 * sizeof(cplx<T>) + 128; // stages + barriers and mailboxes
            SDSP_CUDA(cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int tocc = 0;
            SDSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tocc, tk, 288, smem));
            if (tocc >= 1) {
                p.smem_bytes = smem;
                p.ctas_per_sm = tocc;
                p.threads = 288;
                p.staged = true;
                p.launch = &launch_fused_tma<T, N1>;
            }
            // two slots that double as exchange buffers: measured +1.5 % (2^16) / +3.6 % (2^15), -1.7 % at 2^17 (profiles/r02_fft_fused_variants.txt):
            // the default up to N1 = 256; SDSP_B200_FFT_FUSED_TMA=1 / 2 pins the one- / two-slot kernel
            if constexpr (N1 == 256) { // forward real-input frames: the half-work queue (SDSP_B200_FFT_REAL64K=0 keeps the complex kernels)
                const char *e = getenv("SDSP_B200_FFT_REAL64K");
                if (!e || atoi(e) != 0) {
                    auto rk = fft_real64k_kernel<T, SDSP_REAL_MINB, SDSP_REAL_LAG>;
                    auto rkh = fft_real64k_kernel<T, SDSP_REAL_MINB, SDSP_REAL_LAG_HALF>;
                    const size_t slot = ((size_t)16 * LargeStride<Cfg>::value + 15) / 16 * 16;
                    const size_t rsmem = (SDSP_REAL_NST * slot + 512) * sizeof(cplx<T>) + 128;
                    int rocc = 0;
                    if (cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem) == cudaSuccess &&
                        cudaFuncSetAttribute(rkh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem) == cudaSuccess &&
                        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rocc, rk, 288, rsmem) == cudaSuccess && rocc >= 1 &&
                        (size_t)REAL_RING_MAX * REAL_ROWS * 256 <= p.scratch_frames * (size_t)N1 * 256) {
                        p.real64k_ctas = rocc;
                        p.real64k_smem = rsmem;
                    }
                    cudaGetLastError();
                }
            }
            const int mode = fused_tma_mode();
            if (mode == 2 || (mode == SDSP_FUSED_TMA_DEFAULT && !getenv("SDSP_B200_FFT_FUSED_TMA") && N1 <= 256)) {
                auto tk2 = fft_fused_tma2_kernel<T, N1, MINB>;
                const size_t slot = (XBUF + 15) / 16 * 16;
                const size_t smem2 = (2 * slot + N1 + 256) * sizeof(cplx<T>) + 128;
                SDSP_CUDA(cudaFuncSetAttribute(tk2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                int occ2 = 0;
                SDSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, tk2, 288, smem2));
                if (occ2 >= 1) {
                    p.smem_bytes = smem2;
                    p.ctas_per_sm = occ2;
                    p.threads = 288;
                    p.staged = true;
                    p.two_slot = true;
                    p.launch = &launch_fused_tma<T, N1>;
                }
            }
        }
    }
    return SDSP_B200_OK;
}

// cluster size 8 (portable, measured best: profiles/r01_fft65536_cluster_sweep.txt); SDSP_B200_FFT_CLUSTER=16|162 selects the
// 16-CTA variants (tuning aid)
template <typename T>
static int setup_cluster64k_auto(FftPlan &p)
{
    static int pin = -1;
    if (pin < 0) {
        const char *e = getenv("SDSP_B200_FFT_CLUSTER");
        pin = e ? atoi(e) : 0;
    }
    int rc = SDSP_B200_ERR_UNSUPPORTED;
    if constexpr (sizeof(T) == 4) {
        if (pin == 162) // 16-CTA clusters, two CTAs per SM (128 registers, no spills)
            return setup_cluster64k<T, 16, 2>(p);
    }
    if (pin == 16)
        rc = setup_cluster64k<T, 16, sizeof(T) == 4 ? 3 : 1>(p);
    else
        rc = setup_cluster64k<T, 8, sizeof(T) == 4 ? 2 : 1>(p);
    if (rc == SDSP_B200_ERR_UNSUPPORTED)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "fft: the thread-block-cluster kernel for n=65536 cannot be scheduled on this device");
    return rc;
}

template <typename T>
static int setup_large_n1(FftPlan &p)
{
    // n = 65536: SDSP_B200_FFT_65536=fused|cluster|twokernel (comparison aid)
    static int which = -1;
    if (which < 0) {
        const char *e = getenv("SDSP_B200_FFT_65536");
        which = !e ? 0 : !strcmp(e, "cluster") ? 1 : !strcmp(e, "twokernel") ? 2 : 0;
    }
    if (which == 0) { // one persistent kernel, intermediate in an L2-resident ring (n = n1 x 256, n1 = 128 ... 1024)
        switch (p.n / 256) {
        case 32: return setup_fused<T, 32>(p);
        case 64: return setup_fused<T, 64>(p);
        case 128: return setup_fused<T, 128>(p);
        case 256: return setup_fused<T, 256>(p);
        case 512: return setup_fused<T, 512>(p);
        case 1024:
            if constexpr (sizeof(T) == 4)
                return setup_fused<T, 1024>(p);
        default: break;
        }
    }
    if (p.n == 65536 && which == 1)
        return setup_cluster64k_auto<T>(p);
    switch (p.n / 256) {
    case 64: return setup_large<FftCfg<64, 16, 16, 4>, T>(p);
    case 128: return setup_large<FftCfg<128, 16, 16, 8>, T>(p);
    case 256: return setup_large<FftCfg<256, 16, 16, 16>, T>(p);
    case 512: return setup_large<FftCfg<512, 16, 16, 16, 2>, T>(p);
    case 1024:
        if constexpr (sizeof(T) == 4)
            return setup_large<FftCfg<1024, 16, 16, 16, 4>, T>(p);
    default: break;
    }
    return set_error(SDSP_B200_ERR_UNSUPPORTED, "fft: n=%u %s exceeds the multi-pass path (up to 2^18 f32 / 2^17 f64)", p.n,
                     p.precision == SDSP_B200_F32 ? "f32" : "f64");
}

// the factorisation table: log2(n) -> configuration.  16 points per thread wherever the frame has them.
template <int LG>
struct CfgFor;
#define SDSP_FFT_CFG(LG, THREADS_, MINB_, ...)   \
    template <>                                  \
    struct CfgFor<LG> {                          \
        using type = FftCfg<__VA_ARGS__>;        \
        static constexpr int THREADS = THREADS_; \
        static constexpr int MINB = MINB_;       \
    };
SDSP_FFT_CFG(1, 256, 4, 2, 2, 2)
SDSP_FFT_CFG(2, 256, 4, 4, 4, 4)
SDSP_FFT_CFG(3, 256, 4, 8, 8, 8)
SDSP_FFT_CFG(4, 256, 3, 16, 16, 16)
SDSP_FFT_CFG(5, 256, 3, 32, 16, 16, 2)
SDSP_FFT_CFG(6, 256, 3, 64, 16, 16, 4)
SDSP_FFT_CFG(7, 256, 3, 128, 16, 16, 8)
SDSP_FFT_CFG(8, 256, 3, 256, 16, 16, 16)
SDSP_FFT_CFG(9, 256, 3, 512, 16, 16, 16, 2)
SDSP_FFT_CFG(10, 256, 3, 1024, 16, 16, 16, 4)
SDSP_FFT_CFG(11, 256, 3, 2048, 16, 16, 16, 8)
SDSP_FFT_CFG(12, 256, 3, 4096, 16, 16, 16, 16)
SDSP_FFT_CFG(13, 512, 2, 8192, 16, 16, 16, 16, 2)
SDSP_FFT_CFG(14, 1024, 1, 16384, 16, 16, 16, 16, 4)
#undef SDSP_FFT_CFG
// fp32 frames of 8192 / 16384 points: 32 points per thread and a radix-32 first pass make them three passes (two exchanges through
// shared memory per sample, like the 4096-point frames) instead of four (three exchanges).  These kernels sit at the L1 / shared-memory
// data path's limit of about 64 B per cycle and SM, so the fourth crossing is what kept them at 0.68 / 0.64.
struct CfgR32_13 {
    using type = FftCfg<8192, 32, 32, 16, 16, 1, 5>;
    static constexpr int THREADS = 256, MINB = 2;
};
struct CfgR32_14 {
    using type = FftCfg<16384, 32, 32, 32, 16, 1, 5>;
    static constexpr int THREADS = 512, MINB = 1;
};
static bool fft_r32_enabled() // SDSP_B200_FFT_R32=0: the four-pass, 16-points-per-thread kernels (comparison aid)
{
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("SDSP_B200_FFT_R32");
        on = e ? atoi(e) : 1;
    }
    return on != 0;
}

static constexpr int MAX_LOG2N_F32 = 14;
static constexpr int MAX_LOG2N_F64 = 13;

template <int LG>
static int setup_for(FftPlan &p)
{
    using C = CfgFor<LG>;
    if constexpr (LG == 12) { // SDSP_B200_FFT_TUNE: kernel-tuning aid for the headline size
        static int tune = -1;
        if (tune < 0) {
            const char *e = getenv("SDSP_B200_FFT_TUNE");
            tune = e ? atoi(e) : 0;
        }
        if (p.precision == SDSP_B200_F32) {
            if (tune == 1)
                return setup_cta<typename C::type, float, C::THREADS, 3, true>(p);
            if (tune == 2)
                return setup_cta<typename C::type, float, C::THREADS, 4, false>(p);
            if (tune == 3)
                return setup_cta<typename C::type, float, C::THREADS, 4, true>(p);
            if (tune == 4)
                return setup_cta<typename C::type, float, C::THREADS, 2, true>(p);
            if (tune == 5) {
                const int rc = setup_cta<typename C::type, float, C::THREADS, C::MINB>(p);
                return rc ? rc : enable_alias<typename C::type, float, C::THREADS, C::MINB>(p);
            }
        } else {
            if (tune == 1)
                return setup_cta<typename C::type, double, C::THREADS, 2, true>(p);
            if (tune == 2)
                return setup_cta<typename C::type, double, C::THREADS, 3, false>(p);
            if (tune == 3)
                return setup_cta<typename C::type, double, C::THREADS, 1, true>(p);
            if (tune == 4) {
                const int rc = setup_cta<typename C::type, double, C::THREADS, 2>(p);
                return rc ? rc : enable_alias<typename C::type, double, C::THREADS, 2>(p);
            }
        }
    }
    if (p.precision == SDSP_B200_F32) {
        if constexpr (LG == 13) {
            if (fft_r32_enabled())
                return setup_cta<CfgR32_13::type, float, CfgR32_13::THREADS, CfgR32_13::MINB>(p);
        }
        if constexpr (LG == 14) {
            if (fft_r32_enabled()) // (staging the next frame into the idle exchange buffer, fft_cta_alias_kernel: 0.65 against 0.69)
                return setup_cta<CfgR32_14::type, float, CfgR32_14::THREADS, CfgR32_14::MINB>(p);
        }
        return setup_cta<typename C::type, float, C::THREADS, C::MINB>(p);
    }
    if constexpr (LG <= MAX_LOG2N_F64)
    {
        constexpr int MB = LG == 13 ? 1 : (C::MINB > 2 ? 2 : C::MINB); // (8192 points fp64: 128 registers, one CTA)
        const int rc = setup_cta<typename C::type, double, C::THREADS, MB>(p);
        if constexpr (LG >= 8 && LG <= 12) { // next group staged into the exchange buffer (+5 %, profiles/r01_fft_l2_prefetch_comparison.txt)
            static const bool plain = getenv("SDSP_B200_FFT_F64_PLAIN") != nullptr; // comparison aid
            if (rc == SDSP_B200_OK && !plain)
                return enable_alias<typename C::type, double, C::THREADS, MB>(p);
        }
        return rc;
    }
    else
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "fft: n=%u in f64 is larger than one CTA can hold and the multi-pass path is not built", p.n);
}

template <int LG>
static int emulate_for(int precision, bool inverse, void *frame)
{
    using C = CfgFor<LG>;
    if (precision == SDSP_B200_F32) {
        if (LG == 13 && fft_r32_enabled())
            emulate_cfg<CfgR32_13::type, float>(frame, inverse);
        else if (LG == 14 && fft_r32_enabled())
            emulate_cfg<CfgR32_14::type, float>(frame, inverse);
        else
            emulate_cfg<typename C::type, float>(frame, inverse);
    } else
        emulate_cfg<typename C::type, double>(frame, inverse);
    return SDSP_B200_OK;
}


// ---- half-spectrum plans (sdsp_b200_fft_exec_r2c): the M = n / 2 point configuration of the table above behind fft_r2c_kernel
constexpr int R2C_NO_DIRECT_KERNEL = -12345; // (internal) the size has no M-point one-CTA configuration: take the path through the full spectrum
template <class Cfg, typename T, int THREADS, int MINB>
static int launch_r2c(const FftPlan &p, const void *real_in, void *half_out, size_t n_frames, cudaStream_t stream)
{
    constexpr int FPC = THREADS / Cfg::TPF;
    const size_t groups = (n_frames + FPC - 1) / FPC;
    if (groups == 0)
        return SDSP_B200_OK;
    const size_t resident = (size_t)p.sm_count * (size_t)p.r2c_ctas;
    const size_t grid = groups < resident * 4 ? groups : resident * 4;
    const int prefetch = fft_prefetch_enabled() && (Cfg::N * sizeof(cplx<T>)) % 16 == 0 && reinterpret_cast<uintptr_t>(real_in) % 16 == 0 ? 1 : 0;
    fft_r2c_kernel<Cfg, T, THREADS, MINB><<<(unsigned)grid, THREADS, p.r2c_smem, stream>>>(
        static_cast<const cplx<T> *>(real_in), static_cast<cplx<T> *>(half_out), static_cast<const cplx<T> *>(p.d_r2c_tw),
        static_cast<const cplx<T> *>(p.d_r2c_twn), n_frames, prefetch);
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

static int upload_w64()
{
    float2 f[32];
    double2 d[32];
    for (int j = 0; j < 32; j++) {
        long double re, im;
        unit_root((uint64_t)j, 64, re, im);
        f[j] = make_float2((float)re, (float)im);
        d[j] = make_double2((double)re, (double)im);
    }
    SDSP_CUDA(cudaMemcpyToSymbol(c_w64_f32, f, sizeof(f)));
    SDSP_CUDA(cudaMemcpyToSymbol(c_w64_f64, d, sizeof(d)));
    return SDSP_B200_OK;
}

template <class Cfg, typename T, int THREADS, int MINB>
static int launch_c2r(const FftPlan &p, const void *half_in, void *real_out, size_t n_frames, cudaStream_t stream)
{
    constexpr int FPC = THREADS / Cfg::TPF;
    const size_t groups = (n_frames + FPC - 1) / FPC;
    if (groups == 0)
        return SDSP_B200_OK;
    const size_t resident = (size_t)p.sm_count * (size_t)p.r2c_ctas;
    const size_t grid = groups < resident * 4 ? groups : resident * 4;
    fft_c2r_kernel<Cfg, T, THREADS, MINB><<<(unsigned)grid, THREADS, p.r2c_smem, stream>>>(
        static_cast<const cplx<T> *>(half_in), static_cast<cplx<T> *>(real_out), static_cast<const cplx<T> *>(p.d_r2c_tw),
        static_cast<const cplx<T> *>(p.d_r2c_twn), n_frames, fft_prefetch_enabled() && reinterpret_cast<uintptr_t>(half_in) % 16 == 0 ? 1 : 0);
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

template <class Cfg, typename T, int THREADS, int MINB>
static int setup_r2c_cfg(FftPlan &p)
{
    constexpr int FPC = THREADS / Cfg::TPF;
    const bool back = p.direction == SDSP_B200_REVERSE; // reverse plans: half spectrum -> real frame
    const void *kern = back ? reinterpret_cast<const void *>(fft_c2r_kernel<Cfg, T, THREADS, MINB>) :
                              reinterpret_cast<const void *>(fft_r2c_kernel<Cfg, T, THREADS, MINB>);
    p.r2c_smem = (size_t)FPC * Cfg::PADDED_N * sizeof(cplx<T>);
    if (p.r2c_smem > 48 * 1024)
        SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.r2c_smem));
    int occ = 0;
    SDSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, p.r2c_smem));
    if (occ < 1)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "fft r2c kernel for n=%u does not fit on an SM", p.n);
    p.r2c_ctas = occ;
    p.r2c_threads = THREADS;
    p.r2c_fpc = FPC;
    p.r2c_e = Cfg::E;
    p.r2c_npass = Cfg::NPASS;
    int rc = upload_w64();
    if (rc)
        return rc;
    int radices[4] = { Cfg::R0, Cfg::R1, Cfg::R2, Cfg::R3 };
    std::vector<cplx<T>> tw;
    build_twiddles<T>(Cfg::N, radices, Cfg::NPASS, tw);
    tw.push_back(cplx<T>{ 1, 0 }); // (single-pass frames have no table; keep the pointer valid)
    SDSP_CUDA(cudaMalloc(&p.d_r2c_tw, tw.size() * sizeof(cplx<T>)));
    SDSP_CUDA(cudaMemcpy(p.d_r2c_tw, tw.data(), tw.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice));
    std::vector<cplx<T>> wn(Cfg::TPF); // W_n^t, n = 2 M, for the threads' first bins
    for (int t = 0; t < Cfg::TPF; t++) {
        long double re, im;
        unit_root((uint64_t)t, (uint64_t)2 * Cfg::N, re, im);
        wn[t] = cplx<T>{ (T)re, (T)im };
    }
    SDSP_CUDA(cudaMalloc(&p.d_r2c_twn, wn.size() * sizeof(cplx<T>)));
    SDSP_CUDA(cudaMemcpy(p.d_r2c_twn, wn.data(), wn.size() * sizeof(cplx<T>), cudaMemcpyHostToDevice));
    p.r2c_launch = back ? &launch_c2r<Cfg, T, THREADS, MINB> : &launch_r2c<Cfg, T, THREADS, MINB>;
    return SDSP_B200_OK;
}

static int launch_r2c_real64k(const FftPlan &p, const void *real_in, void *half_out, size_t n_frames, cudaStream_t stream)
{
    if (reinterpret_cast<uintptr_t>(real_in) % 16 != 0 || n_frames > 0x7fffffffu)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_exec_r2c: frames of 65536 points need a 16-byte-aligned input");
    return launch_real64k(p, half_out, real_in, n_frames, stream, true);
}

template <int LG>
static int setup_r2c_for(FftPlan &p) // LG = log2(n / 2)
{
    using C = CfgFor<LG>;
    if (p.precision == SDSP_B200_F32) {
        if constexpr (LG == 13)
            return setup_r2c_cfg<CfgR32_13::type, float, CfgR32_13::THREADS, CfgR32_13::MINB>(p);
        else if constexpr (LG == 14)
            return setup_r2c_cfg<CfgR32_14::type, float, CfgR32_14::THREADS, CfgR32_14::MINB>(p);
        else
            return setup_r2c_cfg<typename C::type, float, C::THREADS, C::MINB>(p);
    }
    if constexpr (LG <= MAX_LOG2N_F64) {
        constexpr int MB = LG == 13 ? 1 : (C::MINB > 2 ? 2 : C::MINB);
        return setup_r2c_cfg<typename C::type, double, C::THREADS, MB>(p);
    } else
        return R2C_NO_DIRECT_KERNEL;
}


// host emulation of the half-spectrum kernels (sdsp_b200_debug_emulate_r2c): the M-point frame code of fft_emulate_frame plus the
// separation step with the factors formed the way the kernels form them (W_n^t from the table, times a 64th root)
template <class Cfg, typename T>
static void emulate_half_cfg(const void *in, void *out, bool back)
{
    constexpr int M = Cfg::N;
    int radices[4] = { Cfg::R0, Cfg::R1, Cfg::R2, Cfg::R3 };
    std::vector<cplx<T>> tw;
    build_twiddles<T>(M, radices, Cfg::NPASS, tw);
    tw.push_back(cplx<T>{ 1, 0 });
    auto factor = [&](int t, int e) {
        long double re, im;
        unit_root((uint64_t)t, (uint64_t)2 * M, re, im);
        const cplx<T> wt{ (T)re, (T)im };
        if (e == 0)
            return wt;
        unit_root((uint64_t)(e * (32 / Cfg::E)), 64, re, im);
        return cmul(wt, cplx<T>{ (T)re, (T)im });
    };
    std::vector<cplx<T>> z(M);
    if (!back) {
        const T *x = static_cast<const T *>(in);
        cplx<T> *half = static_cast<cplx<T> *>(out);
        for (int j = 0; j < M; j++)
            z[j] = cplx<T>{ x[2 * j], x[2 * j + 1] };
        fft_emulate_frame<Cfg, T>(z.data(), tw.data(), false, (T)1);
        for (int t = 0; t < Cfg::TPF; t++)
            for (int e = 0; e < Cfg::E; e++) {
                const int k = t + Cfg::S * e;
                half[k] = r2c_bin(z[k], z[(M - k) & (M - 1)], factor(t, e));
            }
        half[M] = cplx<T>{ z[0].x - z[0].y, (T)0 };
    } else {
        const cplx<T> *half = static_cast<const cplx<T> *>(in);
        T *x = static_cast<T *>(out);
        for (int t = 0; t < Cfg::TPF; t++)
            for (int e = 0; e < Cfg::E; e++) {
                const int k = t + Cfg::S * e;
                z[k] = c2r_bin(half[k], half[M - k], factor(t, e));
            }
        fft_emulate_frame<Cfg, T>(z.data(), tw.data(), true, (T)(0.5 / (double)M));
        for (int j = 0; j < M; j++) {
            x[2 * j] = z[j].x;
            x[2 * j + 1] = z[j].y;
        }
    }
}

template <int LG>
static int emulate_half_for(int precision, bool back, const void *in, void *out) // LG = log2(n / 2)
{
    using C = CfgFor<LG>;
    if (precision == SDSP_B200_F32) {
        if (LG == 13)
            emulate_half_cfg<CfgR32_13::type, float>(in, out, back);
        else if (LG == 14)
            emulate_half_cfg<CfgR32_14::type, float>(in, out, back);
        else
            emulate_half_cfg<typename C::type, float>(in, out, back);
    } else
        emulate_half_cfg<typename C::type, double>(in, out, back);
    return SDSP_B200_OK;
}

#define SDSP_FOR_EACH_LG(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14)

static int setup_plan(FftPlan &p)
{
    const int lg = ilog2(p.n);
    static const bool fused_small = getenv("SDSP_B200_FFT_FUSED_SMALL") && atoi(getenv("SDSP_B200_FFT_FUSED_SMALL")) != 0; // comparison aid: 8192 / 16384 through the fused kernel
    if (fused_small && (lg == 13 || lg == 14))
        return p.precision == SDSP_B200_F32 ? setup_large_n1<float>(p) : setup_large_n1<double>(p);
    // (fp32 frames of 16384 points: one 1024-thread CTA per frame, 0.64 of the copy peak since the first pass loads four twiddles
    // instead of fifteen; the two-slot data-mover queue, n = 64 x 256, reaches 0.60 -- profiles/r02_fft_fused_variants.txt,
    // r02_bench_workloads_v1.txt.  8192 points: 0.68 against 0.60.)
    if (lg > (p.precision == SDSP_B200_F32 ? MAX_LOG2N_F32 : MAX_LOG2N_F64))
        return p.precision == SDSP_B200_F32 ? setup_large_n1<float>(p) : setup_large_n1<double>(p);
    switch (lg) {
#define X(LG) \
    case LG: return setup_for<LG>(p);
        SDSP_FOR_EACH_LG(X)
#undef X
    default:
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "fft: n=%u is larger than one CTA can hold and the multi-pass path is not built", p.n);
    }
}

static int emulate_dispatch(uint32_t n, int precision, bool inverse, void *frame)
{
    switch (ilog2(n)) {
#define X(LG) \
    case LG: return emulate_for<LG>(precision, inverse, frame);
        SDSP_FOR_EACH_LG(X)
#undef X
    default: return set_error(SDSP_B200_ERR_UNSUPPORTED, "emulate_fft: n=%u not built", n);
    }
}



// Sizes without a direct half-spectrum kernel (fp32 frames of 2^17 / 2^18 points, fp32 reverse of 2^16, fp64 from 2^15): through
// the plan's own full-length transform and a temporary of full spectra, a slab of frames at a time -- complete rather than fast
// (two extra passes over a slab that stays in L2 when it is small).
template <typename T>
__global__ void half_compact_kernel(const cplx<T> *__restrict__ full, cplx<T> *__restrict__ half, uint32_t n, size_t frames)
{
    const size_t bins = (size_t)n / 2 + 1, total = frames * bins;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / bins, k = i % bins;
        half[i] = full[f * n + k];
    }
}
template <typename T>
__global__ void half_expand_kernel(const cplx<T> *__restrict__ half, cplx<T> *__restrict__ full, uint32_t n, size_t frames)
{
    const size_t bins = (size_t)n / 2 + 1, total = frames * (size_t)n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / n, k = i % n;
        cplx<T> v;
        if (k < bins) {
            v = half[f * bins + k];
            if (k == 0 || k == n / 2)
                v.y = 0; // a real signal's spectrum is real there; whatever the caller left in the imaginary part is ignored
        } else {
            const cplx<T> m = half[f * bins + (n - k)];
            v = cplx<T>{ m.x, -m.y };
        }
        full[i] = v;
    }
}
template <typename T>
__global__ void real_part_kernel(const cplx<T> *__restrict__ full, T *__restrict__ out, size_t total)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = full[i].x;
}

template <typename T>
static int launch_half_through_full(const FftPlan &p, const void *in, void *out, size_t n_frames, cudaStream_t stream)
{
    FftPlan &mp = const_cast<FftPlan &>(p);
    const bool back = p.direction == SDSP_B200_REVERSE;
    const size_t frame_bytes = (size_t)p.n * sizeof(cplx<T>), bins = (size_t)p.n / 2 + 1;
    size_t slab = (256u << 20) / frame_bytes;
    slab = slab < 1 ? 1 : slab > n_frames ? n_frames : slab;
    if (mp.half_tmp_bytes < slab * frame_bytes) {
        if (mp.d_half_tmp)
            cudaFree(mp.d_half_tmp);
        mp.d_half_tmp = nullptr;
        mp.half_tmp_bytes = 0;
        if (cudaMalloc(&mp.d_half_tmp, slab * frame_bytes) != cudaSuccess) {
            cudaGetLastError();
            return set_error(SDSP_B200_ERR_OOM, "fft: cannot allocate %zu bytes for the full-spectrum temporary", slab * frame_bytes);
        }
        mp.half_tmp_bytes = slab * frame_bytes;
    }
    cplx<T> *tmp = static_cast<cplx<T> *>(mp.d_half_tmp);
    const unsigned grid = (unsigned)p.sm_count * 8;
    for (size_t done = 0; done < n_frames; done += slab) {
        const size_t cnt = (n_frames - done) < slab ? (n_frames - done) : slab;
        int rc;
        if (!back) {
            rc = p.launch(p, tmp, static_cast<const T *>(in) + done * p.n, cnt, stream);
            if (rc)
                return rc;
            half_compact_kernel<T><<<grid, 256, 0, stream>>>(tmp, static_cast<cplx<T> *>(out) + done * bins, p.n, cnt);
        } else {
            half_expand_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const cplx<T> *>(in) + done * bins, tmp, p.n, cnt);
            rc = p.launch(p, tmp, nullptr, cnt, stream);
            if (rc)
                return rc;
            real_part_kernel<T><<<grid, 256, 0, stream>>>(tmp, static_cast<T *>(out) + done * p.n, cnt * (size_t)p.n);
        }
        SDSP_CUDA(cudaGetLastError());
    }
    return SDSP_B200_OK;
}

static int setup_r2c(FftPlan &p)
{
    if (p.r2c_ready)
        return SDSP_B200_OK;
    if (p.n < 4)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "fft_exec_r2c: n=%u (needs at least 4 points)", p.n);
    const int lg = ilog2(p.n) - 1;
    int rc;
    if (lg == 15 && p.precision == SDSP_B200_F32 && p.real64k_ctas > 0 && p.direction == SDSP_B200_FORWARD) {
        p.r2c_launch = &launch_r2c_real64k;
        rc = SDSP_B200_OK;
    } else
        switch (lg) {
#define X(LG) \
    case LG: rc = setup_r2c_for<LG>(p); break;
            SDSP_FOR_EACH_LG(X)
#undef X
        default:
            rc = R2C_NO_DIRECT_KERNEL;
        }
    if (rc == R2C_NO_DIRECT_KERNEL) { // any size the plan itself transforms: through the full spectrum
        p.r2c_launch = p.precision == SDSP_B200_F32 ? &launch_half_through_full<float> : &launch_half_through_full<double>;
        p.r2c_through_full = true;
        rc = SDSP_B200_OK;
    }
    if (rc == SDSP_B200_OK)
        p.r2c_ready = true;
    return rc;
}

static int check_fft_args(uint32_t n, int radix, int precision, int direction)
{
    if (radix != 2 && radix != 4)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft: radix must be 2 or 4 (got %d)", radix);
    if (precision != SDSP_B200_F32 && precision != SDSP_B200_F64)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft: bad precision %d", precision);
    if (direction != SDSP_B200_FORWARD && direction != SDSP_B200_REVERSE)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft: bad direction %d", direction);
    if (!is_pow2(n) || n < 2) // reference fft.h:261 "FFT size must be a power of 2!"
        return set_error(SDSP_B200_ERR_INVALID_ARG, "FFT size must be a power of 2! (n=%u)", n);
    if (radix == 4 && (ilog2(n) % 2) != 0) // reference fft.h:304
        return set_error(SDSP_B200_ERR_INVALID_ARG, "FFT radix 4 size must be a power of 4! (n=%u)", n);
    return SDSP_B200_OK;
}
} // namespace sdsp_b200

using namespace sdsp_b200;

struct sdsp_b200_fft_plan_s {
    FftPlan p;
};

// host-side view of the fused kernels' work queue (tests/test_capi_cpu.py checks the dependency argument on it)
template <typename T, int N1>
static void fused_queue_item(unsigned long long q, int *geom, int *item)
{
    using R = FusedRing<T, N1>;
    geom[0] = R::TILES;
    geom[1] = R::COLS;
    geom[2] = R::LAG;
    geom[3] = R::RING;
    bool cols;
    size_t f;
    int tile;
    fused_decode<R::LAG, R::TILES>((size_t)q, cols, f, tile);
    item[0] = cols ? 1 : 0;
    item[1] = tile;
    item[2] = (int)(f & 0x7fffffff);
}

extern "C" {

int sdsp_b200_fft_plan_create(sdsp_b200_fft_plan *plan, uint32_t n, int radix, int precision, int direction, int device)
{
    if (!plan)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_plan_create: null out pointer");
    *plan = nullptr;
    int rc = check_fft_args(n, radix, precision, direction);
    if (rc)
        return rc;
    rc = ensure_device(device);
    if (rc)
        return rc;
    auto *h = new sdsp_b200_fft_plan_s();
    h->p.n = n;
    h->p.radix = radix;
    h->p.precision = precision;
    h->p.direction = direction;
    h->p.device = device;
    h->p.sm_count = device_sm_count(device);
    rc = setup_plan(h->p);
    if (rc) {
        for (void *q : { h->p.d_tw, h->p.d_tw_cols, h->p.d_tw_rows, h->p.d_tw_hi, h->p.d_tw_lo, h->p.d_scratch, h->p.d_fused_counters })
            if (q)
                cudaFree(q);
        delete h;
        return rc;
    }
    *plan = h;
    return SDSP_B200_OK;
}

int sdsp_b200_fft_plan_destroy(sdsp_b200_fft_plan plan)
{
    if (!plan)
        return SDSP_B200_OK;
    cudaSetDevice(plan->p.device);
    if (plan->p.d_tw)
        cudaFree(plan->p.d_tw);
    if (plan->p.d_stage)
        cudaFree(plan->p.d_stage);
    plan->p.host.release();
    for (void *q : { plan->p.d_tw_cols, plan->p.d_tw_rows, plan->p.d_tw_hi, plan->p.d_tw_lo, plan->p.d_scratch, plan->p.d_fused_counters, plan->p.d_r2c_tw,
                     plan->p.d_r2c_twn, plan->p.d_half_tmp })
        if (q)
            cudaFree(q);
    delete plan;
    return SDSP_B200_OK;
}

int sdsp_b200_fft_exec(sdsp_b200_fft_plan plan, void *data, size_t n_frames, int ptr_kind, void *stream)
{
    if (!plan)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_exec: null plan");
    if (n_frames == 0)
        return SDSP_B200_OK;
    if (!data)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_exec: null data");
    FftPlan &p = plan->p;
    SDSP_CUDA(cudaSetDevice(p.device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t elem = (p.precision == SDSP_B200_F32 ? sizeof(float) : sizeof(double)) * 2;
    if ((reinterpret_cast<uintptr_t>(data) % elem) != 0)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_exec: data must be aligned to one complex element (%zu bytes)", elem);
    if (ptr_kind == SDSP_B200_PTR_DEVICE)
        return p.launch(p, data, nullptr, n_frames, s);
    if (ptr_kind != SDSP_B200_PTR_HOST)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_exec: bad ptr_kind %d", ptr_kind);

    // host data (see host_stage.h): synchronous; anything the caller queued on `stream` is waited for first
    std::lock_guard<std::mutex> lock(p.mu);
    if (s)
        SDSP_CUDA(cudaStreamSynchronize(s));
    int rc = p.host.ensure();
    if (rc)
        return rc;
    const size_t frame_bytes = (size_t)p.n * elem;
    const size_t total = n_frames * frame_bytes;
    if (total <= HostStage::BOUNCE_BYTES) { // one frame / a few frames: latency path through the pinned bounce buffer
        rc = ensure_device_stage(p.d_stage, p.stage_bytes, total, "fft_exec");
        if (rc)
            return rc;
        cudaStream_t cs = p.host.stream[0];
        memcpy(p.host.bounce, data, total);
        SDSP_CUDA(cudaMemcpyAsync(p.d_stage, p.host.bounce, total, cudaMemcpyHostToDevice, cs));
        rc = p.launch(p, p.d_stage, nullptr, n_frames, cs);
        if (rc)
            return rc;
        SDSP_CUDA(cudaMemcpyAsync(p.host.bounce, p.d_stage, total, cudaMemcpyDeviceToHost, cs));
        SDSP_CUDA(cudaStreamSynchronize(cs));
        memcpy(data, p.host.bounce, total);
        return SDSP_B200_OK;
    }
    size_t slab_frames = host_slab_bytes() / frame_bytes;
    if (slab_frames < 1)
        slab_frames = 1;
    if (slab_frames > n_frames)
        slab_frames = n_frames;
    const int nbuf = n_frames > slab_frames ? 2 : 1;
    rc = ensure_device_stage(p.d_stage, p.stage_bytes, slab_frames * frame_bytes * nbuf, "fft_exec");
    if (rc)
        return rc;
    size_t done = 0;
    int which = 0;
    bool first = true;
    while (done < n_frames && rc == SDSP_B200_OK) {
        const size_t cnt = (n_frames - done) < slab_frames ? (n_frames - done) : slab_frames;
        char *h = static_cast<char *>(data) + done * frame_bytes;
        char *d = static_cast<char *>(p.d_stage) + (size_t)which * slab_frames * frame_bytes;
        cudaStream_t cs = p.host.stream[which];
        if (cudaMemcpyAsync(d, h, cnt * frame_bytes, cudaMemcpyHostToDevice, cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "H2D", __FILE__, __LINE__);
        // the transforms of consecutive slabs share the plan's scratch ring and counters (n >= 32768): one at a time
        if (rc == SDSP_B200_OK && !first && cudaStreamWaitEvent(cs, p.host.kernel_done[which ^ 1], 0) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK)
            rc = p.launch(p, d, nullptr, cnt, cs);
        if (rc == SDSP_B200_OK && cudaEventRecord(p.host.kernel_done[which], cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "cudaEventRecord", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK && cudaMemcpyAsync(h, d, cnt * frame_bytes, cudaMemcpyDeviceToHost, cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "D2H", __FILE__, __LINE__);
        done += cnt;
        which = (which + 1) % nbuf;
        first = false;
    }
    return p.host.drain(rc);
}

// real frames in, spectra out (out of place): the imaginary part of the input is zero, as in every call site of
// the reference (test/testFFT.cpp:24,86 fill only the real part) -- half the input bytes of fft_exec
int sdsp_b200_fft_exec_real(sdsp_b200_fft_plan plan, const void *real_in, void *spectrum_out, size_t n_frames, int ptr_kind, void *stream)
{
    if (!plan)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_exec_real: null plan");
    if (n_frames == 0)
        return SDSP_B200_OK;
    if (!real_in || !spectrum_out)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_exec_real: null data");
    FftPlan &p = plan->p;
    SDSP_CUDA(cudaSetDevice(p.device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t es = p.precision == SDSP_B200_F32 ? sizeof(float) : sizeof(double);
    if ((reinterpret_cast<uintptr_t>(real_in) % es) != 0 || (reinterpret_cast<uintptr_t>(spectrum_out) % (2 * es)) != 0)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_exec_real: buffers must be aligned to their element size");
    if (ptr_kind == SDSP_B200_PTR_DEVICE)
        return p.launch(p, spectrum_out, real_in, n_frames, s);
    if (ptr_kind != SDSP_B200_PTR_HOST)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_exec_real: bad ptr_kind %d", ptr_kind);
    // host buffers (see host_stage.h): slabs through the plan's staging memory, each buffer = spectrum slab + real slab behind it
    std::lock_guard<std::mutex> lock(p.mu);
    if (s)
        SDSP_CUDA(cudaStreamSynchronize(s));
    int rc = p.host.ensure();
    if (rc)
        return rc;
    const size_t in_bytes = (size_t)p.n * es, out_bytes = 2 * in_bytes;
    size_t slab = host_slab_bytes() / out_bytes;
    slab = slab < 1 ? 1 : slab > n_frames ? n_frames : slab;
    const int nbuf = n_frames > slab ? 2 : 1;
    const size_t buf_bytes = slab * (in_bytes + out_bytes);
    rc = ensure_device_stage(p.d_stage, p.stage_bytes, buf_bytes * nbuf, "fft_exec_real");
    if (rc)
        return rc;
    int which = 0;
    bool first = true;
    for (size_t done = 0; done < n_frames && rc == SDSP_B200_OK; done += slab) {
        const size_t cnt = (n_frames - done) < slab ? (n_frames - done) : slab;
        char *d_out = static_cast<char *>(p.d_stage) + (size_t)which * buf_bytes, *d_in = d_out + slab * out_bytes;
        cudaStream_t cs = p.host.stream[which];
        if (cudaMemcpyAsync(d_in, static_cast<const char *>(real_in) + done * in_bytes, cnt * in_bytes, cudaMemcpyHostToDevice, cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "H2D", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK && !first && cudaStreamWaitEvent(cs, p.host.kernel_done[which ^ 1], 0) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK)
            rc = p.launch(p, d_out, d_in, cnt, cs);
        if (rc == SDSP_B200_OK && cudaEventRecord(p.host.kernel_done[which], cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "cudaEventRecord", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK &&
            cudaMemcpyAsync(static_cast<char *>(spectrum_out) + done * out_bytes, d_out, cnt * out_bytes, cudaMemcpyDeviceToHost, cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "D2H", __FILE__, __LINE__);
        which = (which + 1) % nbuf;
        first = false;
    }
    return p.host.drain(rc);
}


// real frames <-> half spectra (n / 2 + 1 bins per frame, frames n / 2 + 1 bins apart), out of place.  back = false: real in, half
// spectra out (forward plans); back = true: half spectra in, real frames out (reverse plans, 1/n included)
static int exec_half(sdsp_b200_fft_plan plan, const void *in, void *out, size_t n_frames, int ptr_kind, void *stream, bool back)
{
    const char *who = back ? "fft_exec_c2r" : "fft_exec_r2c";
    if (!plan)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "%s: null plan", who);
    if (n_frames == 0)
        return SDSP_B200_OK;
    if (!in || !out)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "%s: null data", who);
    FftPlan &p = plan->p;
    if ((p.direction == SDSP_B200_REVERSE) != back)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "%s: the plan must be a %s one", who, back ? "reverse" : "forward");
    SDSP_CUDA(cudaSetDevice(p.device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t es = p.precision == SDSP_B200_F32 ? sizeof(float) : sizeof(double);
    if (ptr_kind != SDSP_B200_PTR_DEVICE && ptr_kind != SDSP_B200_PTR_HOST)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "%s: bad ptr_kind %d", who, ptr_kind);
    const void *real_side = back ? out : in, *cplx_side = back ? in : out;
    if ((reinterpret_cast<uintptr_t>(cplx_side) % (2 * es)) != 0 ||
        (reinterpret_cast<uintptr_t>(real_side) % (ptr_kind == SDSP_B200_PTR_DEVICE ? 2 * es : es)) != 0)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "%s: device buffers must be aligned to one complex element (the real frame is accessed in pairs)", who);
    std::lock_guard<std::mutex> lock(p.mu);
    int rc = setup_r2c(p);
    if (rc)
        return rc;
    if (ptr_kind == SDSP_B200_PTR_DEVICE)
        return p.r2c_launch(p, in, out, n_frames, s);
    // host buffers: slabs through the plan's staging memory, each buffer = output slab + input slab behind it
    if (s)
        SDSP_CUDA(cudaStreamSynchronize(s));
    rc = p.host.ensure();
    if (rc)
        return rc;
    const size_t real_bytes = (size_t)p.n * es, half_bytes = ((size_t)p.n / 2 + 1) * 2 * es;
    const size_t in_bytes = back ? half_bytes : real_bytes, out_bytes = back ? real_bytes : half_bytes;
    size_t slab = host_slab_bytes() / (in_bytes + out_bytes);
    slab = slab < 1 ? 1 : slab > n_frames ? n_frames : slab;
    const int nbuf = n_frames > slab ? 2 : 1;
    const size_t out_slab = (slab * out_bytes + 255) / 256 * 256, in_slab = (slab * in_bytes + 255) / 256 * 256;
    const size_t buf_bytes = out_slab + in_slab;
    rc = ensure_device_stage(p.d_stage, p.stage_bytes, buf_bytes * nbuf, who);
    if (rc)
        return rc;
    int which = 0;
    bool first = true;
    for (size_t done = 0; done < n_frames && rc == SDSP_B200_OK; done += slab) {
        const size_t cnt = (n_frames - done) < slab ? (n_frames - done) : slab;
        char *d_out = static_cast<char *>(p.d_stage) + (size_t)which * buf_bytes, *d_in = d_out + out_slab;
        cudaStream_t cs = p.host.stream[which];
        if (cudaMemcpyAsync(d_in, static_cast<const char *>(in) + done * in_bytes, cnt * in_bytes, cudaMemcpyHostToDevice, cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "H2D", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK && !first && cudaStreamWaitEvent(cs, p.host.kernel_done[which ^ 1], 0) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK)
            rc = p.r2c_launch(p, d_in, d_out, cnt, cs);
        if (rc == SDSP_B200_OK && cudaEventRecord(p.host.kernel_done[which], cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "cudaEventRecord", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK && cudaMemcpyAsync(static_cast<char *>(out) + done * out_bytes, d_out, cnt * out_bytes, cudaMemcpyDeviceToHost, cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "D2H", __FILE__, __LINE__);
        which = (which + 1) % nbuf;
        first = false;
    }
    return p.host.drain(rc);
}

int sdsp_b200_fft_exec_r2c(sdsp_b200_fft_plan plan, const void *real_in, void *half_spectrum_out, size_t n_frames, int ptr_kind, void *stream)
{
    return exec_half(plan, real_in, half_spectrum_out, n_frames, ptr_kind, stream, false);
}

int sdsp_b200_fft_exec_c2r(sdsp_b200_fft_plan plan, const void *half_spectrum_in, void *real_out, size_t n_frames, int ptr_kind, void *stream)
{
    return exec_half(plan, half_spectrum_in, real_out, n_frames, ptr_kind, stream, true);
}

int sdsp_b200_fft_plan_describe(sdsp_b200_fft_plan plan, char *buf, size_t buf_len)
{
    if (!plan || !buf || buf_len == 0)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_plan_describe: bad arguments");
    const FftPlan &p = plan->p;
    if (p.fused) {
        const int tiles = p.n1 / 16, cols = 256 / tiles;
        snprintf(buf, buf_len,
                 "fft n=%u %s %s radix-arg=%d: one persistent kernel, single pass over HBM: n = %d x 256; per frame %d column tiles (%d columns "
                 "x %d-point transforms, twiddle) -> ring of %zu scratch frames resident in L2 -> %d row tiles (16 rows x 256-point "
                 "transforms, threads re-mapped between the passes), ordered by an atomic work queue with per-frame completion counters; "
                 "%d threads/CTA%s, smem/CTA=%zuB, CTAs/SM=%d, SMs=%d",
                 p.n, p.precision == SDSP_B200_F32 ? "f32" : "f64", p.direction == SDSP_B200_FORWARD ? "forward" : "reverse", p.radix, p.n1,
                 tiles, cols, p.n1, p.scratch_frames, tiles, p.threads, p.staged ? " (256 compute + a TMA data-mover warp)" : "", p.smem_bytes,
                 p.ctas_per_sm, p.sm_count);
        return SDSP_B200_OK;
    }
    if (p.cluster) {
        snprintf(buf, buf_len,
                 "fft n=%u %s %s radix-arg=%d: one frame per %d-CTA thread-block cluster, single pass over HBM: 256 column transforms "
                 "(%d per CTA, read from HBM) -> twiddle -> rows exchanged through distributed shared memory -> 256 row transforms "
                 "(%d per CTA) -> HBM; 256 threads/CTA, smem/CTA=%zuB, %d clusters resident, SMs=%d",
                 p.n, p.precision == SDSP_B200_F32 ? "f32" : "f64", p.direction == SDSP_B200_FORWARD ? "forward" : "reverse", p.radix,
                 p.cluster_ctas, 256 / p.cluster_ctas, 256 / p.cluster_ctas, p.smem_bytes, p.cluster_slots, p.sm_count);
        return SDSP_B200_OK;
    }
    if (p.large) {
        snprintf(buf, buf_len,
                 "fft n=%u %s %s radix-arg=%d: multi-pass, n = %d x 256: column pass (%d-point transforms, 16 columns/CTA, %d threads, "
                 "smem %zuB) -> L2-resident scratch slab of %zu frames -> row pass (256-point transforms, transposed store, smem %zuB); "
                 "2 launches per slab; SMs=%d",
                 p.n, p.precision == SDSP_B200_F32 ? "f32" : "f64", p.direction == SDSP_B200_FORWARD ? "forward" : "reverse", p.radix, p.n1,
                 p.n1, p.large_threads_a, p.large_smem_a, p.scratch_frames, p.large_smem_b, p.sm_count);
        return SDSP_B200_OK;
    }
    snprintf(buf, buf_len,
             "fft n=%u %s %s radix-arg=%d: single-CTA kernel, passes=%d radices=[%d,%d,%d,%d] points/thread=%d "
             "threads/CTA=%d frames/CTA=%d smem/CTA=%zuB%s CTAs/SM=%d SMs=%d twiddle-table=%zuB",
             p.n, p.precision == SDSP_B200_F32 ? "f32" : "f64", p.direction == SDSP_B200_FORWARD ? "forward" : "reverse", p.radix,
             p.npass, p.radices[0], p.radices[1], p.radices[2], p.radices[3], p.e, p.threads, p.frames_per_cta, p.smem_bytes,
             p.double_buffered ? " (double-buffered exchange)" : p.staged ? " (next group staged into the exchange buffer)" : "", p.ctas_per_sm,
             p.sm_count, p.tw_bytes);
    return SDSP_B200_OK;
}

int sdsp_b200_debug_fft_queue_item(unsigned n, int precision, unsigned long long q, int *geom, int *item)
{
    if (!geom || !item || (precision != SDSP_B200_F32 && precision != SDSP_B200_F64) || n % 256 != 0)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "debug_fft_queue_item: bad arguments");
    const bool f32 = precision == SDSP_B200_F32;
    switch (n / 256) {
#define X(N1) \
    case N1: \
        if (f32) \
            fused_queue_item<float, N1>(q, geom, item); \
        else \
            fused_queue_item<double, N1>(q, geom, item); \
        return SDSP_B200_OK;
        X(32) X(64) X(128) X(256) X(512) X(1024)
#undef X
    default: return set_error(SDSP_B200_ERR_UNSUPPORTED, "debug_fft_queue_item: n=%u is not a size the work-queue kernels take", n);
    }
}

// the same view of the real-input 65536-point kernel's queue (8 column + 9 row tiles per frame): geom = {column tiles, row tiles, lag, ring}
int sdsp_b200_debug_fft_real_queue_item(int half_spectrum, unsigned long long q, int *geom, int *item)
{
    if (!geom || !item)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "debug_fft_real_queue_item: bad arguments");
    const int lag = half_spectrum ? SDSP_REAL_LAG_HALF : SDSP_REAL_LAG;
    geom[0] = REAL_CT;
    geom[1] = REAL_RT;
    geom[2] = lag;
    geom[3] = 2 * lag;
    bool cols = false;
    size_t f = 0;
    int tile = 0;
    if (half_spectrum)
        real_decode<SDSP_REAL_LAG_HALF>((size_t)q, cols, f, tile);
    else
        real_decode<SDSP_REAL_LAG>((size_t)q, cols, f, tile);
    item[0] = cols ? 1 : 0;
    item[1] = tile;
    item[2] = (int)f;
    return SDSP_B200_OK;
}

int sdsp_b200_fft_plan_launches(sdsp_b200_fft_plan plan, size_t n_frames, int *launches)
{
    if (!plan || !launches)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "fft_plan_launches: bad arguments");
    if (plan->p.large)
        *launches = (int)(2 * ((n_frames + plan->p.scratch_frames - 1) / plan->p.scratch_frames));
    else
        *launches = n_frames ? 1 : 0;
    return SDSP_B200_OK;
}

int sdsp_b200_digit_reverse_table(uint32_t n, uint32_t base, int half_table, uint32_t *out, int device)
{
    if (!out)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "digit_reverse_table: null out");
    if (base != 2 && base != 4)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "digit_reverse_table: base must be 2 or 4");
    if (!is_pow2(n) || n < 2 || (base == 4 && ilog2(n) % 2))
        return set_error(SDSP_B200_ERR_INVALID_ARG, "digit_reverse_table: n=%u is not a power of %u", n, base);
    int rc = ensure_device(device);
    if (rc)
        return rc;
    uint32_t *d = nullptr;
    SDSP_CUDA(cudaMalloc(&d, (size_t)n * sizeof(uint32_t)));
    digit_reverse_table_kernel<<<(n + 255) / 256, 256>>>(d, n, ilog2(n), base, half_table);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpy(out, d, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess)
        return cuda_fail((int)e, "digit_reverse_table", __FILE__, __LINE__);
    return SDSP_B200_OK;
}

int sdsp_b200_digit_reverse_permute(void *data, uint32_t n, uint32_t base, int precision, size_t n_frames, int ptr_kind,
                                    int device, void *stream)
{
    if (base != 2 && base != 4)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "digit_reverse_permute: base must be 2 or 4");
    if (!is_pow2(n) || n < 2 || (base == 4 && ilog2(n) % 2))
        return set_error(SDSP_B200_ERR_INVALID_ARG, "digit_reverse_permute: n=%u is not a power of %u", n, base);
    if (n_frames == 0)
        return SDSP_B200_OK;
    if (!data)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "digit_reverse_permute: null data");
    int rc = ensure_device(device);
    if (rc)
        return rc;
    const size_t elem = (precision == SDSP_B200_F32 ? sizeof(float) : sizeof(double)) * 2;
    const size_t bytes = n_frames * (size_t)n * elem;
    void *d = data;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (ptr_kind == SDSP_B200_PTR_HOST) {
        SDSP_CUDA(cudaMalloc(&d, bytes));
        cudaError_t e = cudaMemcpy(d, data, bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d);
            return cuda_fail((int)e, "H2D", __FILE__, __LINE__);
        }
        s = nullptr;
    }
    const size_t total = n_frames * (size_t)n;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (precision == SDSP_B200_F32)
        digit_reverse_permute_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<cplx<float> *>(d), n, ilog2(n), base, n_frames);
    else
        digit_reverse_permute_kernel<double><<<grid, 256, 0, s>>>(reinterpret_cast<cplx<double> *>(d), n, ilog2(n), base, n_frames);
    cudaError_t e = cudaGetLastError();
    if (ptr_kind == SDSP_B200_PTR_HOST) {
        if (e == cudaSuccess)
            e = cudaMemcpy(data, d, bytes, cudaMemcpyDeviceToHost);
        cudaFree(d);
    }
    if (e != cudaSuccess)
        return cuda_fail((int)e, "digit_reverse_permute", __FILE__, __LINE__);
    return SDSP_B200_OK;
}

// W[i][j] = exp(-i * Sign * 2 pi j / 2^(i+1)), i < log2(n), j < n: the table of reference fft.h:197-214,
// produced by the generator that fills the device tables (host only, no device needed)
int sdsp_b200_twiddle_table(uint32_t n, int direction, double *out)
{
    if (!out || !is_pow2(n) || n < 2)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "twiddle_table: n=%u must be a power of 2 and out non-null", n);
    if (direction != SDSP_B200_FORWARD && direction != SDSP_B200_REVERSE)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "twiddle_table: bad direction %d", direction);
    const int rows = ilog2(n);
    for (int i = 0; i < rows; i++) {
        const uint64_t den = 2ull << i;
        for (uint32_t j = 0; j < n; j++) {
            long double re, im;
            unit_root(j, den, re, im);
            out[2 * ((size_t)i * n + j)] = (double)re;
            out[2 * ((size_t)i * n + j) + 1] = direction == SDSP_B200_REVERSE ? (double)-im : (double)im;
        }
    }
    return SDSP_B200_OK;
}

// host emulation of sdsp_b200_fft_exec_r2c (back = 0) / _c2r (back = 1) for the sizes with a direct kernel: n = 4 .. 32768
int sdsp_b200_debug_emulate_r2c(uint32_t n, int precision, int back, const void *in, void *out, size_t n_frames)
{
    if (!in || !out || (precision != SDSP_B200_F32 && precision != SDSP_B200_F64) || !is_pow2(n) || n < 4 || n > 32768)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "debug_emulate_r2c: bad arguments (n = 4 .. 32768, a power of two)");
    const size_t es = precision == SDSP_B200_F32 ? sizeof(float) : sizeof(double);
    const size_t real_bytes = (size_t)n * es, half_bytes = ((size_t)n / 2 + 1) * 2 * es;
    for (size_t f = 0; f < n_frames; f++) {
        const char *ip = static_cast<const char *>(in) + f * (back ? half_bytes : real_bytes);
        char *op = static_cast<char *>(out) + f * (back ? real_bytes : half_bytes);
        switch (ilog2(n) - 1) {
#define X(LG) \
    case LG: emulate_half_for<LG>(precision, back != 0, ip, op); break;
            SDSP_FOR_EACH_LG(X)
#undef X
        default: return set_error(SDSP_B200_ERR_UNSUPPORTED, "debug_emulate_r2c: n=%u not built", n);
        }
    }
    return SDSP_B200_OK;
}

int sdsp_b200_debug_emulate_fft(uint32_t n, int precision, int direction, void *data, size_t n_frames)
{
    int rc = check_fft_args(n, 2, precision, direction);
    if (rc)
        return rc;
    const size_t elem = (precision == SDSP_B200_F32 ? sizeof(float) : sizeof(double)) * 2;
    for (size_t f = 0; f < n_frames; f++) {
        rc = emulate_dispatch(n, precision, direction == SDSP_B200_REVERSE, static_cast<char *>(data) + f * (size_t)n * elem);
        if (rc)
            return rc;
    }
    return SDSP_B200_OK;
}
}
