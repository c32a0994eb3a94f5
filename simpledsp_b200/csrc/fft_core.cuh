// fft_core.cuh -- per-thread arithmetic of the batched FFT kernels (host + device).
//
// What it replaces: the butterfly loops of sdsp::fft_radix2 (reference include/sdsp/fft.h:276-294)
// and sdsp::fft_radix4 (:311-349) together with their digit-reversal sweeps (:269-273, :351-355).
//
// How (B200 design, not the reference's): a frame of N = R0*R1*...*R(p-1) points is transformed by p
// register-resident passes.  Every thread owns E points; in pass i it performs E/Ri radix-Ri
// butterflies entirely in registers.  Between passes the frame is exchanged through shared memory in
// the Stockham auto-sort arrangement, so
//   * every pass reads  position  t + (N/E)*e            (t = thread in frame, e < E)
//   * pass i writes     position  K + Pi*k + Pi*Ri*m      (Pi = R0*..*R(i-1), K = b mod Pi, m = b div Pi,
//                                                          b = t + (N/E)*q the butterfly, k its output)
// and the last pass writes natural-order output with the same coalesced pattern the first pass read
// with.  The digit reversal of the reference is therefore absorbed into the exchange addressing; no
// separate permutation pass touches memory.  Twiddles W_{N/Pi}^(m*k) come from per-pass tables laid
// out [k-1][m] so that a warp reads them contiguously.
//
// Everything here is __host__ __device__ so that sdsp_b200_debug_emulate_fft can run the identical
// code on the CPU (index arithmetic and rounding can be checked without a GPU).
#pragma once
#include "common.h"

#ifndef SDSP_FFT_TW4_F32
#define SDSP_FFT_TW4_F32 1 // first-pass factors from four loads + products in fp32 as well as in fp64 (see fft_pass)
#endif
#ifndef SDSP_FFT_TW4_MIN_M
#define SDSP_FFT_TW4_MIN_M 64 // ... for tables of at least this many entries per row (1024-point frames and up; 256 and 64 measured: profiles/r02_fft_twiddle_loads_ab.txt)
#endif

namespace sdsp_b200
{
// ------------------------------------------------------------------------------------------------
// factorisation of one frame
// PADSH_: the exchange buffer carries one spare element after every 2^PADSH_ (16 for the radix-16 factorisations; 32 where the first
// pass is a radix-32 one, whose writes stride by 32 elements)
template <int N_, int E_, int R0_, int R1_ = 1, int R2_ = 1, int R3_ = 1, int PADSH_ = 4>
struct FftCfg {
    static constexpr int N = N_;  // points per frame
    static constexpr int E = E_;  // points per thread
    static constexpr int R0 = R0_, R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int NPASS = (R0_ > 1) + (R1_ > 1) + (R2_ > 1) + (R3_ > 1);
    static constexpr int TPF = N_ / E_; // threads per frame
    static constexpr int S = N_ / E_;   // distance between the points a thread owns
    static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
    static_assert(R0_ <= E_ && R1_ <= E_ && R2_ <= E_ && R3_ <= E_, "radix larger than points per thread");
    static_assert(N_ % E_ == 0, "E must divide N");

    SDSP_HD static constexpr int radix(int p)
    {
        return p == 0 ? R0_ : p == 1 ? R1_ : p == 2 ? R2_ : R3_;
    }
    // product of the radices of the passes before p
    SDSP_HD static constexpr int pprev(int p)
    {
        return p == 0 ? 1 : p == 1 ? R0_ : p == 2 ? R0_ * R1_ : R0_ * R1_ * R2_;
    }
    // entries of the twiddle table of pass p: (R-1) rows of N/(pprev*R) values; the last pass has none
    SDSP_HD static constexpr int tw_count(int p)
    {
        return p + 1 >= NPASS ? 0 : (radix(p) - 1) * (N_ / (pprev(p) * radix(p)));
    }
    SDSP_HD static constexpr int tw_offset(int p)
    {
        return p == 0 ? 0 : tw_offset(p - 1) + tw_count(p - 1);
    }
    // shared-memory exchange: one spare element after every 16 keeps both the strided writes of a pass
    // and the unit-stride reads of the next one free of bank conflicts (8-byte and 16-byte elements)
    static constexpr int PADSH = PADSH_, PADW = 1 << PADSH_;
    SDSP_HD static constexpr int pad(int pos)
    {
        return pos + (pos >> PADSH_);
    }
    static constexpr int PADDED_N = N_ + (N_ >> PADSH_);
};

// ------------------------------------------------------------------------------------------------
// butterflies, forward sign (e^{-i...}); natural-order in, natural-order out, in place

template <typename T>
SDSP_HD cplx<T> mul_neg_i(cplx<T> a) // a * (-i)
{
    return { a.y, -a.x };
}

template <typename T>
SDSP_HD void dft2(cplx<T> &a, cplx<T> &b)
{
    const cplx<T> t = a;
    a = t + b;
    b = t - b;
}

template <typename T>
SDSP_HD void dft4(cplx<T> &a0, cplx<T> &a1, cplx<T> &a2, cplx<T> &a3)
{
    const cplx<T> s02 = a0 + a2, d02 = a0 - a2;
    const cplx<T> s13 = a1 + a3, d13 = mul_neg_i(a1 - a3);
    a0 = s02 + s13;
    a1 = d02 + d13;
    a2 = s02 - s13;
    a3 = d02 - d13;
}

template <typename T>
struct FftConst {
    static constexpr T SQRT1_2 = (T)0.70710678118654752440084436210484903928L;
    static constexpr T COS_PI_8 = (T)0.92387953251128675612818318939678828682L;
    static constexpr T SIN_PI_8 = (T)0.38268343236508977172845998403039886676L;
};

template <int R, typename T>
struct Dft;

template <typename T>
struct Dft<2, T> {
    SDSP_HD static void run(cplx<T> (&a)[2])
    {
        dft2(a[0], a[1]);
    }
};

template <typename T>
struct Dft<4, T> {
    SDSP_HD static void run(cplx<T> (&a)[4])
    {
        dft4(a[0], a[1], a[2], a[3]);
    }
};

template <typename T>
struct Dft<8, T> {
    SDSP_HD static void run(cplx<T> (&a)[8])
    {
        constexpr T h = FftConst<T>::SQRT1_2;
        // decimation in frequency: even outputs from sums, odd outputs from twiddled differences
        cplx<T> e0 = a[0] + a[4], e1 = a[1] + a[5], e2 = a[2] + a[6], e3 = a[3] + a[7];
        cplx<T> o0 = a[0] - a[4], d1 = a[1] - a[5], d2 = a[2] - a[6], d3 = a[3] - a[7];
        cplx<T> o1 = { h * (d1.x + d1.y), h * (d1.y - d1.x) };  // * W8^1 = (1 - i)/sqrt2
        cplx<T> o2 = mul_neg_i(d2);                              // * W8^2 = -i
        cplx<T> o3 = { h * (d3.y - d3.x), -(h * (d3.x + d3.y)) }; // * W8^3 = (-1 - i)/sqrt2
        dft4(e0, e1, e2, e3);
        dft4(o0, o1, o2, o3);
        a[0] = e0;
        a[1] = o0;
        a[2] = e1;
        a[3] = o1;
        a[4] = e2;
        a[5] = o2;
        a[6] = e3;
        a[7] = o3;
    }
};

template <typename T>
struct Dft<16, T> {
    SDSP_HD static void run(cplx<T> (&a)[16])
    {
        constexpr T h = FftConst<T>::SQRT1_2, c = FftConst<T>::COS_PI_8, s = FftConst<T>::SIN_PI_8;
        // n = b + 4c, k = k1 + 4 k2:  X[k1 + 4 k2] = sum_b W4^(b k2) [ W16^(b k1) sum_c a[b + 4c] W4^(c k1) ]
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int b = 0; b < 4; b++)
            dft4(a[b], a[b + 4], a[b + 8], a[b + 12]); // a[b + 4 k1] = inner sum
        // W16^(b k1), b, k1 in 1..3 : exponents 1 2 3 / 2 4 6 / 3 6 9
        a[1 + 4] = cmul(a[1 + 4], cplx<T>{ c, -s });
        a[1 + 8] = cplx<T>{ h * (a[1 + 8].x + a[1 + 8].y), h * (a[1 + 8].y - a[1 + 8].x) };
        a[1 + 12] = cmul(a[1 + 12], cplx<T>{ s, -c });
        a[2 + 4] = cplx<T>{ h * (a[2 + 4].x + a[2 + 4].y), h * (a[2 + 4].y - a[2 + 4].x) };
        a[2 + 8] = mul_neg_i(a[2 + 8]);
        a[2 + 12] = cplx<T>{ h * (a[2 + 12].y - a[2 + 12].x), -(h * (a[2 + 12].x + a[2 + 12].y)) };
        a[3 + 4] = cmul(a[3 + 4], cplx<T>{ s, -c });
        a[3 + 8] = cplx<T>{ h * (a[3 + 8].y - a[3 + 8].x), -(h * (a[3 + 8].x + a[3 + 8].y)) };
        a[3 + 12] = cmul(a[3 + 12], cplx<T>{ -c, s });
        cplx<T> o[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k1 = 0; k1 < 4; k1++) {
            cplx<T> x0 = a[0 + 4 * k1], x1 = a[1 + 4 * k1], x2 = a[2 + 4 * k1], x3 = a[3 + 4 * k1];
            dft4(x0, x1, x2, x3); // over b -> k2
            o[k1] = x0;
            o[k1 + 4] = x1;
            o[k1 + 8] = x2;
            o[k1 + 12] = x3;
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 16; k++)
            a[k] = o[k];
    }
};

template <typename T>
struct Dft<32, T> {
    // decimation in frequency once, then two 16-point transforms: even outputs from a[j] + a[j + 16], odd outputs from
    // (a[j] - a[j + 16]) W32^j
    SDSP_HD static void run(cplx<T> (&a)[32])
    {
        constexpr T h = FftConst<T>::SQRT1_2;
        constexpr T c1 = (T)0.98078528040323044912618223613423903697L, s1 = (T)0.19509032201612826784828486847702224093L;
        constexpr T c2 = FftConst<T>::COS_PI_8, s2 = FftConst<T>::SIN_PI_8;
        constexpr T c3 = (T)0.83146961230254523707878837761790575673L, s3 = (T)0.55557023301960222474283081394853287438L;
        cplx<T> ev[16], od[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 16; j++) {
            ev[j] = a[j] + a[j + 16];
            od[j] = a[j] - a[j + 16];
        }
        // W32^j = (cos(pi j / 16), -sin(pi j / 16)), j = 1 .. 7; W32^(8 + j) = -i W32^j
        od[1] = cmul(od[1], cplx<T>{ c1, -s1 });
        od[2] = cmul(od[2], cplx<T>{ c2, -s2 });
        od[3] = cmul(od[3], cplx<T>{ c3, -s3 });
        od[4] = cplx<T>{ h * (od[4].x + od[4].y), h * (od[4].y - od[4].x) };
        od[5] = cmul(od[5], cplx<T>{ s3, -c3 });
        od[6] = cmul(od[6], cplx<T>{ s2, -c2 });
        od[7] = cmul(od[7], cplx<T>{ s1, -c1 });
        od[8] = mul_neg_i(od[8]);
        od[9] = cmul(od[9], cplx<T>{ -s1, -c1 });
        od[10] = cmul(od[10], cplx<T>{ -s2, -c2 });
        od[11] = cmul(od[11], cplx<T>{ -s3, -c3 });
        od[12] = cplx<T>{ h * (od[12].y - od[12].x), -(h * (od[12].x + od[12].y)) };
        od[13] = cmul(od[13], cplx<T>{ -c3, -s3 });
        od[14] = cmul(od[14], cplx<T>{ -c2, -s2 });
        od[15] = cmul(od[15], cplx<T>{ -c1, -s1 });
        Dft<16, T>::run(ev);
        Dft<16, T>::run(od);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 16; j++) {
            a[2 * j] = ev[j];
            a[2 * j + 1] = od[j];
        }
    }
};

// ------------------------------------------------------------------------------------------------
// one pass over the E points a thread holds.
//   in : v[e] = element at position t + S*e of the current arrangement
//   out: v[q + (E/R)*k] = output k of butterfly b = t + S*q, already multiplied by the pass twiddle
template <class Cfg, int P, typename T>
SDSP_HD void fft_pass(cplx<T> (&v)[Cfg::E], int t, const cplx<T> *__restrict__ tw)
{
    constexpr int R = Cfg::radix(P);
    constexpr int G = Cfg::E / R;              // butterflies per thread
    constexpr int PP = Cfg::pprev(P);          // points already resolved per sub-transform
    constexpr int M = Cfg::N / (PP * R);       // range of m
    constexpr bool LAST = (P + 1 == Cfg::NPASS);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < G; q++) {
        cplx<T> a[R];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < R; j++)
            a[j] = v[q + G * j];
        Dft<R, T>::run(a);
        if (!LAST) {
            const int b = t + Cfg::S * q;
            const int m = b / PP;
            const cplx<T> *row = tw + Cfg::tw_offset(P) + m;
            if (R == 32 && M >= SDSP_FFT_TW4_MIN_M) {
                // radix-32 first pass (fp32 frames of 8192 / 16384 points): the same idea with eight loads (k = 1, 4, 8 .. 28) and
                // 23 products, three roundings deep at most
                const cplx<T> w1 = row[0];
                const cplx<T> w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                const cplx<T> lo[4] = { w1, w1, w2, w3 };
                cplx<T> hi[8];
                hi[0] = w1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = 1; j < 8; j++)
                    hi[j] = row[(4 * j - 1) * M];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int k = 1; k < R; k++) {
                    const cplx<T> w = k < 4 ? lo[k] : (k % 4 == 0 ? hi[k / 4] : cmul(hi[k / 4], lo[k % 4]));
                    a[k] = cmul(a[k], w);
                }
            } else if ((sizeof(T) == 8 || SDSP_FFT_TW4_F32) && R == 16 && M >= SDSP_FFT_TW4_MIN_M) {
                // fp64 frames of 4096 points and more, first pass: the 15 factors W^(m k) of a butterfly come from a table of
                // 15 x M x 16 bytes (61 KB at 4096 points) that does not stay in what the exchange buffers leave of L1, and
                // loading them all costs as many bytes from L2 as the frame itself.  Four of them are loaded (k = 1, 4, 8, 12),
                // the rest are products W^(4j m) W^(m r) with r = k mod 4 -- two roundings deep; eleven more complex products on
                // an FP64 pipe that is a third busy (profiles/r01_ncu_fft4096_f64_v3.txt).
                const cplx<T> w1 = row[0], w4 = row[3 * M], w8 = row[7 * M], w12 = row[11 * M];
                const cplx<T> w2 = cmul(w1, w1), w3 = cmul(w2, w1);
                const cplx<T> lo[4] = { w1, w1, w2, w3 }, hi[4] = { w1, w4, w8, w12 };
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int k = 1; k < R; k++) {
                    const cplx<T> w = k < 4 ? lo[k] : (k % 4 == 0 ? hi[k / 4] : cmul(hi[k / 4], lo[k % 4]));
                    a[k] = cmul(a[k], w);
                }
            } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int k = 1; k < R; k++)
                    a[k] = cmul(a[k], row[(k - 1) * M]);
            }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < R; k++)
            v[q + G * k] = a[k];
    }
}

// where output slot e = q + (E/R)*k of pass P lands in the next arrangement
template <class Cfg, int P>
SDSP_HD int fft_out_pos(int t, int e)
{
    constexpr int R = Cfg::radix(P);
    constexpr int G = Cfg::E / R;
    constexpr int PP = Cfg::pprev(P);
    const int q = e % G, k = e / G;
    const int b = t + Cfg::S * q;
    return (b % PP) + PP * k + PP * R * (b / PP);
}

// Exchange addresses in closed form.  pad(pos) = pos + pos/16; for the factorisations used here the padded
// address splits into a per-thread base plus a compile-time multiple of the slot index, so the compiler
// emits one base computation per pass and immediate offsets for the 16 accesses (the generic pad(pos) form
// costs an LEA/LOP3/SHF chain per access -- 22 % of the instruction stream in the first profile).
template <class Cfg>
SDSP_HD int fft_read_phys(int t, int e)
{
    if (Cfg::S % Cfg::PADW == 0)
        return t + (t >> Cfg::PADSH) + e * (Cfg::S + Cfg::S / Cfg::PADW);
    return Cfg::pad(t + Cfg::S * e);
}
template <class Cfg, int P>
SDSP_HD int fft_out_phys(int t, int e)
{
    constexpr int R = Cfg::radix(P);
    constexpr int G = Cfg::E / R;
    constexpr int PP = Cfg::pprev(P);
    constexpr int W = Cfg::PADW;
    if (G == 1 && PP % W == 0) { // b = t, K = t % PP, m = t / PP, k = e
        const int K = t % PP, m = t / PP;
        return K + (K >> Cfg::PADSH) + (PP * R + PP * R / W) * m + e * (PP + PP / W);
    }
    if (G == 1 && PP == 1 && R == W) // pos = e + R t
        return (R + 1) * t + e;
    if (G > 1 && PP % W == 0 && Cfg::S % PP == 0) { // b = t + S q: K = t % PP, m = t / PP + (S / PP) q, with q = e % G, k = e / G
        const int K = t % PP, m = t / PP;
        const int q = e % G, k = e / G;
        return K + (K >> Cfg::PADSH) + (PP * R + PP * R / W) * m + (PP * R + PP * R / W) * (Cfg::S / PP) * q + k * (PP + PP / W);
    }
    return Cfg::pad(fft_out_pos<Cfg, P>(t, e));
}

// ------------------------------------------------------------------------------------------------
// half spectra of real frames (fft_r2c_kernel / fft_c2r_kernel): a frame of n = 2M reals read as M complex numbers z[j] = x[2j] + i x[2j+1].
//   forward: X[k] = 1/2 [ (Z[k] + conj Z[M-k]) - i W_n^k (Z[k] - conj Z[M-k]) ]                      (zk = Z[k], zm = Z[M-k], w = W_n^k)
//   back:    Z[k] = 1/2 [ (X[k] + conj X[M-k]) + i conj(W_n^k) (X[k] - conj X[M-k]) ]                 (the 1/2 is left to the caller's scale)
template <typename T>
SDSP_HD cplx<T> r2c_bin(cplx<T> zk, cplx<T> zm, cplx<T> w)
{
    const cplx<T> b = cplx<T>{ zm.x, -zm.y };
    const cplx<T> ye = zk + b, d = zk - b;
    const cplx<T> yo = cplx<T>{ d.y, -d.x }; // -i (zk - conj zm)
    const cplx<T> x = ye + cmul(yo, w);
    return cplx<T>{ (T)0.5 * x.x, (T)0.5 * x.y };
}
template <typename T>
SDSP_HD cplx<T> c2r_bin(cplx<T> xk, cplx<T> xm, cplx<T> w)
{
    const cplx<T> b = cplx<T>{ xm.x, -xm.y };
    const cplx<T> s = xk + b, d = xk - b;
    const cplx<T> r = cmul(d, cplx<T>{ w.x, -w.y }); // conj(W_n^k) (xk - conj xm)
    return cplx<T>{ s.x - r.y, s.y + r.x };        // + i r
}

// ------------------------------------------------------------------------------------------------
// host emulation of one frame: every "thread" runs the very pass code the kernel runs, the shared
// memory exchange is an array indexed through the same pad() function.
template <class Cfg, typename T, int P>
inline void fft_emulate_pass(cplx<T> *cur, cplx<T> *nxt, const cplx<T> *tw)
{
    for (int t = 0; t < Cfg::TPF; t++) {
        cplx<T> v[Cfg::E];
        for (int e = 0; e < Cfg::E; e++)
            v[e] = cur[P == 0 ? (t + Cfg::S * e) : fft_read_phys<Cfg>(t, e)];
        fft_pass<Cfg, P, T>(v, t, tw);
        for (int e = 0; e < Cfg::E; e++) {
            if (P + 1 == Cfg::NPASS)
                nxt[t + Cfg::S * e] = v[e]; // natural order: slot e of thread t is output t + S*e
            else
                nxt[fft_out_phys<Cfg, P>(t, e)] = v[e];
        }
    }
    if (P + 1 < Cfg::NPASS)
        fft_emulate_pass<Cfg, T, (P + 1 < Cfg::NPASS ? P + 1 : P)>(nxt, cur, tw);
}

// Returns with the result in `frame` (natural order).  inverse is the re/im swap identity
// IDFT(x) = swap(DFT(swap(x))) / N, which is exact, so only forward butterflies are ever compiled.
template <class Cfg, typename T>
inline void fft_emulate_frame(cplx<T> *frame, const cplx<T> *tw, bool inverse, T scale)
{
    constexpr int BUF = Cfg::PADDED_N > Cfg::N ? Cfg::PADDED_N : Cfg::N;
    cplx<T> *a = new cplx<T>[BUF]();
    cplx<T> *b = new cplx<T>[BUF]();
    for (int i = 0; i < Cfg::N; i++)
        a[i] = inverse ? cplx<T>{ frame[i].y, frame[i].x } : frame[i];
    fft_emulate_pass<Cfg, T, 0>(a, b, tw);
    const cplx<T> *res = (Cfg::NPASS % 2) ? b : a;
    for (int i = 0; i < Cfg::N; i++) {
        cplx<T> r = res[i];
        if (inverse)
            r = cplx<T>{ r.y * scale, r.x * scale };
        frame[i] = r;
    }
    delete[] a;
    delete[] b;
}
} // namespace sdsp_b200
