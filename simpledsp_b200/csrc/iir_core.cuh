// iir_core.cuh -- per-sample arithmetic of the cascaded second-order-section IIR (host + device).
//
// What it replaces: the inner loop of sdsp::casc_2o_iir<m_t>::process (reference
// include/sdsp/casc_2o_iir.h:51-77) and of the fixed-numerator classes
// (casc_2o_iir_base::process_base :228-263 with process_spec :286-295 lp, :344-353 hp, :402-411 bp).
//
// Recurrence per section j (b0 == 1 implicitly, a0 never read -- casc_2o_iir.h:64-69):
//     v_j[n] = u_j[n] + b1*u_j[n-1] + b2*u_j[n-2] - a1*v_j[n-1] - a2*v_j[n-2],   u_0 = gain*x,  u_{j+1} = v_j
// evaluated as a chain of fused multiply-adds in a FIXED order, so the result of a stream does not
// depend on how it is cut into process() calls (reference test/testIIR.cpp:61-75 demands exact
// equality of block-wise and whole-buffer runs).  The history is kept as the two most recent values of
// every row of the reference's m_mem; its 3-slot ring and m_pos (casc_2o_iir.h:11,15,54-60,73-75) are
// a private representation, replaced here by register rotation.
//
// fp64 evaluates that line as it stands.  fp32 evaluates the SAME recurrence in difference ("delta") form:
//     d_j[n] = g*d_j[n-1] + (numerator - c0*v_j[n-2]),      v_j[n] = v_j[n-1] + d_j[n]
//     with  c0 = 1 + a1 + a2   and   g = a2 - c0 = -(1 + a1)       (both formed in double on the host, rounded once)
// which is algebraically identical (substitute d = v[n] - v[n-1]) but keeps what fp32 cannot afford to lose: for a
// narrow-band section a1 -> -2, a2 -> 1 and the quantity that places the poles is c0 ~ (2 pi f0/fs)^2, i.e. the last
// bits of a1 and a2.  Rounding a1, a2 to fp32 moves c0 by up to 2^-24 ABSOLUTE (10 % of c0 at f0/fs = 2.5e-4, and
// 1e-4 of the impulse response already at f0/fs = 0.005: reference test_data/impulse_response/LPimpulse.csv, SURVEY
// H3); rounding c0 itself moves it by 2^-24 RELATIVE, and a rounding error of g multiplies (1 - z^-1), which
// vanishes where such a filter has its gain.  The running difference d is carried as state next to v, so rounding
// errors of v never re-enter the resonant part of the recurrence.  Measured against the fp64 reference on the nine
// golden fixtures and on noise: direct form 1e-4 (f0/fs = 0.005) ... 1e-3 (f0/fs = 0.001), delta form <= 3e-6 for
// every low-pass / band-pass case and <= 1.3e-5 for the high-pass ones (profiles/r02_iir_f32_delta_form.txt).
#pragma once
#include "common.h"

namespace sdsp_b200
{
enum : int { NUM_GENERIC = 0, NUM_LP = 1, NUM_HP = 2, NUM_BP = 3,
              // internal: a generic bank whose every b2 is exactly 1 (all Butterworth low-/high-pass designs, in any mix).  fma(1, in2, t)
              // IS t + in2, so the result has the same bits; an add has two register operands and issues at full rate where a
              // three-register FFMA takes 1.83 cycles of a warp that is alone on its sub-partition (profiles/r02_ubench_fp32_issue.txt)
              NUM_GENERIC_B2ONE = 4 };

// does precision T run the difference form (and carry d next to the history)?
template <typename T>
struct IirDelta {
    static constexpr bool value = sizeof(T) == 4;
};

// coefficients of one channel, feedback pair pre-arranged so that every update is a plain fma:
//   fp64 (direct form):  fa = -a1,          fb = -a2
//   fp32 (delta form):   fa = g = -(1+a1),  fb = -c0 = -(1+a1+a2)
template <typename T, int M>
struct IirCoef {
    T gain;
    T b1[M], b2[M]; // only read by NUM_GENERIC
    T fa[M], fb[M];
};

// history of one channel: row 0 = scaled input, row j+1 = output of section j; [.][0] = x[n-1], [.][1] = x[n-2];
// d[j] = the running difference of section j's output (delta form only; never read in fp64)
template <typename T, int M>
struct IirState {
    T h[M + 1][2];
    T d[M];
};

// number of coefficient / state scalars per channel in the device-side structure-of-arrays banks.
// State rows: 2r, 2r+1 = history row r (as the C ABI's mem[r][0..1]); delta form: row 2(m+1)+j = d[j].
SDSP_HD constexpr int iir_coef_count(int m)
{
    return 1 + 4 * m;
}
SDSP_HD constexpr int iir_hist_count(int m) // the part of the state the C ABI exposes (reference m_mem)
{
    return 2 * (m + 1);
}
SDSP_HD constexpr int iir_state_count(int m, bool delta)
{
    return 2 * (m + 1) + (delta ? m : 0);
}
// the feedback pair of one section from the reference's a1, a2 (host, double)
inline void iir_feedback_pair(bool delta, double a1, double a2, double &fa, double &fb)
{
    if (delta) {
        fa = -(1.0 + a1);
        fb = -(1.0 + a1 + a2);
    } else {
        fa = -a1;
        fb = -a2;
    }
}

// ---- bank rows <-> registers (structure-of-arrays banks: row k of channel ch at [k * pitch + ch]) ------------------
#if defined(__CUDA_ARCH__)
#define SDSP_UNROLL _Pragma("unroll")
#else
#define SDSP_UNROLL
#endif
template <typename T, int M>
SDSP_HD void iir_load_coef(IirCoef<T, M> &c, const T *__restrict__ coef, size_t pitch, size_t ch)
{
    c.gain = coef[ch];
    SDSP_UNROLL
    for (int j = 0; j < M; j++) {
        c.b1[j] = coef[(size_t)(1 + j) * pitch + ch];
        c.b2[j] = coef[(size_t)(1 + M + j) * pitch + ch];
        c.fa[j] = coef[(size_t)(1 + 2 * M + j) * pitch + ch];
        c.fb[j] = coef[(size_t)(1 + 3 * M + j) * pitch + ch];
    }
}
template <typename T, int M>
SDSP_HD void iir_zero_coef(IirCoef<T, M> &c)
{
    c.gain = 0;
    SDSP_UNROLL
    for (int j = 0; j < M; j++)
        c.b1[j] = c.b2[j] = c.fa[j] = c.fb[j] = 0;
}
template <typename T, int M>
SDSP_HD void iir_load_state(IirState<T, M> &s, const T *__restrict__ state, size_t pitch, size_t ch)
{
    SDSP_UNROLL
    for (int r = 0; r <= M; r++) {
        s.h[r][0] = state[(size_t)(2 * r) * pitch + ch];
        s.h[r][1] = state[(size_t)(2 * r + 1) * pitch + ch];
    }
    SDSP_UNROLL
    for (int j = 0; j < M; j++)
        s.d[j] = IirDelta<T>::value ? state[(size_t)(2 * (M + 1) + j) * pitch + ch] : (T)0;
}
template <typename T, int M>
SDSP_HD void iir_zero_state(IirState<T, M> &s)
{
    SDSP_UNROLL
    for (int r = 0; r <= M; r++)
        s.h[r][0] = s.h[r][1] = 0;
    SDSP_UNROLL
    for (int j = 0; j < M; j++)
        s.d[j] = 0;
}
template <typename T, int M>
SDSP_HD void iir_store_state(const IirState<T, M> &s, T *__restrict__ state, size_t pitch, size_t ch)
{
    SDSP_UNROLL
    for (int r = 0; r <= M; r++) {
        state[(size_t)(2 * r) * pitch + ch] = s.h[r][0];
        state[(size_t)(2 * r + 1) * pitch + ch] = s.h[r][1];
    }
    if (IirDelta<T>::value) {
        SDSP_UNROLL
        for (int j = 0; j < M; j++)
            state[(size_t)(2 * (M + 1) + j) * pitch + ch] = s.d[j];
    }
}
// the C ABI's history mem[r][0..1] (double) <-> a register state; the running differences are not part of that
// interface (the reference's m_mem has no such thing): a history that comes in from outside starts them at
// v[n-1] - v[n-2], which is what they are up to the rounding of one addition
template <typename T, int M>
inline void iir_state_from_mem(IirState<T, M> &s, const double *mem)
{
    for (int r = 0; r <= M; r++) {
        s.h[r][0] = (T)mem[2 * r];
        s.h[r][1] = (T)mem[2 * r + 1];
    }
    for (int j = 0; j < M; j++)
        s.d[j] = IirDelta<T>::value ? s.h[j + 1][0] - s.h[j + 1][1] : (T)0;
}
template <typename T, int M>
inline void iir_state_to_mem(const IirState<T, M> &s, double *mem)
{
    for (int r = 0; r <= M; r++) {
        mem[2 * r] = (double)s.h[r][0];
        mem[2 * r + 1] = (double)s.h[r][1];
    }
}

// one section, one sample.  Every kernel (sequential, skewed, packed, scan) and the host emulation call this one
// function with a fixed evaluation order, which is what makes their results bit-identical to one another.
//   numerator   acc = in0 + b1*in1 + b2*in2        (fixed kinds: {1,2,1}, {1,-2,1}, {1,0,-1} without multiplies)
//   fp64        v = fma(fa, v1, fma(fb, v2, acc))                       fa = -a1, fb = -a2
//               v1, the only loop-carried operand that is one sample old, enters the LAST operation: the recurrence
//               costs one FMA latency per sample.
//   fp32        d = fma(fa, d, fma(fb, v2, acc));  v = v1 + d           fa = g,   fb = -c0   (see the header comment)
//               loop-carried: d -> d one FMA, v -> v one add, and d -> v -> (two samples later) d three operations.
template <int KIND, typename T>
SDSP_HD T iir_numerator(T in0, T in1, T in2, T b1, T b2)
{
    if (KIND == NUM_GENERIC)
        return fma_t(b2, in2, fma_t(b1, in1, in0));
    if (KIND == NUM_GENERIC_B2ONE)
        return add_t(fma_t(b1, in1, in0), in2);
    if (KIND == NUM_LP)
        return fma_t((T)2, in1, in0) + in2;
    if (KIND == NUM_HP)
        return fma_t((T)-2, in1, in0) + in2;
    return in0 - in2; // NUM_BP
}
template <int KIND>
SDSP_HD double iir_section(double in0, double in1, double in2, double v1, double v2, double &d, double b1, double b2, double fa, double fb)
{
    (void)d;
    return fma_t(fa, v1, fma_t(fb, v2, iir_numerator<KIND, double>(in0, in1, in2, b1, b2)));
}
template <int KIND>
SDSP_HD float iir_section(float in0, float in1, float in2, float v1, float v2, float &d, float b1, float b2, float fa, float fb)
{
    d = fma_t(fa, d, fma_t(fb, v2, iir_numerator<KIND, float>(in0, in1, in2, b1, b2)));
    return add_t(v1, d);
}

// one input sample through the whole cascade; returns the output sample and advances the history
template <typename T, int M, int KIND>
SDSP_HD T iir_step(T x, const IirCoef<T, M> &c, IirState<T, M> &s)
{
    T in0 = mul_t(x, c.gain);
    T in1 = s.h[0][0], in2 = s.h[0][1];
    s.h[0][1] = in1;
    s.h[0][0] = in0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < M; j++) {
        const T v1 = s.h[j + 1][0], v2 = s.h[j + 1][1];
        const T v = iir_section<KIND>(in0, in1, in2, v1, v2, s.d[j], c.b1[j], c.b2[j], c.fa[j], c.fb[j]);
        s.h[j + 1][1] = v1;
        s.h[j + 1][0] = v;
        in0 = v;
        in1 = v1;
        in2 = v2;
    }
    return in0;
}

// TS consecutive samples with the sections software-skewed: in iteration i section j works on sample
// i - j, so the M section updates of one iteration are independent of one another and a single warp
// keeps M fma chains in flight instead of one (the loop-carried dependency of the reference's
// sample-by-sample order, casc_2o_iir.h:51-77, is what bounds a lane-per-channel kernel).  Each
// (section, sample) update is the same iir_section() call on the same operands as in iir_step(), so the
// results are bit-identical to the plain loop.  `load(i)` returns input sample i, `store(i, y)` takes
// output sample i; both are called with compile-time-constant i after unrolling.
template <typename T, int M, int KIND, int TS, typename Load, typename Store>
SDSP_HD void iir_tile_skewed(const IirCoef<T, M> &c, IirState<T, M> &s, Load &&load, Store &&store)
{
    T inh[M][2], vh[M][2], dh[M], pipe[M + 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < M; j++) {
        inh[j][0] = s.h[j][0];
        inh[j][1] = s.h[j][1];
        vh[j][0] = s.h[j + 1][0];
        vh[j][1] = s.h[j + 1][1];
        dh[j] = s.d[j];
        pipe[j] = 0;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < TS + M - 1; i++) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = M - 1; j >= 0; j--) { // downstream first: pipe[j] still holds last iteration's value
            const int smp = i - j;
            if (smp >= 0 && smp < TS) {
                const T in0 = (j == 0) ? mul_t(load(smp), c.gain) : pipe[j];
                const T v = iir_section<KIND>(in0, inh[j][0], inh[j][1], vh[j][0], vh[j][1], dh[j], c.b1[j], c.b2[j], c.fa[j], c.fb[j]);
                inh[j][1] = inh[j][0];
                inh[j][0] = in0;
                vh[j][1] = vh[j][0];
                vh[j][0] = v;
                if (j == M - 1)
                    store(smp, v);
                else
                    pipe[j + 1] = v;
            }
        }
    }
    s.h[0][0] = inh[0][0];
    s.h[0][1] = inh[0][1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < M; j++) {
        s.h[j + 1][0] = vh[j][0];
        s.h[j + 1][1] = vh[j][1];
        s.d[j] = dh[j];
    }
}

// ---- two-wide fp32 helpers: one FFMA2 / FADD2 on sm_100, two scalar ops on the host --------------
#if defined(__CUDACC__)
typedef float2 f32x2;
#else
struct f32x2 {
    float x, y;
};
#endif
SDSP_HD f32x2 mk2(float x, float y)
{
    f32x2 r;
    r.x = x;
    r.y = y;
    return r;
}
SDSP_HD f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return mk2(__builtin_fmaf(a.x, b.x, c.x), __builtin_fmaf(a.y, b.y, c.y));
#endif
}
SDSP_HD f32x2 mul2(f32x2 a, f32x2 b)
{
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return mk2(a.x * b.x, a.y * b.y);
#endif
}
SDSP_HD f32x2 add2(f32x2 a, f32x2 b)
{
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return mk2(a.x + b.x, a.y + b.y);
#endif
}

// ---- fp32 tiles with the sections packed two per instruction (fma.rn.f32x2, new on sm_100) ----------
// Pair p holds sections (p, p + M/2) in the two halves of a 64-bit register.  With the skew of
// iir_tile_skewed (section j works on sample i - j) the input of pair p >= 1 is exactly the previous
// output of pair p - 1, both halves at once, so handing values down the cascade costs no instruction;
// only pair 0 needs one move (its high half is the previous output of section M/2 - 1).  Every lane
// still owns one channel and every (section, sample) update is the same fused-multiply-add chain as
// iir_section(), so the bits equal the scalar path; the instruction count per sample drops from
// 1 + 4M to about 2 + 2M.
template <int KIND>
SDSP_HD f32x2 iir_numerator_x2(f32x2 in0, f32x2 in1, f32x2 in2, f32x2 b1, f32x2 b2)
{
    if (KIND == NUM_GENERIC)
        return fma2(b2, in2, fma2(b1, in1, in0));
    if (KIND == NUM_GENERIC_B2ONE)
        return add2(fma2(b1, in1, in0), in2);
    if (KIND == NUM_LP)
        return add2(fma2(mk2(2.f, 2.f), in1, in0), in2);
    if (KIND == NUM_HP)
        return add2(fma2(mk2(-2.f, -2.f), in1, in0), in2);
    return add2(in0, mk2(-in2.x, -in2.y));
}
// the delta-form update of iir_section(float ...), both halves at once
template <int KIND>
SDSP_HD f32x2 iir_section_x2(f32x2 in0, f32x2 in1, f32x2 in2, f32x2 v1, f32x2 v2, f32x2 &d, f32x2 b1, f32x2 b2, f32x2 fa, f32x2 fb)
{
    d = fma2(fa, d, fma2(fb, v2, iir_numerator_x2<KIND>(in0, in1, in2, b1, b2)));
    return add2(v1, d);
}

template <int M, int KIND, int TS, typename Load, typename Store>
SDSP_HD void iir_tile_skewed_x2(const IirCoef<float, M> &c, IirState<float, M> &s, Load &&load, Store &&store)
{
    static_assert(M % 2 == 0, "pairs of sections");
    constexpr int P = M / 2;
    f32x2 b1[P], b2[P], fa[P], fb[P], inh0[P], inh1[P], vh0[P], vh1[P], dh[P];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < P; p++) {
        b1[p] = mk2(c.b1[p], c.b1[p + P]);
        b2[p] = mk2(c.b2[p], c.b2[p + P]);
        fa[p] = mk2(c.fa[p], c.fa[p + P]);
        fb[p] = mk2(c.fb[p], c.fb[p + P]);
        dh[p] = mk2(s.d[p], s.d[p + P]);
        inh0[p] = mk2(s.h[p][0], s.h[p + P][0]);
        inh1[p] = mk2(s.h[p][1], s.h[p + P][1]);
        vh0[p] = mk2(s.h[p + 1][0], s.h[p + P + 1][0]);
        vh1[p] = mk2(s.h[p + 1][1], s.h[p + P + 1][1]);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < TS + M - 1; i++) {
        // what section P-1 produced in the previous iteration feeds section P (high half of pair 0) now;
        // pair P-1 is updated first below, so take the value before that happens
        const float from_mid = vh0[P - 1].x;
        // descending: pair p reads the not-yet-updated output of pair p - 1 (= last iteration's value)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int p = P - 1; p >= 0; p--) {
            const int slo = i - p, shi = i - p - P; // samples the two halves work on
            const bool lo_on = slo >= 0 && slo < TS, hi_on = shi >= 0 && shi < TS;
            if (!lo_on && !hi_on)
                continue;
            f32x2 in0;
            if (p == 0) {
                in0.x = lo_on ? mul_t(load(slo), c.gain) : 0.f;
                in0.y = from_mid;
            } else {
                in0 = vh0[p - 1];
            }
            if (lo_on && hi_on) {
                const f32x2 v = iir_section_x2<KIND>(in0, inh0[p], inh1[p], vh0[p], vh1[p], dh[p], b1[p], b2[p], fa[p], fb[p]);
                inh1[p] = inh0[p];
                inh0[p] = in0;
                vh1[p] = vh0[p];
                vh0[p] = v;
                if (p == P - 1)
                    store(shi, v.y);
            } else if (lo_on) { // ramp-up: only the low section of the pair has a sample
                const float v = iir_section<KIND>(in0.x, inh0[p].x, inh1[p].x, vh0[p].x, vh1[p].x, dh[p].x, b1[p].x, b2[p].x, fa[p].x, fb[p].x);
                inh1[p].x = inh0[p].x;
                inh0[p].x = in0.x;
                vh1[p].x = vh0[p].x;
                vh0[p].x = v;
            } else { // ramp-down: only the high section still has samples
                const float v = iir_section<KIND>(in0.y, inh0[p].y, inh1[p].y, vh0[p].y, vh1[p].y, dh[p].y, b1[p].y, b2[p].y, fa[p].y, fb[p].y);
                inh1[p].y = inh0[p].y;
                inh0[p].y = in0.y;
                vh1[p].y = vh0[p].y;
                vh0[p].y = v;
                if (p == P - 1)
                    store(shi, v);
            }
        }
    }
    // all sections are aligned on the same sample again: row 0 is the input history of section 0, row j + 1 the
    // output history of section j
    s.h[0][0] = inh0[0].x;
    s.h[0][1] = inh1[0].x;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < P; p++) {
        s.h[p + 1][0] = vh0[p].x;
        s.h[p + 1][1] = vh1[p].x;
        s.h[p + P + 1][0] = vh0[p].y;
        s.h[p + P + 1][1] = vh1[p].y;
        s.d[p] = dh[p].x;
        s.d[p + P] = dh[p].y;
    }
}

// PACK (fp32 only): two sections per instruction with fma.rn.f32x2.  Same operations, same bits either way; which is
// faster depends on what bounds the kernel -- packed halves the instruction count (wins when many warps share an SM
// and issue slots are the limit), scalar has twice the independent instructions in flight per warp (wins when a warp
// is alone on its scheduler and latency is the limit).  profiles/r01_iir_pack_vs_scalar.txt
template <typename T, int M, int KIND, int TS, bool PACK = true, typename Load, typename Store>
SDSP_HD void iir_tile_dispatch(const IirCoef<T, M> &c, IirState<T, M> &s, Load &&load, Store &&store)
{
#ifndef SDSP_IIR_NOPACK
    if constexpr (sizeof(T) == 4 && M % 2 == 0 && PACK)
        iir_tile_skewed_x2<M, KIND, TS>(c, s, load, store);
    else
#endif
        iir_tile_skewed<T, M, KIND, TS>(c, s, load, store);
}

// host-side packing helpers shared by the bank upload and the emulator
template <typename T, int M>
inline void iir_pack_coef(IirCoef<T, M> &c, double gain, const double *b /*[M][3] or null*/, const double *a /*[M][3]*/)
{
    c.gain = (T)gain;
    for (int j = 0; j < M; j++) {
        c.b1[j] = b ? (T)b[3 * j + 1] : (T)0;
        c.b2[j] = b ? (T)b[3 * j + 2] : (T)0;
        double fa, fb;
        iir_feedback_pair(IirDelta<T>::value, a[3 * j + 1], a[3 * j + 2], fa, fb);
        c.fa[j] = (T)fa;
        c.fb[j] = (T)fb;
    }
}
} // namespace sdsp_b200
