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
#pragma once
#include "common.h"

namespace sdsp_b200
{
enum : int { NUM_GENERIC = 0, NUM_LP = 1, NUM_HP = 2, NUM_BP = 3 };

// coefficients of one channel, negated denominators so that every update is a plain fma
template <typename T, int M>
struct IirCoef {
    T gain;
    T b1[M], b2[M];   // only read by NUM_GENERIC
    T na1[M], na2[M]; // -a1, -a2
};

// history of one channel: row 0 = scaled input, row j+1 = output of section j; [.][0] = x[n-1], [.][1] = x[n-2]
template <typename T, int M>
struct IirState {
    T h[M + 1][2];
};

// number of coefficient / state scalars per channel in the device-side structure-of-arrays banks
SDSP_HD constexpr int iir_coef_count(int m)
{
    return 1 + 4 * m;
}
SDSP_HD constexpr int iir_state_count(int m)
{
    return 2 * (m + 1);
}

// one section, one sample:  v = in0 + b1*in1 + b2*in2 - a2*v2 - a1*v1.
// The evaluation order is fixed (SDSP_IIR_ORDER): every kernel (sequential, skewed, packed, scan) and the host
// emulation call this one function, which is what makes their results bit-identical to one another.
// Order 'A' -- fma(na1, v1, fma(na2, v2, fma(b2, in2, fma(b1, in1, in0)))) -- puts v1, this section's previous
// output and the only loop-carried operand that is one sample old, into the LAST operation: the recurrence costs one
// FMA latency per sample and a section four instructions.  The four orders measured on the nine golden fixtures and
// on noise in fp32 are equally accurate (worst fixture 9.6e-5 .. 1.07e-4 of peak), so
// the shortest chain wins; the lane-per-channel kernel with 512 warps is bound by exactly this latency.
#ifndef SDSP_IIR_ORDER
#define SDSP_IIR_ORDER 'A'
#endif
template <int KIND, typename T>
SDSP_HD T iir_numpart(T in1, T in2, T b1, T b2) // b1*in1 + b2*in2
{
    if (KIND == NUM_GENERIC)
        return fma_t(b2, in2, b1 * in1);
    if (KIND == NUM_LP) // {1, 2, 1}
        return fma_t((T)2, in1, in2);
    if (KIND == NUM_HP) // {1, -2, 1}
        return fma_t((T)-2, in1, in2);
    return -in2; // NUM_BP {1, 0, -1}
}
// fp64: four fused multiply-adds in the obvious order (accuracy is not at stake there -- 7e-15 absolute on the
// golden fixtures -- and the FP64 pipe is what bounds the fp64 kernels, so the operation count matters)
template <int KIND>
SDSP_HD double iir_section(double in0, double in1, double in2, double v1, double v2, double b1, double b2, double na1, double na2)
{
    double acc;
    if (KIND == NUM_GENERIC)
        acc = fma_t(b2, in2, fma_t(b1, in1, in0));
    else if (KIND == NUM_LP)
        acc = fma_t(2.0, in1, in0) + in2;
    else if (KIND == NUM_HP)
        acc = fma_t(-2.0, in1, in0) + in2;
    else
        acc = in0 - in2;
    return fma_t(na1, v1, fma_t(na2, v2, acc));
}
// fp32: SDSP_IIR_ORDER
template <int KIND>
SDSP_HD float iir_section(float in0, float in1, float in2, float v1, float v2, float b1, float b2, float na1, float na2)
{
    using T = float;
#if SDSP_IIR_ORDER == 'A' // numerator chain from in0, then both feedback terms
    T acc;
    if (KIND == NUM_GENERIC)
        acc = fma_t(b2, in2, fma_t(b1, in1, in0));
    else if (KIND == NUM_LP)
        acc = fma_t((T)2, in1, in0) + in2;
    else if (KIND == NUM_HP)
        acc = fma_t((T)-2, in1, in0) + in2;
    else
        acc = in0 - in2;
    return fma_t(na1, v1, fma_t(na2, v2, acc));
#elif SDSP_IIR_ORDER == 'B' // everything old first, then (part + in0), then the v1 term
    const T part = fma_t(na2, v2, iir_numpart<KIND, T>(in1, in2, b1, b2));
    return fma_t(na1, v1, part + in0);
#elif SDSP_IIR_ORDER == 'J' // feedback pair combined on its own (the cancelling terms), numerator and in0 added last
    const T fb = fma_t(na1, v1, na2 * v2);
    return (fb + iir_numpart<KIND, T>(in1, in2, b1, b2)) + in0;
#else // 'K'
    const T t = na2 * v2 + (iir_numpart<KIND, T>(in1, in2, b1, b2) + in0);
    return fma_t(na1, v1, t);
#endif
}

// one input sample through the whole cascade; returns the output sample and advances the history
template <typename T, int M, int KIND>
SDSP_HD T iir_step(T x, const IirCoef<T, M> &c, IirState<T, M> &s)
{
    T in0 = x * c.gain;
    T in1 = s.h[0][0], in2 = s.h[0][1];
    s.h[0][1] = in1;
    s.h[0][0] = in0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < M; j++) {
        const T v1 = s.h[j + 1][0], v2 = s.h[j + 1][1];
        const T v = iir_section<KIND>(in0, in1, in2, v1, v2, c.b1[j], c.b2[j], c.na1[j], c.na2[j]);
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
    T inh[M][2], vh[M][2], pipe[M + 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < M; j++) {
        inh[j][0] = s.h[j][0];
        inh[j][1] = s.h[j][1];
        vh[j][0] = s.h[j + 1][0];
        vh[j][1] = s.h[j + 1][1];
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
                const T in0 = (j == 0) ? load(smp) * c.gain : pipe[j];
                const T v = iir_section<KIND>(in0, inh[j][0], inh[j][1], vh[j][0], vh[j][1], c.b1[j], c.b2[j], c.na1[j], c.na2[j]);
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
SDSP_HD f32x2 iir_numpart_x2(f32x2 in1, f32x2 in2, f32x2 b1, f32x2 b2)
{
    if (KIND == NUM_GENERIC)
        return fma2(b2, in2, mul2(b1, in1));
    if (KIND == NUM_LP)
        return fma2(mk2(2.f, 2.f), in1, in2);
    if (KIND == NUM_HP)
        return fma2(mk2(-2.f, -2.f), in1, in2);
    return mk2(-in2.x, -in2.y);
}
template <int KIND>
SDSP_HD f32x2 iir_section_x2(f32x2 in0, f32x2 in1, f32x2 in2, f32x2 v1, f32x2 v2, f32x2 b1, f32x2 b2, f32x2 na1, f32x2 na2)
{
#if SDSP_IIR_ORDER == 'A'
    f32x2 acc;
    if (KIND == NUM_GENERIC)
        acc = fma2(b2, in2, fma2(b1, in1, in0));
    else if (KIND == NUM_LP)
        acc = add2(fma2(mk2(2.f, 2.f), in1, in0), in2);
    else if (KIND == NUM_HP)
        acc = add2(fma2(mk2(-2.f, -2.f), in1, in0), in2);
    else
        acc = add2(in0, mk2(-in2.x, -in2.y));
    return fma2(na1, v1, fma2(na2, v2, acc));
#elif SDSP_IIR_ORDER == 'B'
    const f32x2 part = fma2(na2, v2, iir_numpart_x2<KIND>(in1, in2, b1, b2));
    return fma2(na1, v1, add2(part, in0));
#elif SDSP_IIR_ORDER == 'J'
    const f32x2 fb = fma2(na1, v1, mul2(na2, v2));
    return add2(add2(fb, iir_numpart_x2<KIND>(in1, in2, b1, b2)), in0);
#else
    const f32x2 t = add2(mul2(na2, v2), add2(iir_numpart_x2<KIND>(in1, in2, b1, b2), in0));
    return fma2(na1, v1, t);
#endif
}

template <int M, int KIND, int TS, typename Load, typename Store>
SDSP_HD void iir_tile_skewed_x2(const IirCoef<float, M> &c, IirState<float, M> &s, Load &&load, Store &&store)
{
    static_assert(M % 2 == 0, "pairs of sections");
    constexpr int P = M / 2;
    f32x2 b1[P], b2[P], na1[P], na2[P], inh0[P], inh1[P], vh0[P], vh1[P];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < P; p++) {
        b1[p] = mk2(c.b1[p], c.b1[p + P]);
        b2[p] = mk2(c.b2[p], c.b2[p + P]);
        na1[p] = mk2(c.na1[p], c.na1[p + P]);
        na2[p] = mk2(c.na2[p], c.na2[p + P]);
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
                in0.x = lo_on ? load(slo) * c.gain : 0.f;
                in0.y = from_mid;
            } else {
                in0 = vh0[p - 1];
            }
            if (lo_on && hi_on) {
                const f32x2 v = iir_section_x2<KIND>(in0, inh0[p], inh1[p], vh0[p], vh1[p], b1[p], b2[p], na1[p], na2[p]);
                inh1[p] = inh0[p];
                inh0[p] = in0;
                vh1[p] = vh0[p];
                vh0[p] = v;
                if (p == P - 1)
                    store(shi, v.y);
            } else if (lo_on) { // ramp-up: only the low section of the pair has a sample
                const float v = iir_section<KIND>(in0.x, inh0[p].x, inh1[p].x, vh0[p].x, vh1[p].x, b1[p].x, b2[p].x, na1[p].x, na2[p].x);
                inh1[p].x = inh0[p].x;
                inh0[p].x = in0.x;
                vh1[p].x = vh0[p].x;
                vh0[p].x = v;
            } else { // ramp-down: only the high section still has samples
                const float v = iir_section<KIND>(in0.y, inh0[p].y, inh1[p].y, vh0[p].y, vh1[p].y, b1[p].y, b2[p].y, na1[p].y, na2[p].y);
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
        c.na1[j] = (T)(-a[3 * j + 1]);
        c.na2[j] = (T)(-a[3 * j + 2]);
    }
}
} // namespace sdsp_b200
