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

// feed-forward part of one section for the four numerator classes
template <int KIND, typename T>
SDSP_HD T iir_numerator(T in0, T in1, T in2, T b1, T b2)
{
    if (KIND == NUM_GENERIC)
        return fma_t(b2, in2, fma_t(b1, in1, in0));
    if (KIND == NUM_LP) // {1, 2, 1}
        return fma_t((T)2, in1, in0) + in2;
    if (KIND == NUM_HP) // {1, -2, 1}
        return fma_t((T)-2, in1, in0) + in2;
    return in0 - in2; // NUM_BP {1, 0, -1}
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
        T acc = iir_numerator<KIND, T>(in0, in1, in2, c.b1[j], c.b2[j]);
        acc = fma_t(c.na2[j], v2, acc);
        const T v = fma_t(c.na1[j], v1, acc);
        s.h[j + 1][1] = v1;
        s.h[j + 1][0] = v;
        in0 = v;
        in1 = v1;
        in2 = v2;
    }
    return in0;
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
