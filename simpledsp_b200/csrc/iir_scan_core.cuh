// iir_scan_core.cuh -- the algebra of the time-parallel ("scan") IIR path (host + device).
//
// What it replaces: casc_2o_iir<m_t>::process (reference include/sdsp/casc_2o_iir.h:36-80) when there
// are too few channels to fill the GPU with one lane per channel -- in the limit one channel of 2^30
// samples (BASELINE config 4).  The reference loop is strictly serial in time; it is linear, though, so a
// stretch of L samples started from the wrong history can be repaired afterwards:
//
//   state:   the cascade's memory is (u[n-1], u[n-2]) -- the scaled input, known without filtering --
//            plus the 2m values c = (v_j[n-1], v_j[n-2]), j < m, the outputs of the sections.
//   chunk:   run L samples with the true input history but c = 0  ->  y0[i], final state c0.
//            truth:  y[i] = y0[i] + sum_k H[i][k] c_in[k],     c_out = c0 + A_L c_in
//            where column k of H / A_L is the output / final state of the input-free cascade started
//            from the unit state e_k (tables, per channel, computed once on the host in double).
//   carry:   c_in of chunk l+1 is c_out of chunk l: an affine recurrence with one constant matrix, so a
//            Kogge-Stone scan over the 32 lanes of a warp needs only A_L^(2^j), j < 5:
//              P <- P + A_(L 2^j) * shfl_up(P, 2^j)            (zero-input prefix of the tile)
//              Q <- A_(L 2^j) * Q  where bit j of the lane is set   (A_L^lane * state entering the tile)
//   tiles:   a warp owns 32 consecutive chunks; tiles of one channel are chained through small
//            records in global memory (decoupled look-back).  With Mt = A_L^32:
//              c_in(tile t) = sum_{k>=1} Mt^(k-1) agg(t-k)
//            For every stable filter Mt^K falls below 2^-80 (fp64) / 2^-50 (fp32) after a few tiles
//            (K = "reach", found on the host); the sum is then cut after K terms -- orders of magnitude
//            below one ulp -- and tiles do not wait for one another's final result.  Filters whose reach exceeds the window fall back
//            to waiting for the predecessor's inclusive state (correct for any filter, slower).
//
// All matrices are block lower-triangular (section j never feeds back into sections < j); only those
// blocks are multiplied.
#pragma once
#include <vector>

#include "iir_core.cuh"

namespace sdsp_b200
{
constexpr int SCAN_LANES = 32;
constexpr int SCAN_KS_STEPS = 5;   // log2(32)
constexpr int SCAN_MAX_REACH = 8;  // look-back window of the fast path
template <typename T>
constexpr double scan_negligible()
{
    return sizeof(T) == 4 ? 8.8817841970012523e-16 /* 2^-50 */ : 8.2718061255302767e-25 /* 2^-80 */;
}

SDSP_HD constexpr int scan_sd(int m) // dimension of the carried state
{
    return 2 * m;
}
// table of one channel: H[L][sd], A_(L 2^j)[sd][sd] for j < 5, Mt[sd][sd]
SDSP_HD constexpr int scan_table_count(int m, int L)
{
    return L * scan_sd(m) + (SCAN_KS_STEPS + 1) * scan_sd(m) * scan_sd(m);
}
SDSP_HD constexpr int scan_off_A(int m, int L, int j)
{
    return L * scan_sd(m) + j * scan_sd(m) * scan_sd(m);
}

// y += A x for a block lower-triangular A (row-major [SD][SD]); rows 2j, 2j+1 read columns < 2j + 2
template <typename T, int SD>
SDSP_HD void tri_matvec_acc(const T *__restrict__ A, const T (&x)[SD], T (&y)[SD])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < SD; r++) {
        T acc = y[r];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < SD; k++)
            if (k < (r / 2) * 2 + 2)
                acc = fma_t(A[r * SD + k], x[k], acc);
        y[r] = acc;
    }
}

// The carried vector.  fp64: c = (v_j[n-1], v_j[n-2]).  fp32 (delta form, iir_core.cuh): c = (v_j[n-1], d_j[n-1]) with
// d = v[n-1] - v[n-2] the running difference -- in that basis the propagation matrices of a narrow-band section are
// well conditioned (in the (v1, v2) basis A_L ~ [[1+L, -L], [L, 1-L]]: its product with a state rounded to fp32
// cancels L-fold), and it is what the fp32 section carries anyway.  The host tables are built in the same basis.
template <typename T, int M>
SDSP_HD void scan_state_to_vec(const IirState<T, M> &s, T (&c)[2 * M])
{
    SDSP_UNROLL
    for (int j = 0; j < M; j++) {
        c[2 * j] = s.h[j + 1][0];
        c[2 * j + 1] = IirDelta<T>::value ? s.d[j] : s.h[j + 1][1];
    }
}
template <typename T, int M>
SDSP_HD void scan_vec_to_state(const T (&c)[2 * M], IirState<T, M> &s)
{
    SDSP_UNROLL
    for (int j = 0; j < M; j++) {
        s.h[j + 1][0] = c[2 * j];
        if (IirDelta<T>::value) {
            s.d[j] = c[2 * j + 1];
            s.h[j + 1][1] = c[2 * j] - c[2 * j + 1];
        } else {
            s.h[j + 1][1] = c[2 * j + 1];
            s.d[j] = 0;
        }
    }
}

// ---- host: tables ---------------------------------------------------------------------------------
// delta_basis: tables for the carried vector (v1, d) instead of (v1, v2): with S = blockdiag([[1, 0], [1, -1]]) (its own
// inverse; (v1, v2) = S (v1, d)),  H' = H S  and  A' = S A S
template <int M, int KIND>
inline void scan_build_tables_mk(double gain, const double *b, const double *a, int L, double negligible, bool delta_basis,
                                 std::vector<double> &out, int &reach)
{
    constexpr int SD = 2 * M;
    IirCoef<double, M> c;
    iir_pack_coef<double, M>(c, gain, b, a);
    out.assign(scan_table_count(M, L), 0.0);
    double *H = out.data();
    double *A0 = out.data() + scan_off_A(M, L, 0);
    for (int k = 0; k < SD; k++) { // input-free cascade started from the unit state e_k
        IirState<double, M> s;
        for (int r = 0; r <= M; r++)
            s.h[r][0] = s.h[r][1] = 0.0;
        for (int j = 0; j < M; j++)
            s.d[j] = 0.0;
        s.h[1 + k / 2][k % 2] = 1.0;
        for (int i = 0; i < L; i++)
            H[i * SD + k] = iir_step<double, M, KIND>(0.0, c, s);
        for (int j = 0; j < M; j++) {
            A0[(2 * j) * SD + k] = s.h[j + 1][0];
            A0[(2 * j + 1) * SD + k] = s.h[j + 1][1];
        }
    }
    if (delta_basis) {
        for (int i = 0; i < L; i++) // H S: column pair (2j, 2j+1) -> (h0 + h1, -h1)
            for (int j = 0; j < M; j++) {
                const double h0 = H[i * SD + 2 * j], h1 = H[i * SD + 2 * j + 1];
                H[i * SD + 2 * j] = h0 + h1;
                H[i * SD + 2 * j + 1] = -h1;
            }
        for (int r = 0; r < SD; r++) // A S
            for (int j = 0; j < M; j++) {
                const double x0 = A0[r * SD + 2 * j], x1 = A0[r * SD + 2 * j + 1];
                A0[r * SD + 2 * j] = x0 + x1;
                A0[r * SD + 2 * j + 1] = -x1;
            }
        for (int k = 0; k < SD; k++) // S (A S): row pair (2j, 2j+1) -> (r0, r0 - r1)
            for (int j = 0; j < M; j++)
                A0[(2 * j + 1) * SD + k] = A0[(2 * j) * SD + k] - A0[(2 * j + 1) * SD + k];
    }
    auto square = [&](const double *X, double *Y) {
        for (int r = 0; r < SD; r++)
            for (int k = 0; k < SD; k++) {
                long double acc = 0;
                for (int t = 0; t < SD; t++)
                    acc += (long double)X[r * SD + t] * (long double)X[t * SD + k];
                Y[r * SD + k] = (double)acc;
            }
    };
    for (int j = 1; j <= SCAN_KS_STEPS; j++) // A_(2L), A_(4L), ... A_(32L) = Mt
        square(out.data() + scan_off_A(M, L, j - 1), out.data() + scan_off_A(M, L, j));
    // reach: smallest K with max |Mt^K| <= negligible (2^-80 for fp64 data, 2^-50 for fp32: eight and more
    // decimal orders below one ulp of anything the carried state is added to)
    const double *Mt = out.data() + scan_off_A(M, L, SCAN_KS_STEPS);
    std::vector<double> pw(Mt, Mt + SD * SD), nx(SD * SD);
    reach = 1 << 30;
    for (int K = 1; K <= 64; K++) {
        bool zero = true;
        for (double v : pw)
            zero = zero && (v <= negligible && v >= -negligible);
        if (zero) {
            reach = K;
            break;
        }
        for (int r = 0; r < SD; r++)
            for (int k = 0; k < SD; k++) {
                double acc = 0;
                for (int t = 0; t < SD; t++)
                    acc += pw[r * SD + t] * Mt[t * SD + k];
                nx[r * SD + k] = acc;
            }
        pw.swap(nx);
    }
}

inline int scan_build_tables(int m, int kind, double gain, const double *b, const double *a, int L, double negligible, bool delta_basis,
                             std::vector<double> &out, int &reach)
{
#define SDSP_SCAN_CASE(MM)                                                                \
    case MM:                                                                              \
        switch (kind) {                                                                   \
        case NUM_GENERIC: scan_build_tables_mk<MM, NUM_GENERIC>(gain, b, a, L, negligible, delta_basis, out, reach); break; \
        case NUM_LP: scan_build_tables_mk<MM, NUM_LP>(gain, b, a, L, negligible, delta_basis, out, reach); break;  \
        case NUM_HP: scan_build_tables_mk<MM, NUM_HP>(gain, b, a, L, negligible, delta_basis, out, reach); break;  \
        default: scan_build_tables_mk<MM, NUM_BP>(gain, b, a, L, negligible, delta_basis, out, reach); break;      \
        }                                                                                 \
        return 0;
    switch (m) {
        SDSP_SCAN_CASE(2)
        SDSP_SCAN_CASE(4)
        SDSP_SCAN_CASE(6)
        SDSP_SCAN_CASE(8)
    default: return -1;
    }
#undef SDSP_SCAN_CASE
}

// ---- host emulation of one channel: the kernel's algorithm, lanes and tiles played in order ---------
// data: n samples (only whole tiles of 32*L are consumed; returns how many samples were processed).
// s: full history in / out.  force_general: take the wait-for-inclusive path regardless of reach.
template <typename T, int M, int KIND>
inline size_t scan_emulate_channel(const IirCoef<T, M> &c, IirState<T, M> &s, const std::vector<double> &tab64, int reach, int L,
                                   T *data, size_t n, bool force_general)
{
    constexpr int SD = 2 * M;
    std::vector<T> tab(tab64.size());
    for (size_t i = 0; i < tab.size(); i++)
        tab[i] = (T)tab64[i];
    const T *H = tab.data();
    const T *Mt = tab.data() + scan_off_A(M, L, SCAN_KS_STEPS);
    const size_t tile = (size_t)SCAN_LANES * L;
    const size_t n_tiles = n / tile;
    if (n_tiles == 0)
        return 0;
    std::vector<std::vector<T>> agg(n_tiles, std::vector<T>(SD)), incl(n_tiles, std::vector<T>(SD));
    T cin0[SD];
    scan_state_to_vec<T, M>(s, cin0);
    T uh[2] = { s.h[0][0], s.h[0][1] }; // scaled-input history entering the current tile
    const bool fast = !force_general && reach <= SCAN_MAX_REACH;

    for (size_t t = 0; t < n_tiles; t++) {
        T *x = data + t * tile;
        T c0[SCAN_LANES][SD], P[SCAN_LANES][SD], Q[SCAN_LANES][SD];
        T next_uh[2] = { x[tile - 1] * c.gain, x[tile - 2] * c.gain };
        // per-lane halo of scaled inputs, read before anything is overwritten
        T halo[SCAN_LANES][2];
        for (int l = 0; l < SCAN_LANES; l++) {
            if (l == 0) {
                halo[l][0] = uh[0];
                halo[l][1] = uh[1];
            } else {
                halo[l][0] = x[(size_t)l * L - 1] * c.gain;
                halo[l][1] = x[(size_t)l * L - 2] * c.gain;
            }
        }
        // zero-state pass
        for (int l = 0; l < SCAN_LANES; l++) {
            IirState<T, M> z;
            iir_zero_state<T, M>(z);
            z.h[0][0] = halo[l][0];
            z.h[0][1] = halo[l][1];
            T *chunk = x + (size_t)l * L;
            for (int i = 0; i < L; i++)
                chunk[i] = iir_step<T, M, KIND>(chunk[i], c, z);
            scan_state_to_vec<T, M>(z, c0[l]);
            for (int k = 0; k < SD; k++)
                P[l][k] = c0[l][k];
        }
        // Kogge-Stone prefix over lanes
        for (int j = 0; j < SCAN_KS_STEPS; j++) {
            const int d = 1 << j;
            const T *Aj = tab.data() + scan_off_A(M, L, j);
            T nP[SCAN_LANES][SD];
            for (int l = 0; l < SCAN_LANES; l++) {
                for (int k = 0; k < SD; k++)
                    nP[l][k] = P[l][k];
                if (l >= d)
                    tri_matvec_acc<T, SD>(Aj, P[l - d], nP[l]);
            }
            for (int l = 0; l < SCAN_LANES; l++)
                for (int k = 0; k < SD; k++)
                    P[l][k] = nP[l][k];
        }
        for (int k = 0; k < SD; k++)
            agg[t][k] = P[SCAN_LANES - 1][k];
        // state entering the tile
        T cin[SD];
        if (t == 0) {
            for (int k = 0; k < SD; k++)
                cin[k] = cin0[k];
        } else if (fast) {
            const size_t K = (size_t)reach < t ? (size_t)reach : t;
            T acc[SD];
            for (int k = 0; k < SD; k++)
                acc[k] = (K == t) ? cin0[k] : (T)0; // the chain reaches tile 0: its incoming state is the bank's
            for (size_t kk = K; kk >= 1; kk--) {
                T nxt[SD];
                for (int k = 0; k < SD; k++)
                    nxt[k] = agg[t - kk][k];
                tri_matvec_acc<T, SD>(Mt, acc, nxt);
                for (int k = 0; k < SD; k++)
                    acc[k] = nxt[k];
            }
            for (int k = 0; k < SD; k++)
                cin[k] = acc[k];
        } else {
            for (int k = 0; k < SD; k++)
                cin[k] = incl[t - 1][k];
        }
        {
            T v[SD];
            for (int k = 0; k < SD; k++)
                v[k] = agg[t][k];
            tri_matvec_acc<T, SD>(Mt, cin, v);
            for (int k = 0; k < SD; k++)
                incl[t][k] = v[k];
        }
        // A_L^lane * cin by binary decomposition of the lane index
        for (int l = 0; l < SCAN_LANES; l++) {
            for (int k = 0; k < SD; k++)
                Q[l][k] = cin[k];
            for (int j = 0; j < SCAN_KS_STEPS; j++)
                if (l & (1 << j)) {
                    T nq[SD];
                    for (int k = 0; k < SD; k++)
                        nq[k] = 0;
                    tri_matvec_acc<T, SD>(tab.data() + scan_off_A(M, L, j), Q[l], nq);
                    for (int k = 0; k < SD; k++)
                        Q[l][k] = nq[k];
                }
        }
        // correction
        for (int l = 0; l < SCAN_LANES; l++) {
            T ci[SD];
            for (int k = 0; k < SD; k++)
                ci[k] = Q[l][k] + (l ? P[l - 1][k] : (T)0);
            T *chunk = x + (size_t)l * L;
            for (int i = 0; i < L; i++) {
                T y = chunk[i];
                for (int k = 0; k < SD; k++)
                    y = fma_t(H[i * SD + k], ci[k], y);
                chunk[i] = y;
            }
        }
        uh[0] = next_uh[0];
        uh[1] = next_uh[1];
    }
    s.h[0][0] = uh[0];
    s.h[0][1] = uh[1];
    T last[SD];
    for (int k = 0; k < SD; k++)
        last[k] = incl[n_tiles - 1][k];
    scan_vec_to_state<T, M>(last, s);
    return n_tiles * tile;
}
} // namespace sdsp_b200
