// iir.cu -- cascaded second-order-section IIR banks for sm_100a: kernels, banks, C-ABI entry points.
//
// Replaces sdsp::casc_2o_iir<m_t> and casc_2o_iir_lp/_hp/_bp<m_t> (reference
// include/sdsp/casc_2o_iir.h:8-468) for banks of independent channels resident in HBM.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "iir_core.cuh"
#include "iir_internal.h"

namespace sdsp_b200
{
// =================================================================================================
// sequential path, generic addressing: lane per channel, a warp owns 32 channels and walks the time
// axis in tiles of 32 samples.  The tile is transposed through a warp-private shared-memory patch so
// that global traffic is row-contiguous (128 B per channel row for fp32) while every lane consumes its
// own channel in order.  The next tile is fetched into registers while the current one is filtered.
template <typename T, int M>
__device__ __forceinline__ void iir_load_channel(IirCoef<T, M> &c, IirState<T, M> &s, const T *__restrict__ coef,
                                                 const T *__restrict__ state, size_t n_channels, size_t ch, bool active)
{
    if (active) {
        iir_load_coef<T, M>(c, coef, n_channels, ch);
        iir_load_state<T, M>(s, state, n_channels, ch);
    } else {
        iir_zero_coef<T, M>(c);
        iir_zero_state<T, M>(s);
    }
}

template <typename T, int M, int KIND, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
    iir_seq_kernel(T *__restrict__ data, size_t n_samples, size_t stride, const T *__restrict__ coef, T *__restrict__ state,
                   size_t n_channels)
{
    constexpr int TS = 32;
    __shared__ T tile[WARPS][32][TS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t ch0 = ((size_t)blockIdx.x * WARPS + warp) * 32;
    if (ch0 >= n_channels)
        return; // warps are independent: nothing below synchronises across the block
    const size_t ch = ch0 + lane;
    const bool active = ch < n_channels;
    const int rows = (int)((n_channels - ch0) < 32 ? (n_channels - ch0) : 32);

    IirCoef<T, M> c;
    IirState<T, M> s;
    iir_load_channel<T, M>(c, s, coef, state, n_channels, ch, active);

    T(*my)[TS + 1] = tile[warp];
    T *base = data + ch0 * stride + lane;
    T nxt[32];
    // prefetch tile 0
#pragma unroll
    for (int r = 0; r < 32; r++)
        nxt[r] = (r < rows && (size_t)lane < n_samples) ? base[(size_t)r * stride] : (T)0;

    for (size_t n0 = 0; n0 < n_samples; n0 += TS) {
#pragma unroll
        for (int r = 0; r < 32; r++)
            my[r][lane] = nxt[r];
        __syncwarp();
        const size_t n1 = n0 + TS;
        if (n1 < n_samples) {
            const bool col_ok = n1 + lane < n_samples;
#pragma unroll
            for (int r = 0; r < 32; r++)
                nxt[r] = (r < rows && col_ok) ? base[(size_t)r * stride + n1] : (T)0;
        }
        const int cnt = (int)((n_samples - n0) < TS ? (n_samples - n0) : TS);
        if (cnt == TS) {
#pragma unroll 8
            for (int i = 0; i < TS; i++)
                my[lane][i] = iir_step<T, M, KIND>(my[lane][i], c, s);
        } else {
            for (int i = 0; i < cnt; i++)
                my[lane][i] = iir_step<T, M, KIND>(my[lane][i], c, s);
        }
        __syncwarp();
        const bool col_ok = n0 + lane < n_samples;
#pragma unroll
        for (int r = 0; r < 32; r++)
            if (r < rows && col_ok)
                base[(size_t)r * stride + n0] = my[r][lane];
        __syncwarp();
    }
    if (active)
        iir_store_state<T, M>(s, state, n_channels, ch);
}

// =================================================================================================
template <typename T, int M, int KIND>
static int launch_seq(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    constexpr int WARPS = 2;
    const size_t groups = (b.n_channels + 31) / 32;
    const unsigned grid = (unsigned)((groups + WARPS - 1) / WARPS);
    iir_seq_kernel<T, M, KIND, WARPS><<<grid, WARPS * 32, 0, stream>>>(static_cast<T *>(data), n_samples, stride,
                                                                      static_cast<const T *>(b.d_coef), static_cast<T *>(b.d_state),
                                                                      b.n_channels);
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

template <typename T, int M>
static int launch_seq_kind(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    switch (b.numerator) {
    case NUM_GENERIC: return launch_seq<T, M, NUM_GENERIC>(b, data, n_samples, stride, stream);
    case NUM_LP: return launch_seq<T, M, NUM_LP>(b, data, n_samples, stride, stream);
    case NUM_HP: return launch_seq<T, M, NUM_HP>(b, data, n_samples, stride, stream);
    default: return launch_seq<T, M, NUM_BP>(b, data, n_samples, stride, stream);
    }
}

template <typename T>
static int launch_seq_sections(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    switch (b.sections) {
    case 1: return launch_seq_kind<T, 1>(b, data, n_samples, stride, stream);
    case 2: return launch_seq_kind<T, 2>(b, data, n_samples, stride, stream);
    case 3: return launch_seq_kind<T, 3>(b, data, n_samples, stride, stream);
    case 4: return launch_seq_kind<T, 4>(b, data, n_samples, stride, stream);
    case 5: return launch_seq_kind<T, 5>(b, data, n_samples, stride, stream);
    case 6: return launch_seq_kind<T, 6>(b, data, n_samples, stride, stream);
    case 7: return launch_seq_kind<T, 7>(b, data, n_samples, stride, stream);
    case 8: return launch_seq_kind<T, 8>(b, data, n_samples, stride, stream);
    default: return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir: sections=%d not built (1..8)", b.sections);
    }
}

// =================================================================================================
// One channel, a few thousand samples: the single filter object of the drop-in header (BASELINE config 1: casc_2o_iir<4> on a
// 4096-sample buffer, reference test/testIIR.cpp:465-496).  Nothing is parallel but the SECTIONS: warp j runs section j, one lane
// each, over tiles of CH_TILE samples held in shared memory; in step s warp j filters tile s - j in place (the output of section
// j - 1 written one step earlier), a CTA barrier between steps.  One warp doing all sections issues 4 x 5 fp64 operations per
// sample on one FP64 pipe (~90 us for 4096 samples); m warps on m sub-partitions are bound by the recurrence's one-FMA latency
// instead (~20 us).  Every update is the same iir_section() call on the same operands as iir_step(): same bits.
constexpr int CH_TILE = 32;
template <typename T, int KIND>
__global__ void __launch_bounds__(256)
    iir_chain_kernel(T *__restrict__ data, int n_samples, const T *__restrict__ coef, T *__restrict__ state, int m)
{
    extern __shared__ __align__(16) unsigned char chain_smem[];
    T *sx = reinterpret_cast<T *>(chain_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < n_samples; i += blockDim.x)
        sx[i] = data[i];
    // section `warp` of the one channel (bank rows of a one-channel bank are contiguous: pitch 1)
    T gain = 1, b1 = 0, b2 = 0, fa = 0, fb = 0, in1 = 0, in2 = 0, v1 = 0, v2 = 0, d = 0;
    if (lane == 0 && warp < m) {
        gain = warp == 0 ? coef[0] : (T)1; // x * 1 is exact: the stream passes from section to section unchanged
        b1 = coef[1 + warp];
        b2 = coef[1 + m + warp];
        fa = coef[1 + 2 * m + warp];
        fb = coef[1 + 3 * m + warp];
        in1 = state[2 * warp];
        in2 = state[2 * warp + 1];
        v1 = state[2 * (warp + 1)];
        v2 = state[2 * (warp + 1) + 1];
        d = IirDelta<T>::value ? state[2 * (m + 1) + warp] : (T)0;
    }
    __syncthreads();
    const int n_tiles = (n_samples + CH_TILE - 1) / CH_TILE;
    for (int s = 0; s < n_tiles + m - 1; s++) {
        const int k = s - warp;
        if (lane == 0 && warp < m && k >= 0 && k < n_tiles) {
            T *p = sx + k * CH_TILE;
            const int cnt = n_samples - k * CH_TILE < CH_TILE ? n_samples - k * CH_TILE : CH_TILE;
            if (cnt == CH_TILE) {
#pragma unroll
                for (int i = 0; i < CH_TILE; i++) {
                    const T in0 = mul_t(p[i], gain);
                    const T v = iir_section<KIND>(in0, in1, in2, v1, v2, d, b1, b2, fa, fb);
                    in2 = in1, in1 = in0, v2 = v1, v1 = v;
                    p[i] = v;
                }
            } else {
                for (int i = 0; i < cnt; i++) {
                    const T in0 = mul_t(p[i], gain);
                    const T v = iir_section<KIND>(in0, in1, in2, v1, v2, d, b1, b2, fa, fb);
                    in2 = in1, in1 = in0, v2 = v1, v1 = v;
                    p[i] = v;
                }
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n_samples; i += blockDim.x)
        data[i] = sx[i];
    if (lane == 0 && warp < m) {
        if (warp == 0) {
            state[0] = in1; // row 0: the scaled input
            state[1] = in2;
        }
        state[2 * (warp + 1)] = v1;
        state[2 * (warp + 1) + 1] = v2;
        if (IirDelta<T>::value)
            state[2 * (m + 1) + warp] = d;
    }
}

template <typename T>
static int launch_chain(const IirBank &b, void *data, size_t n_samples, cudaStream_t stream)
{
    const size_t smem = n_samples * sizeof(T);
    const int threads = 32 * (b.sections < 4 ? 4 : b.sections); // at least four warps for the copies
#define SDSP_CHAIN(KK)                                                                                                        \
    {                                                                                                                         \
        auto kern = iir_chain_kernel<T, KK>;                                                                                   \
        static bool configured_dev[64] = {};                                                                                  \
        if (!configured_dev[b.device & 63]) {                                                                                 \
            SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IIR_CHAIN_MAX_BYTES));      \
            configured_dev[b.device & 63] = true;                                                                             \
        }                                                                                                                     \
        kern<<<1, threads, smem, stream>>>(static_cast<T *>(data), (int)n_samples, static_cast<const T *>(b.d_coef),          \
                                           static_cast<T *>(b.d_state), b.sections);                                          \
    }
    switch (b.numerator) {
    case NUM_GENERIC: SDSP_CHAIN(NUM_GENERIC) break;
    case NUM_LP: SDSP_CHAIN(NUM_LP) break;
    case NUM_HP: SDSP_CHAIN(NUM_HP) break;
    default: SDSP_CHAIN(NUM_BP) break;
    }
#undef SDSP_CHAIN
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

int iir_launch_sequential(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    if (b.precision == SDSP_B200_F32)
        return launch_seq_sections<float>(b, data, n_samples, stride, stream);
    return launch_seq_sections<double>(b, data, n_samples, stride, stream);
}

// one channel whose samples fit shared memory: sections spread over warps (same bits as every other sequential kernel)
bool iir_chain_applicable(const IirBank &b, size_t n_samples)
{
    static const bool disabled = getenv("SDSP_B200_NO_CHAIN") != nullptr;
    const size_t es = b.precision == SDSP_B200_F32 ? 4 : 8;
    return !disabled && b.n_channels == 1 && b.sections >= 2 && n_samples >= 2 * CH_TILE && n_samples * es <= IIR_CHAIN_MAX_BYTES;
}
int iir_launch_chain(const IirBank &b, void *data, size_t n_samples, cudaStream_t stream)
{
    return b.precision == SDSP_B200_F32 ? launch_chain<float>(b, data, n_samples, stream) : launch_chain<double>(b, data, n_samples, stream);
}

// =================================================================================================
// host emulation of the sequential path (same iir_step, compiled for the host)
template <typename T, int M, int KIND>
static void emulate_seq(double gain, const double *b, const double *a, double *mem, double *diff, void *data, size_t n)
{
    IirCoef<T, M> c;
    IirState<T, M> s;
    iir_pack_coef<T, M>(c, gain, b, a);
    iir_state_from_mem<T, M>(s, mem);
    if (diff && IirDelta<T>::value)
        for (int j = 0; j < M; j++)
            s.d[j] = (T)diff[j];
    T *d = static_cast<T *>(data);
    // as the TMA kernel does: whole tiles with the sections skewed, the ragged tail sample by sample
    constexpr int TS = 2 * 128 / (int)sizeof(T);
    size_t i = 0;
    for (; i + TS <= n; i += TS) {
        T *tile = d + i;
        iir_tile_dispatch<T, M, KIND, TS>(
            c, s, [&](int k) -> T { return tile[k]; }, [&](int k, T y) { tile[k] = y; });
    }
    for (; i < n; i++)
        d[i] = iir_step<T, M, KIND>(d[i], c, s);
    iir_state_to_mem<T, M>(s, mem);
    if (diff)
        for (int j = 0; j < M; j++)
            diff[j] = IirDelta<T>::value ? (double)s.d[j] : 0.0;
}

template <typename T, int M>
static void emulate_seq_kind(int kind, double gain, const double *b, const double *a, double *mem, double *diff, void *data, size_t n)
{
    switch (kind) {
    case NUM_GENERIC: emulate_seq<T, M, NUM_GENERIC>(gain, b, a, mem, diff, data, n); break;
    case NUM_LP: emulate_seq<T, M, NUM_LP>(gain, b, a, mem, diff, data, n); break;
    case NUM_HP: emulate_seq<T, M, NUM_HP>(gain, b, a, mem, diff, data, n); break;
    case NUM_GENERIC_B2ONE: emulate_seq<T, M, NUM_GENERIC_B2ONE>(gain, b, a, mem, diff, data, n); break;
    default: emulate_seq<T, M, NUM_BP>(gain, b, a, mem, diff, data, n); break;
    }
}

template <typename T>
static int emulate_seq_sections(int sections, int kind, double gain, const double *b, const double *a, double *mem, double *diff,
                                void *data, size_t n)
{
    switch (sections) {
#define X(MM) \
    case MM: emulate_seq_kind<T, MM>(kind, gain, b, a, mem, diff, data, n); return SDSP_B200_OK;
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#undef X
    default: return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir: sections=%d not built (1..8)", sections);
    }
}

// =================================================================================================
// Butterworth designers -- host scalar code.  Bilinear-transform prototype sections exactly as the
// reference parameterises them (casc_2o_iir.h:168-194 lp, :140-166 hp, :82-138 bp): section k has
//   beta  = (1 - t)/(1 + t)/2,  t = d_k sin(e)/2,  gamma = (1/2 + beta) cos(e),  a = {1, -2 gamma, 2 beta}
// with e the (warped) centre angle and d_k = 2 sin((2k+1) pi / (4 m)) the Butterworth damping.
static void design_section(double dk, double e, double &beta, double &gamma)
{
    const double t = dk * std::sin(e) / 2;
    beta = (1 - t) / (1 + t) / 2;
    gamma = (0.5 + beta) * std::cos(e);
}

static int design_lp_hp(int m, double f0, double fs, double gain_in, bool hp, double *gain, double *b, double *a)
{
    if (m < 1 || m > SDSP_B200_IIR_MAX_SECTIONS_ONCE || !gain || !b || !a)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_design: bad arguments (sections=%d)", m);
    double g = gain_in;
    const double e0 = 2 * M_PI * f0 / fs;
    for (int k = 0; k < m; k++) {
        const double dk = 2 * std::sin((2 * k + 1) * M_PI / (4.0 * m));
        double beta, gamma;
        design_section(dk, e0, beta, gamma);
        const double alpha = hp ? (0.5 + beta + gamma) / 4 : (0.5 + beta - gamma) / 4;
        g *= 2 * alpha;
        b[3 * k + 0] = 1.0;
        b[3 * k + 1] = hp ? -2.0 : 2.0;
        b[3 * k + 2] = 1.0;
        a[3 * k + 0] = 1.0;
        a[3 * k + 1] = -2 * gamma;
        a[3 * k + 2] = 2 * beta;
    }
    *gain = g;
    return SDSP_B200_OK;
}

static int design_bp(int m, double f0, double fs, double q, double gain_in, double *gain, double *b, double *a)
{
    if (m < 2 || m > SDSP_B200_IIR_MAX_SECTIONS_ONCE || (m % 2) || !gain || !b || !a)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_design_bp: sections must be even, 2..%d (got %d)", SDSP_B200_IIR_MAX_SECTIONS_ONCE, m);
    double g = gain_in;
    const double e0 = 2 * M_PI * f0 / fs;
    const double de = 2 * std::tan(e0 / (2 * q)) / std::sin(e0);
    const double th = std::tan(e0 / 2.0);
    for (int k = 0; k < m / 2; k++) {
        const double d = 2 * std::sin((2 * k + 1) * M_PI / (2.0 * m));
        const double aa = (1 + de * de / 4.0) * 2 / d / de;
        const double dk = std::sqrt(de * d / (aa + std::sqrt(aa * aa - 1)));
        const double bb = d * de / dk / 2.0;
        const double w = bb + std::sqrt(bb * bb - 1);
        const double e1 = 2.0 * std::atan(th / w);
        const double e2 = 2.0 * std::atan(w * th);
        double beta1, gamma1, beta2, gamma2;
        design_section(dk, e1, beta1, gamma1);
        design_section(dk, e2, beta2, gamma2);
        const double t = std::sqrt(1 + (w - 1 / w) / dk * (w - 1 / w) / dk);
        const double alpha1 = (0.5 - beta1) * t / 2.0;
        const double alpha2 = (0.5 - beta2) * t / 2.0;
        g *= 4 * alpha1 * alpha2;
        for (int h = 0; h < 2; h++) {
            double *bs = b + 3 * (2 * k + h), *as = a + 3 * (2 * k + h);
            bs[0] = 1.0;
            bs[1] = 0.0;
            bs[2] = -1.0;
            as[0] = 1.0;
            as[1] = -2 * (h ? gamma2 : gamma1);
            as[2] = 2 * (h ? beta2 : beta1);
        }
    }
    *gain = g;
    return SDSP_B200_OK;
}

// =================================================================================================
static int check_bank_args(int sections, int precision, int numerator)
{
    if (sections < 1 || sections > 8)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir: sections=%d not built (1..8)", sections);
    if (precision != SDSP_B200_F32 && precision != SDSP_B200_F64)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir: bad precision %d", precision);
    if (numerator < 0 || numerator > 3)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir: bad numerator kind %d", numerator);
    return SDSP_B200_OK;
}

template <typename T>
static void pack_coef_soa(std::vector<T> &out, int m, size_t count, const double *gain, const double *b, const double *a)
{
    // [k][c] with k = gain, b1[0..m), b2[0..m), fa[0..m), fb[0..m)  (feedback pair: see IirCoef)
    out.assign((size_t)iir_coef_count(m) * count, (T)0);
    for (size_t c = 0; c < count; c++) {
        out[c] = (T)gain[c];
        for (int j = 0; j < m; j++) {
            const double *bj = b ? b + (c * m + j) * 3 : nullptr;
            const double *aj = a + (c * m + j) * 3;
            out[(size_t)(1 + j) * count + c] = bj ? (T)bj[1] : (T)0;
            out[(size_t)(1 + m + j) * count + c] = bj ? (T)bj[2] : (T)0;
            double fa, fb;
            iir_feedback_pair(IirDelta<T>::value, aj[1], aj[2], fa, fb);
            out[(size_t)(1 + 2 * m + j) * count + c] = (T)fa;
            out[(size_t)(1 + 3 * m + j) * count + c] = (T)fb;
        }
    }
}

static size_t elem_size(int precision)
{
    return precision == SDSP_B200_F32 ? sizeof(float) : sizeof(double);
}

// rows of a [rows][n_channels] device array <-> [rows][count] host block at channel offset `first`
static int copy_rows(const IirBank &b, void *dev, void *host, int rows, size_t first, size_t count, bool to_device)
{
    const size_t es = elem_size(b.precision);
    char *d = static_cast<char *>(dev) + first * es;
    if (to_device)
        SDSP_CUDA(cudaMemcpy2D(d, b.n_channels * es, host, count * es, count * es, rows, cudaMemcpyHostToDevice));
    else
        SDSP_CUDA(cudaMemcpy2D(host, count * es, d, b.n_channels * es, count * es, rows, cudaMemcpyDeviceToHost));
    return SDSP_B200_OK;
}
} // namespace sdsp_b200

using namespace sdsp_b200;

struct sdsp_b200_iir_bank_s {
    IirBank b;
};

extern "C" int sdsp_b200_iir_bank_destroy(sdsp_b200_iir_bank bank);

// host layout mem[c][k] (k < rows)  <->  device rows [row0 + k][channel]
template <typename T>
static int state_rows_io(IirBank &b, int row0, int rows, size_t first, size_t count, double *mem, bool to_device)
{
    std::vector<T> soa((size_t)rows * count);
    void *dev = static_cast<char *>(b.d_state) + (size_t)row0 * b.n_channels * sizeof(T);
    if (to_device) {
        for (size_t c = 0; c < count; c++)
            for (int k = 0; k < rows; k++)
                soa[(size_t)k * count + c] = (T)mem[c * rows + k];
        return copy_rows(b, dev, soa.data(), rows, first, count, true);
    }
    int rc = copy_rows(b, dev, soa.data(), rows, first, count, false);
    for (size_t c = 0; c < count && rc == 0; c++)
        for (int k = 0; k < rows; k++)
            mem[c * rows + k] = (double)soa[(size_t)k * count + c];
    return rc;
}

static int state_io(sdsp_b200_iir_bank bank, size_t first, size_t count, double *mem, bool to_device, bool diff_rows)
{
    if (!bank || !mem)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_state: null argument");
    IirBank &b = bank->b;
    if (first + count > b.n_channels)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_state: channels [%zu,%zu) outside bank of %zu", first, first + count,
                         b.n_channels);
    if (count == 0)
        return SDSP_B200_OK;
    SDSP_CUDA(cudaSetDevice(b.device));
    const int m = b.sections, hist = iir_hist_count(m);
    if (diff_rows) { // the running differences of a delta-form (fp32) bank; an fp64 bank has none: reads give zeros
        if (b.precision != SDSP_B200_F32) {
            if (!to_device)
                memset(mem, 0, sizeof(double) * count * (size_t)m);
            return SDSP_B200_OK;
        }
        return state_rows_io<float>(b, hist, m, first, count, mem, to_device);
    }
    if (b.precision != SDSP_B200_F32)
        return state_rows_io<double>(b, 0, hist, first, count, mem, to_device);
    int rc = state_rows_io<float>(b, 0, hist, first, count, mem, to_device);
    if (rc == SDSP_B200_OK && to_device) {
        // a history that comes from outside starts the running differences at v[n-1] - v[n-2] (iir_state_from_mem)
        std::vector<double> diff(count * (size_t)m);
        for (size_t c = 0; c < count; c++)
            for (int j = 0; j < m; j++)
                diff[c * m + j] = (double)((float)mem[c * hist + 2 * (j + 1)] - (float)mem[c * hist + 2 * (j + 1) + 1]);
        rc = state_rows_io<float>(b, hist, m, first, count, diff.data(), true);
    }
    return rc;
}

static std::mutex g_once_mu;
static std::map<std::tuple<int, int, int>, sdsp_b200_iir_bank> g_once_cache;

void sdsp_b200::iir_release_process_once_cache()
{
    std::lock_guard<std::mutex> lock(g_once_mu);
    for (auto &kv : g_once_cache)
        sdsp_b200_iir_bank_destroy(kv.second);
    g_once_cache.clear();
}

extern "C" {

int sdsp_b200_iir_bank_create(sdsp_b200_iir_bank *bank, int sections, size_t n_channels, int precision, int numerator, int device)
{
    if (!bank)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_create: null out pointer");
    *bank = nullptr;
    int rc = check_bank_args(sections, precision, numerator);
    if (rc)
        return rc;
    if (n_channels == 0)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_create: n_channels must be positive");
    rc = ensure_device(device);
    if (rc)
        return rc;
    auto *h = new sdsp_b200_iir_bank_s();
    IirBank &b = h->b;
    b.sections = sections;
    b.n_channels = n_channels;
    b.precision = precision;
    b.numerator = numerator;
    b.device = device;
    b.sm_count = device_sm_count(device);
    const size_t es = elem_size(precision);
    const size_t coef_bytes = (size_t)iir_coef_count(sections) * n_channels * es;
    const size_t state_bytes = (size_t)iir_bank_state_rows(b) * n_channels * es;
    if (cudaMalloc(&b.d_coef, coef_bytes) != cudaSuccess || cudaMalloc(&b.d_state, state_bytes) != cudaSuccess) {
        cudaGetLastError();
        if (b.d_coef)
            cudaFree(b.d_coef);
        delete h;
        return set_error(SDSP_B200_ERR_OOM, "iir_bank_create: cannot allocate bank for %zu channels", n_channels);
    }
    cudaMemset(b.d_coef, 0, coef_bytes);
    cudaMemset(b.d_state, 0, state_bytes);
    *bank = h;
    return SDSP_B200_OK;
}

int sdsp_b200_iir_bank_destroy(sdsp_b200_iir_bank bank)
{
    if (!bank)
        return SDSP_B200_OK;
    IirBank &b = bank->b;
    cudaSetDevice(b.device);
    iir_bank_release_aux(b);
    if (b.d_coef)
        cudaFree(b.d_coef);
    if (b.d_state)
        cudaFree(b.d_state);
    if (b.d_stage)
        cudaFree(b.d_stage);
    b.host.release();
    delete bank;
    return SDSP_B200_OK;
}

int sdsp_b200_iir_bank_set_coeffs(sdsp_b200_iir_bank bank, size_t first, size_t count, const double *gain, const double *bco,
                                  const double *aco)
{
    if (!bank || !gain || !aco)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_set_coeffs: null argument");
    IirBank &b = bank->b;
    if (b.numerator == NUM_GENERIC && !bco)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_set_coeffs: a generic bank needs b coefficients");
    if (first + count > b.n_channels)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_set_coeffs: channels [%zu,%zu) outside bank of %zu", first, first + count,
                         b.n_channels);
    if (count == 0)
        return SDSP_B200_OK;
    SDSP_CUDA(cudaSetDevice(b.device));
    int rc;
    if (b.precision == SDSP_B200_F32) {
        std::vector<float> soa;
        pack_coef_soa<float>(soa, b.sections, count, gain, bco, aco);
        rc = copy_rows(b, b.d_coef, soa.data(), iir_coef_count(b.sections), first, count, true);
    } else {
        std::vector<double> soa;
        pack_coef_soa<double>(soa, b.sections, count, gain, bco, aco);
        rc = copy_rows(b, b.d_coef, soa.data(), iir_coef_count(b.sections), first, count, true);
    }
    if (rc == SDSP_B200_OK) {
        // keep a host copy in double: the scan path derives its propagation tables from it
        const int m = b.sections;
        if (b.h_gain.size() != b.n_channels) {
            b.h_gain.assign(b.n_channels, 1.0);
            b.h_b.assign(b.n_channels * m * 3, 0.0);
            b.h_a.assign(b.n_channels * m * 3, 0.0);
        }
        for (size_t c = 0; c < count; c++) {
            b.h_gain[first + c] = gain[c];
            for (int k = 0; k < 3 * m; k++) {
                b.h_b[(first + c) * 3 * m + k] = bco ? bco[c * 3 * m + k] : 0.0;
                b.h_a[(first + c) * 3 * m + k] = aco[c * 3 * m + k];
            }
        }
        b.coef_version++;
        static const int b2one_mode = getenv("SDSP_B200_IIR_B2ONE") ? atoi(getenv("SDSP_B200_IIR_B2ONE")) : -1; // comparison aid: 0 off, 1 on
        b.b2_all_one = b.numerator == NUM_GENERIC && b.precision == SDSP_B200_F32 && b2one_mode != 0;
        for (size_t i = 0; b.b2_all_one && i < b.n_channels * (size_t)m; i++)
            b.b2_all_one = b.h_b[3 * i + 2] == 1.0;
    }
    return rc;
}

int sdsp_b200_iir_bank_set_state(sdsp_b200_iir_bank bank, size_t first, size_t count, const double *mem)
{
    return state_io(bank, first, count, const_cast<double *>(mem), true, false);
}

int sdsp_b200_iir_bank_get_state(sdsp_b200_iir_bank bank, size_t first, size_t count, double *mem)
{
    return state_io(bank, first, count, mem, false, false);
}

int sdsp_b200_iir_bank_set_state_diff(sdsp_b200_iir_bank bank, size_t first, size_t count, const double *diff)
{
    return state_io(bank, first, count, const_cast<double *>(diff), true, true);
}

int sdsp_b200_iir_bank_get_state_diff(sdsp_b200_iir_bank bank, size_t first, size_t count, double *diff)
{
    return state_io(bank, first, count, diff, false, true);
}

int sdsp_b200_iir_bank_reset_state(sdsp_b200_iir_bank bank)
{
    if (!bank)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_reset_state: null bank");
    IirBank &b = bank->b;
    SDSP_CUDA(cudaSetDevice(b.device));
    SDSP_CUDA(cudaMemset(b.d_state, 0, (size_t)iir_bank_state_rows(b) * b.n_channels * elem_size(b.precision)));
    return SDSP_B200_OK;
}

int sdsp_b200_iir_bank_process(sdsp_b200_iir_bank bank, void *data, size_t n_samples, size_t channel_stride, int ptr_kind, int path,
                               void *stream)
{
    if (!bank)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_process: null bank");
    if (n_samples == 0)
        return SDSP_B200_OK;
    if (!data)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_process: null data");
    IirBank &b = bank->b;
    if (channel_stride < n_samples && b.n_channels > 1)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_process: channel_stride %zu < n_samples %zu", channel_stride, n_samples);
    if (path < SDSP_B200_IIR_AUTO || path > SDSP_B200_IIR_SCAN_SPLIT)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_process: bad path %d", path);
    const size_t es = elem_size(b.precision);
    if (reinterpret_cast<uintptr_t>(data) % es)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_process: data not aligned to its element size");
    SDSP_CUDA(cudaSetDevice(b.device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);

    if (ptr_kind == SDSP_B200_PTR_DEVICE)
        return iir_dispatch(b, data, n_samples, channel_stride, path, s);
    if (ptr_kind != SDSP_B200_PTR_HOST)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_process: bad ptr_kind %d", ptr_kind);

    // host data (see host_stage.h): synchronous; anything the caller queued on `stream` (device-pointer calls that
    // moved this bank's history) is waited for first
    std::lock_guard<std::mutex> lock(b.mu);
    if (s)
        SDSP_CUDA(cudaStreamSynchronize(s));
    int rc = b.host.ensure();
    if (rc)
        return rc;
    // The stream is cut along TIME into chunks that span all channels; the bank's history carries from chunk to chunk
    // exactly as it does from call to call, so on the sequential path the result is bit-identical to one uncut call.
    // Staging memory is two chunks, whatever the size of the bank (config 3 would need 64 GiB for an uncut copy).
    // Chunk rows are 256-byte multiples so every chunk takes the same kernel (TMA path: 16-byte aligned pitch).
    const size_t row_quantum = 256 / es;
    size_t chunk = host_slab_bytes() / (b.n_channels * es) / row_quantum * row_quantum;
    if (chunk < 4 * row_quantum)
        chunk = 4 * row_quantum;
    if (chunk > n_samples)
        chunk = n_samples;
    const size_t pitch_elems = (chunk * es + 15) / 16 * 16 / es;
    const size_t chunk_bytes = b.n_channels * pitch_elems * es;
    const int nbuf = n_samples > chunk ? 2 : 1;
    rc = ensure_device_stage(b.d_stage, b.stage_bytes, chunk_bytes * nbuf, "iir_bank_process");
    if (rc)
        return rc;
    const bool small = nbuf == 1 && b.n_channels * n_samples * es <= HostStage::BOUNCE_BYTES;
    if (small) { // one filter object / a handful of channels: latency path through the pinned bounce buffer
        cudaStream_t cs = b.host.stream[0];
        char *bounce = static_cast<char *>(b.host.bounce);
        for (size_t c = 0; c < b.n_channels; c++)
            memcpy(bounce + c * n_samples * es, static_cast<char *>(data) + c * channel_stride * es, n_samples * es);
        SDSP_CUDA(cudaMemcpy2DAsync(b.d_stage, pitch_elems * es, bounce, n_samples * es, n_samples * es, b.n_channels, cudaMemcpyHostToDevice, cs));
        rc = iir_dispatch(b, b.d_stage, n_samples, pitch_elems, path, cs);
        if (rc)
            return rc;
        SDSP_CUDA(cudaMemcpy2DAsync(bounce, n_samples * es, b.d_stage, pitch_elems * es, n_samples * es, b.n_channels, cudaMemcpyDeviceToHost, cs));
        SDSP_CUDA(cudaStreamSynchronize(cs));
        for (size_t c = 0; c < b.n_channels; c++)
            memcpy(static_cast<char *>(data) + c * channel_stride * es, bounce + c * n_samples * es, n_samples * es);
        return SDSP_B200_OK;
    }
    int which = 0;
    bool first = true;
    for (size_t done = 0; done < n_samples && rc == SDSP_B200_OK; done += chunk) {
        const size_t cnt = (n_samples - done) < chunk ? (n_samples - done) : chunk;
        char *h = static_cast<char *>(data) + done * es;
        char *d = static_cast<char *>(b.d_stage) + (size_t)which * chunk_bytes;
        cudaStream_t cs = b.host.stream[which];
        if (cudaMemcpy2DAsync(d, pitch_elems * es, h, channel_stride * es, cnt * es, b.n_channels, cudaMemcpyHostToDevice, cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "H2D", __FILE__, __LINE__);
        // the filter history is one object: chunk k + 1 starts where chunk k ended
        if (rc == SDSP_B200_OK && !first && cudaStreamWaitEvent(cs, b.host.kernel_done[which ^ 1], 0) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "cudaStreamWaitEvent", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK)
            rc = iir_dispatch(b, d, cnt, pitch_elems, path, cs);
        if (rc == SDSP_B200_OK && cudaEventRecord(b.host.kernel_done[which], cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "cudaEventRecord", __FILE__, __LINE__);
        if (rc == SDSP_B200_OK &&
            cudaMemcpy2DAsync(h, channel_stride * es, d, pitch_elems * es, cnt * es, b.n_channels, cudaMemcpyDeviceToHost, cs) != cudaSuccess)
            rc = cuda_fail((int)cudaGetLastError(), "D2H", __FILE__, __LINE__);
        which = (which + 1) % nbuf;
        first = false;
    }
    return b.host.drain(rc);
}

int sdsp_b200_iir_bank_describe(sdsp_b200_iir_bank bank, size_t n_samples, size_t channel_stride, int path, char *buf, size_t buf_len)
{
    if (!bank || !buf || !buf_len)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_bank_describe: bad arguments");
    return iir_describe(bank->b, n_samples, channel_stride, path, buf, buf_len);
}

int sdsp_b200_iir_design_lp(int sections, double f0, double fs, double gain_in, double *gain, double *b, double *a)
{
    return design_lp_hp(sections, f0, fs, gain_in, false, gain, b, a);
}
int sdsp_b200_iir_design_hp(int sections, double f0, double fs, double gain_in, double *gain, double *b, double *a)
{
    return design_lp_hp(sections, f0, fs, gain_in, true, gain, b, a);
}
int sdsp_b200_iir_design_bp(int sections, double f0, double fs, double q, double gain_in, double *gain, double *b, double *a)
{
    return design_bp(sections, f0, fs, q, gain_in, gain, b, a);
}

// preload_filter (casc_2o_iir.h:196-214): every history slot of row 0 holds value*gain; for a low-pass
// design each following row is the previous one times the section's DC gain sum(b)/(1+a1+a2); for
// every other type the later rows are zero.
int sdsp_b200_iir_preload_state(int sections, int filter_type, double gain, const double *b, const double *a, double value,
                                double *mem)
{
    if (sections < 1 || sections > SDSP_B200_IIR_MAX_SECTIONS_ONCE || !mem || (filter_type == SDSP_B200_LOW_PASS && (!a || !b)))
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_preload_state: bad arguments");
    double v = value * gain;
    for (int k = 0; k < 2 * (sections + 1); k++)
        mem[k] = 0.0;
    mem[0] = mem[1] = v;
    if (filter_type == SDSP_B200_LOW_PASS) {
        for (int j = 1; j <= sections; j++) {
            v /= 1 + a[3 * (j - 1) + 1] + a[3 * (j - 1) + 2];
            v *= b[3 * (j - 1) + 0] + b[3 * (j - 1) + 1] + b[3 * (j - 1) + 2];
            mem[2 * j] = mem[2 * j + 1] = v;
        }
    }
    return SDSP_B200_OK;
}

// A small cache of one-channel banks serves the drop-in header's single-object process() calls.  The arithmetic is
// fp64 whatever the sample type: the reference object computes and keeps m_mem in double for any iterator value type
// (casc_2o_iir.h:13-18, 45-71), so float samples are widened on the way in and rounded once on the way out, and the
// history the caller holds (double mem[sections+1][2]) round-trips exactly -- block-wise calls stay bit-identical to
// one whole-buffer call (test/testIIR.cpp:61-75).  fp32 ARITHMETIC is what sdsp_b200_iir_bank_* with SDSP_B200_F32 is for.
// one group of at most 8 sections (what the kernels are built for)
static int process_once_group(int sections, int numerator, int precision, double gain, const double *bco, const double *aco, double *mem,
                              void *data, size_t n_samples, int device)
{
    int rc = check_bank_args(sections, precision, numerator);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(g_once_mu);
    auto key = std::make_tuple(sections, numerator, device);
    auto it = g_once_cache.find(key);
    if (it == g_once_cache.end()) {
        sdsp_b200_iir_bank nb = nullptr;
        rc = sdsp_b200_iir_bank_create(&nb, sections, 1, SDSP_B200_F64, numerator, device);
        if (rc)
            return rc;
        it = g_once_cache.emplace(key, nb).first;
    }
    sdsp_b200_iir_bank bk = it->second;
    IirBank &ob = bk->b;
    // coefficients: uploaded only when they differ from what the cached bank holds (a filter object is designed once and
    // then called many times)
    const int m = sections;
    bool same = ob.h_gain.size() == 1 && ob.h_gain[0] == gain;
    for (int k = 0; same && k < 3 * m; k++)
        same = ob.h_a[k] == aco[k] && ob.h_b[k] == (bco ? bco[k] : 0.0);
    if (!same) {
        rc = sdsp_b200_iir_bank_set_coeffs(bk, 0, 1, &gain, bco, aco);
        if (rc)
            return rc;
    }
    const int hist = iir_hist_count(m);
    const size_t data_bytes = n_samples * sizeof(double), hist_bytes = (size_t)hist * sizeof(double);
    if (hist_bytes + data_bytes <= HostStage::BOUNCE_BYTES) {
        // latency path: history and samples bounce through pinned memory, five asynchronous operations, one synchronisation.
        // (one channel: the device rows [2r + i][1] are laid out exactly like mem[r][i])
        std::lock_guard<std::mutex> bank_lock(ob.mu);
        SDSP_CUDA(cudaSetDevice(ob.device));
        rc = ob.host.ensure();
        if (!rc)
            rc = ensure_device_stage(ob.d_stage, ob.stage_bytes, (data_bytes + 15) / 16 * 16, "iir_process_once");
        if (rc)
            return rc;
        cudaStream_t cs = ob.host.stream[0];
        double *bounce = static_cast<double *>(ob.host.bounce);
        memcpy(bounce, mem, hist_bytes);
        if (precision == SDSP_B200_F32) {
            const float *src = static_cast<const float *>(data);
            for (size_t i = 0; i < n_samples; i++)
                bounce[hist + i] = (double)src[i];
        } else {
            memcpy(bounce + hist, data, data_bytes);
        }
        SDSP_CUDA(cudaMemcpyAsync(ob.d_state, bounce, hist_bytes, cudaMemcpyHostToDevice, cs));
        SDSP_CUDA(cudaMemcpyAsync(ob.d_stage, bounce + hist, data_bytes, cudaMemcpyHostToDevice, cs));
        rc = iir_dispatch(ob, ob.d_stage, n_samples, (n_samples + 1) / 2 * 2, SDSP_B200_IIR_SEQUENTIAL, cs);
        if (rc)
            return rc;
        SDSP_CUDA(cudaMemcpyAsync(bounce, ob.d_state, hist_bytes, cudaMemcpyDeviceToHost, cs));
        SDSP_CUDA(cudaMemcpyAsync(bounce + hist, ob.d_stage, data_bytes, cudaMemcpyDeviceToHost, cs));
        SDSP_CUDA(cudaStreamSynchronize(cs));
        memcpy(mem, bounce, hist_bytes);
        if (precision == SDSP_B200_F32) {
            float *dst = static_cast<float *>(data);
            for (size_t i = 0; i < n_samples; i++)
                dst[i] = (float)bounce[hist + i];
        } else {
            memcpy(data, bounce + hist, data_bytes);
        }
        return SDSP_B200_OK;
    }
    std::vector<double> wide;
    double *work = static_cast<double *>(data);
    if (precision == SDSP_B200_F32) {
        const float *src = static_cast<const float *>(data);
        wide.assign(src, src + n_samples);
        work = wide.data();
    }
    rc = sdsp_b200_iir_bank_set_state(bk, 0, 1, mem);
    if (!rc)
        rc = sdsp_b200_iir_bank_process(bk, work, n_samples, n_samples, SDSP_B200_PTR_HOST, SDSP_B200_IIR_SEQUENTIAL, nullptr);
    if (!rc)
        rc = sdsp_b200_iir_bank_get_state(bk, 0, 1, mem);
    if (!rc && precision == SDSP_B200_F32) {
        float *dst = static_cast<float *>(data);
        for (size_t i = 0; i < n_samples; i++)
            dst[i] = (float)wide[i];
    }
    return rc;
}

int sdsp_b200_iir_process_once(int sections, int numerator, int precision, double gain, const double *bco, const double *aco,
                               double *mem, void *data, size_t n_samples, int device)
{
    if (sections < 1 || sections > SDSP_B200_IIR_MAX_SECTIONS_ONCE)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir_process_once: sections=%d (1..%d)", sections, SDSP_B200_IIR_MAX_SECTIONS_ONCE);
    if ((precision != SDSP_B200_F32 && precision != SDSP_B200_F64) || numerator < 0 || numerator > 3)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_process_once: bad precision / numerator");
    if (!aco || !mem || (numerator == NUM_GENERIC && !bco))
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_process_once: null argument");
    if (n_samples == 0)
        return SDSP_B200_OK;
    if (!data)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir_process_once: null data");
    if (sections <= 8)
        return process_once_group(sections, numerator, precision, gain, bco, aco, mem, data, n_samples, device);
    // More sections than a kernel holds (the reference's template takes any even m_t, casc_2o_iir.h:23-26): a cascade is a chain,
    // so groups of eight run one after another over the block -- group g + 1 with gain 1 (x * 1 is exact) on group g's output.
    // Row `first` of the history belongs to both neighbours (output history of one, input history of the other): the later
    // group must start from its value BEFORE this block, which the earlier group has overwritten by then.
    std::vector<double> wide;
    double *work = static_cast<double *>(data);
    if (precision == SDSP_B200_F32) { // keep fp64 between the groups; round once at the end
        const float *src = static_cast<const float *>(data);
        wide.assign(src, src + n_samples);
        work = wide.data();
    }
    double shared_before[2] = { 0, 0 };
    for (int first = 0; first < sections; first += 8) {
        const int cnt = sections - first < 8 ? sections - first : 8;
        double gmem[2 * 9];
        memcpy(gmem, mem + 2 * first, sizeof(double) * 2 * (size_t)(cnt + 1));
        if (first > 0) {
            gmem[0] = shared_before[0];
            gmem[1] = shared_before[1];
        }
        shared_before[0] = mem[2 * (first + cnt)];
        shared_before[1] = mem[2 * (first + cnt) + 1];
        const int rc = process_once_group(cnt, numerator, SDSP_B200_F64, first == 0 ? gain : 1.0, bco ? bco + 3 * first : nullptr,
                                          aco + 3 * first, gmem, work, n_samples, device);
        if (rc)
            return rc;
        memcpy(mem + 2 * (first + (first > 0 ? 1 : 0)), gmem + (first > 0 ? 2 : 0), sizeof(double) * 2 * (size_t)(cnt + (first > 0 ? 0 : 1)));
    }
    if (precision == SDSP_B200_F32) {
        float *dst = static_cast<float *>(data);
        for (size_t i = 0; i < n_samples; i++)
            dst[i] = (float)wide[i];
    }
    return SDSP_B200_OK;
}

// host emulation of the sequential kernels.  diff (may be null): the running differences of the fp32 delta form,
// [sections] doubles in / out next to mem -- with them a stream cut into calls reproduces the uncut run bit for bit;
// without, each call starts them at v[n-1] - v[n-2] as sdsp_b200_iir_bank_set_state does.
int sdsp_b200_debug_emulate_iir_diff(int sections, int numerator, int precision, double gain, const double *bco, const double *aco,
                                     double *mem, double *diff, void *data, size_t n_samples)
{
    // (numerator 4 = the kernels' internal "generic with every b2 == 1" kind, so that its bit-identity with the generic kind
    // can be checked on the host)
    int rc = check_bank_args(sections, precision, numerator == NUM_GENERIC_B2ONE ? NUM_GENERIC : numerator);
    if (rc)
        return rc;
    if (!aco || !mem || !data || ((numerator == NUM_GENERIC || numerator == NUM_GENERIC_B2ONE) && !bco))
        return set_error(SDSP_B200_ERR_INVALID_ARG, "emulate_iir: null argument");
    if (precision == SDSP_B200_F32)
        return emulate_seq_sections<float>(sections, numerator, gain, bco, aco, mem, diff, data, n_samples);
    return emulate_seq_sections<double>(sections, numerator, gain, bco, aco, mem, diff, data, n_samples);
}

int sdsp_b200_debug_emulate_iir(int sections, int numerator, int precision, double gain, const double *bco, const double *aco,
                                double *mem, void *data, size_t n_samples)
{
    return sdsp_b200_debug_emulate_iir_diff(sections, numerator, precision, gain, bco, aco, mem, nullptr, data, n_samples);
}
}
