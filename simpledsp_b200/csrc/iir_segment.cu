// iir_segment.cu -- the time-split IIR path for filters whose memory is shorter than a segment.
//
// What it replaces: casc_2o_iir<m_t>::process (reference include/sdsp/casc_2o_iir.h:36-80) on long streams
// with few channels (BASELINE config 4: one channel of 2^30 samples; the >= 4096-channel point of config 3),
// where lane-per-channel alone cannot fill 148 SMs.
//
// The cascade is linear, so for a stretch of samples entered with history (u, c) -- u the two last scaled inputs
// (row 0 of the reference's m_mem), c the two last outputs of every section (rows 1..m) --
//       response(x; u, c) = response(x; u, 0)  +  response(0; 0, c).
// The second term is the filter's natural response: it decays like |pole|^n.  Let K be the number of samples
// after which the one-step transition matrix of the cascade, raised to the n-th power, is below 2^-62 (fp64) /
// 2^-32 (fp32) in every entry -- 2^-9 (fp64) / 2^-8 (fp32) of one ulp of a value as large as the state it multiplies.  Cut each
// channel into `segs` segments of seg_len >= K samples.  Then
//   1. seg_gather: every segment learns its incoming scaled-input history u from the two samples before it
//      (read before anything is overwritten: the filter runs in place); segment 0 takes the bank's history;
//   2. the lane-per-row TMA kernel (iir_tma.cu, ROWS_SEG) runs all channels x segs rows at once from (u, 0):
//      this is the bandwidth-bound pass, identical in speed to a bank of channels x segs channels;
//   3. seg_carry: c entering segment s is the section history segment s-1 ended with -- exactly the state
//      hand-off of a chunked state-space scan, whose propagation term A^seg_len c_in(s-1) has underflowed below
//      the threshold above because seg_len >= K; the last segment's history becomes the bank's;
//   4. the same TMA kernel in ROWS_SEG_ACC mode adds response(0; 0, c) to the first K samples of every segment.
// HBM traffic is (1 + K/seg_len) x the algorithmic bytes; no pass is serial in time.  Filters with K larger
// than any reasonable segment (|pole| -> 1) go to the look-back scan kernel in iir_scan.cu instead, which
// carries the propagation term explicitly.  Leftover samples (n not a multiple of segs*seg_len) are split
// again, and the final few go through the sequential kernel, which continues from the bank history.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "iir_core.cuh"
#include "iir_internal.h"

namespace sdsp_b200
{
// ---- host: how long does the natural response take to vanish -----------------------------------------------
// One step of the cascade with zero input on the section history c (rows 1..m), in double, from the bank's host
// copy of the coefficients; the fixed-numerator kinds take their numerators from the kind.
static void natural_step_matrix(int m, int kind, const double *b, const double *a, std::vector<double> &t)
{
    const int sd = 2 * m;
    t.assign((size_t)sd * sd, 0.0);
    std::vector<double> h(sd);
    for (int col = 0; col < sd; col++) {
        for (int k = 0; k < sd; k++)
            h[k] = k == col ? 1.0 : 0.0;
        double in0 = 0, in1 = 0, in2 = 0;
        for (int j = 0; j < m; j++) {
            double b1, b2;
            switch (kind) {
            case NUM_LP: b1 = 2, b2 = 1; break;
            case NUM_HP: b1 = -2, b2 = 1; break;
            case NUM_BP: b1 = 0, b2 = -1; break;
            default: b1 = b[3 * j + 1], b2 = b[3 * j + 2]; break;
            }
            const double v1 = h[2 * j], v2 = h[2 * j + 1];
            const double v = in0 + b1 * in1 + b2 * in2 - a[3 * j + 1] * v1 - a[3 * j + 2] * v2;
            h[2 * j + 1] = v1;
            h[2 * j] = v;
            in0 = v;
            in1 = v1;
            in2 = v2;
        }
        for (int k = 0; k < sd; k++)
            t[(size_t)k * sd + col] = h[k];
    }
}
static void matmul(int sd, const std::vector<double> &x, const std::vector<double> &y, std::vector<double> &z)
{
    z.assign((size_t)sd * sd, 0.0);
    for (int i = 0; i < sd; i++)
        for (int k = 0; k < sd; k++) {
            const double xv = x[(size_t)i * sd + k];
            if (xv != 0.0)
                for (int j = 0; j < sd; j++)
                    z[(size_t)i * sd + j] += xv * y[(size_t)k * sd + j];
        }
}
static double maxabs(const std::vector<double> &x)
{
    double m = 0;
    for (double v : x) {
        if (!(v == v))
            return INFINITY;
        m = fabs(v) > m ? fabs(v) : m;
    }
    return m;
}
// smallest n with max|T^n| <= negligible (by squaring, then descending over the binary digits); 0 = never (within 2^40)
static unsigned long long decay_length_one(int m, int kind, const double *b, const double *a, double negligible)
{
    const int sd = 2 * m;
    std::vector<std::vector<double>> pw(1);
    natural_step_matrix(m, kind, b, a, pw[0]);
    if (maxabs(pw[0]) <= negligible)
        return 1;
    int j = 0;
    for (;;) {
        if (j >= 40)
            return 0;
        pw.emplace_back();
        matmul(sd, pw[j], pw[j], pw[j + 1]);
        j++;
        const double nm = maxabs(pw[j]);
        if (!(nm < 1e100))
            return 0; // growing: unstable filter
        if (nm <= negligible)
            break;
    }
    // T^(2^j) is negligible, T^(2^(j-1)) is not: find the largest n in between that is not
    unsigned long long n = 1ull << (j - 1);
    std::vector<double> q = pw[j - 1], r;
    for (int i = j - 2; i >= 0; i--) {
        matmul(sd, q, pw[i], r);
        if (maxabs(r) > negligible) {
            q.swap(r);
            n += 1ull << i;
        }
    }
    return n + 1;
}

static double negligible_for(int precision)
{
    return precision == SDSP_B200_F32 ? 2.3283064365386963e-10 /* 2^-32 */ : 2.1684043449710089e-19 /* 2^-62 */;
}

// the bank's decay length: the slowest channel's, with a margin; cached per coefficient version
unsigned long long iir_decay_length(IirBank &b)
{
    if (b.decay_version == b.coef_version && b.decay_len_valid)
        return b.decay_len;
    const double negligible = negligible_for(b.precision);
    unsigned long long worst = 1;
    const int m = b.sections;
    // identical coefficient sets are common in a bank (bench: a few thousand distinct designs): memoise the last one
    const double *pb = nullptr, *pa = nullptr;
    unsigned long long last = 0;
    for (size_t ch = 0; ch < b.n_channels && worst; ch++) {
        const double *cb = &b.h_b[ch * 3 * m], *ca = &b.h_a[ch * 3 * m];
        bool same = pb != nullptr;
        for (int k = 0; same && k < 3 * m; k++)
            same = cb[k] == pb[k] && ca[k] == pa[k];
        const unsigned long long d = same ? last : decay_length_one(m, b.numerator, cb, ca, negligible);
        pb = cb, pa = ca, last = d;
        if (d == 0)
            worst = 0;
        else if (d > worst)
            worst = d;
    }
    b.decay_len = worst ? worst + worst / 8 + 16 : 0; // 0 = some channel never decays
    b.decay_version = b.coef_version;
    b.decay_len_valid = true;
    return b.decay_len;
}

// ---- device: the two small kernels around the row passes --------------------------------------------------
template <typename T>
__global__ void seg_gather_kernel(const T *__restrict__ data, size_t stride, size_t seg_len, unsigned segs, size_t rows, int state_rows,
                                  const T *__restrict__ coef, const T *__restrict__ bank_state, size_t n_channels, T *__restrict__ row_state)
{
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= rows)
        return;
    const size_t ch = v / segs;
    const unsigned s = (unsigned)(v % segs);
    if (s == 0) {
        for (int k = 0; k < state_rows; k++)
            row_state[(size_t)k * rows + v] = bank_state[(size_t)k * n_channels + ch];
        return;
    }
    const T *p = data + ch * stride + (size_t)s * seg_len;
    const T gain = coef[ch];
    row_state[v] = mul_t(p[-1], gain); // the same product iir_step() forms for row 0 of the history
    row_state[rows + v] = mul_t(p[-2], gain);
    for (int k = 2; k < state_rows; k++)
        row_state[(size_t)k * rows + v] = (T)0;
}

template <typename T>
__global__ void seg_carry_kernel(const T *__restrict__ row_state, unsigned segs, size_t rows, int state_rows, T *__restrict__ acc_state,
                                 T *__restrict__ bank_state, size_t n_channels)
{
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= rows)
        return;
    const size_t ch = v / segs;
    const unsigned s = (unsigned)(v % segs);
    acc_state[v] = (T)0;
    acc_state[rows + v] = (T)0;
    for (int k = 2; k < state_rows; k++)
        acc_state[(size_t)k * rows + v] = s ? row_state[(size_t)k * rows + v - 1] : (T)0;
    if (s + 1 == segs)
        for (int k = 0; k < state_rows; k++)
            bank_state[(size_t)k * n_channels + ch] = row_state[(size_t)k * rows + v];
}

// ---- host: planning --------------------------------------------------------------------------------------
// How many segments per channel?  The row pass keeps slots = SMs x (row-warps per SM) warps resident, 32 rows
// each; rows are long streams, so a last, partly filled wave costs a whole wave.  Cost model in row-samples:
//   waves(segs) x (seg_len + corr)  +  leftover samples x 4       (leftovers are re-split at low occupancy)
// minimised over segs = 8, 16, ... (8-row TMA boxes must stay inside one channel) with seg_len >= min_len.
// SDSP_B200_SEG_ROWS=<rows> (tuning aid) pins segs = rows / channels instead.
bool iir_segment_plan(IirBank &b, size_t n_samples, bool first_round, size_t *segs_out, size_t *seg_len_out, size_t *corr_out)
{
    // the search below is a pure function of (coefficients, n_samples, first_round): remember the last two answers
    for (auto &m : b.seg_plan_memo)
        if (m.valid && m.coef_version == b.coef_version && m.n_samples == n_samples && m.first_round == first_round) {
            *segs_out = m.segs;
            *seg_len_out = m.seg_len;
            *corr_out = m.corr;
            return m.ok;
        }
    auto remember = [&](bool ok, size_t segs, size_t seg_len, size_t corr) {
        auto &m = b.seg_plan_memo[first_round ? 0 : 1];
        m = { true, ok, first_round, b.coef_version, n_samples, segs, seg_len, corr };
        return ok;
    };
    const unsigned long long K = iir_decay_length(b);
    if (K == 0)
        return remember(false, 0, 0, 0);
    const size_t es = b.precision == SDSP_B200_F32 ? 4 : 8;
    const size_t cts = 2 * (128 / es); // the row kernel's compute tile = one ring stage
    const size_t align = cts;           // segments are whole stages (no ragged tail in every row), starts 256-byte aligned
    const size_t corr = (size_t)((K + cts - 1) / cts * cts);
    // the correction pass costs corr/seg_len extra traffic: <= 25 % on the bulk, anything on the leftovers
    const size_t min_len = first_round ? (corr * 4 > 1024 ? corr * 4 : 1024) : corr;
    const size_t max_segs = n_samples / min_len / 8 * 8;
    if (max_segs < 8)
        return remember(false, 0, 0, 0);
    static long pinned = -1;
    if (pinned < 0) {
        const char *e = getenv("SDSP_B200_SEG_ROWS");
        pinned = e ? atol(e) : 0;
    }
    const size_t slots = (size_t)b.sm_count * iir_tma_rows_slots_per_sm(b);
    size_t best = 0;
    if (pinned > 0) {
        best = (size_t)pinned / b.n_channels / 8 * 8;
        best = best > max_segs ? max_segs : best;
    } else {
        double best_cost = 0;
        const size_t limit = slots * 32 * 16 / b.n_channels; // beyond 16 waves nothing is left to gain
        bool four_waves_seen = false; // measured (profiles/r01_iir_split_ring_sweep.txt): once a bank fills four waves, cutting it
                                      // finer only adds correction traffic -- the wave model overrates what more segments buy
        for (size_t segs = 8; segs <= max_segs && segs <= limit && !four_waves_seen; segs += 8) {
            const size_t len = n_samples / segs / align * align;
            if (len < min_len)
                break;
            const size_t warps = (b.n_channels * segs + 31) / 32;
            const size_t waves = (warps + slots - 1) / slots;
            four_waves_seen = warps >= 4 * slots;
            const double cost = (double)waves * (double)(len + corr) + 4.0 * (double)(n_samples - segs * len);
            if (best == 0 || cost < best_cost) {
                best = segs;
                best_cost = cost;
            }
        }
    }
    if (best < 8)
        return remember(false, 0, 0, 0);
    const size_t seg_len = n_samples / best / align * align;
    if (seg_len < corr || seg_len >= (1ull << 31))
        return remember(false, 0, 0, 0);
    *segs_out = best;
    *seg_len_out = seg_len;
    *corr_out = corr;
    return remember(true, best, seg_len, corr);
}

bool iir_segment_applicable(IirBank &b, const void *data, size_t n_samples, size_t stride)
{
    if (!iir_tma_applicable(b, data, n_samples, stride) || !iir_tma_built_for(b.sections))
        return false;
    if (b.h_gain.size() != b.n_channels)
        return false;
    size_t segs, seg_len, corr;
    return iir_segment_plan(b, n_samples, true, &segs, &seg_len, &corr);
}

template <typename T>
static int segment_round(IirBank &b, T *data, size_t stride, size_t segs, size_t seg_len, size_t corr, cudaStream_t stream)
{
    // all rows of the bank's state array travel: the section history AND (fp32) the running differences a segment ends
    // with are what its successor's natural response starts from
    const int state_rows = iir_bank_state_rows(b);
    const size_t rows = b.n_channels * segs;
    const size_t need = 2 * (size_t)state_rows * rows * sizeof(T);
    if (b.seg_state_bytes < need) {
        if (b.d_seg_state)
            cudaFree(b.d_seg_state);
        b.d_seg_state = nullptr;
        b.seg_state_bytes = 0;
        if (cudaMalloc(&b.d_seg_state, need) != cudaSuccess) {
            cudaGetLastError();
            return set_error(SDSP_B200_ERR_OOM, "iir time-split: cannot allocate %zu bytes of segment history", need);
        }
        b.seg_state_bytes = need;
    }
    T *row_state = static_cast<T *>(b.d_seg_state);
    T *acc_state = row_state + (size_t)state_rows * rows;
    const unsigned tb = 256, grid = (unsigned)((rows + tb - 1) / tb);
    seg_gather_kernel<T><<<grid, tb, 0, stream>>>(data, stride, seg_len, (unsigned)segs, rows, state_rows, static_cast<const T *>(b.d_coef),
                                                  static_cast<const T *>(b.d_state), b.n_channels, row_state);
    SDSP_CUDA(cudaGetLastError());
    int rc = iir_launch_tma_rows(b, data, seg_len, segs, stride, row_state, seg_len, false, stream);
    if (rc)
        return rc;
    seg_carry_kernel<T><<<grid, tb, 0, stream>>>(row_state, (unsigned)segs, rows, state_rows, acc_state, static_cast<T *>(b.d_state),
                                                 b.n_channels);
    SDSP_CUDA(cudaGetLastError());
    return iir_launch_tma_rows(b, data, seg_len, segs, stride, acc_state, corr < seg_len ? corr : seg_len, true, stream);
}

// whole stream: rounds of segments while they pay, then the sequential kernel for what is left
int iir_launch_segmented(IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    const size_t es = b.precision == SDSP_B200_F32 ? 4 : 8;
    size_t done = 0;
    for (int round = 0; round < 8 && done < n_samples; round++) {
        size_t segs, seg_len, corr;
        if (!iir_segment_plan(b, n_samples - done, round == 0, &segs, &seg_len, &corr))
            break;
        char *p = static_cast<char *>(data) + done * es;
        const int rc = b.precision == SDSP_B200_F32 ? segment_round<float>(b, reinterpret_cast<float *>(p), stride, segs, seg_len, corr, stream) :
                                                      segment_round<double>(b, reinterpret_cast<double *>(p), stride, segs, seg_len, corr, stream);
        if (rc)
            return rc;
        done += segs * seg_len;
    }
    if (done == 0)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir time-split path: stream too short for the filter's memory (or unstable filter)");
    if (done < n_samples) {
        char *p = static_cast<char *>(data) + done * es;
        if (iir_tma_applicable(b, p, n_samples - done, stride) && n_samples - done >= 256)
            return iir_launch_tma(b, p, n_samples - done, stride, stream);
        return iir_launch_sequential(b, p, n_samples - done, stride, stream);
    }
    return SDSP_B200_OK;
}

// host only (no device): the memory of one filter in samples, as the planner sees it (0 = does not decay)
unsigned long long iir_decay_length_of(int sections, int numerator, int precision, const double *b, const double *a)
{
    const unsigned long long d = decay_length_one(sections, numerator, b, a, negligible_for(precision));
    return d ? d + d / 8 + 16 : 0; // the same margin the bank-wide figure carries (the envelope of the response is not monotonic)
}

int iir_segment_describe(IirBank &b, size_t n_samples, char *buf, size_t buf_len)
{
    size_t segs = 0, seg_len = 0, corr = 0;
    if (!iir_segment_plan(b, n_samples, true, &segs, &seg_len, &corr))
        return -1;
    snprintf(buf, buf_len, "time-split: %zu segments x %zu samples per channel as rows of the lane-per-row TMA kernel, natural-response "
                           "correction over the first %zu samples of each (filter memory %llu samples)",
             segs, seg_len, corr, iir_decay_length(b));
    return 0;
}
} // namespace sdsp_b200

using namespace sdsp_b200;

// verification aid (host only): samples after which the natural response of a cascade is below the time-split path's threshold
extern "C" int sdsp_b200_debug_iir_decay_length(int sections, int numerator, int precision, const double *b, const double *a,
                                                unsigned long long *samples)
{
    if (!a || !samples || sections < 1 || sections > 8 || numerator < 0 || numerator > 3 || (numerator == NUM_GENERIC && !b) ||
        (precision != SDSP_B200_F32 && precision != SDSP_B200_F64))
        return set_error(SDSP_B200_ERR_INVALID_ARG, "debug_iir_decay_length: bad argument");
    *samples = iir_decay_length_of(sections, numerator, precision, b, a);
    return SDSP_B200_OK;
}
