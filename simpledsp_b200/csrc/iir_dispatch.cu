// iir_dispatch.cu -- chooses how a bank walks the time axis.
//
// Reference test/testIIR.cpp:61-75 cuts a stream into 32-sample calls and demands bit-identical output.  Every
// sequential kernel evaluates iir_step() in the same order, so any mix of them keeps that promise; the time-parallel
// kernels reassociate (within the parity tolerance, not bit for bit).  SDSP_B200_IIR_AUTO therefore resolves to a
// sequential kernel EXCEPT where that kernel cannot use the machine and the call is far longer than anything a
// block-streaming caller issues -- the rule (iir_auto_time_split):
//     warps = ceil(channels / 32) <= 2 x SMs   (fewer than half the GPU's 4 x SMs warp schedulers would have a warp:
//                                               9472 channels on a B200; such a bank runs below 45 % of the roofline)
//     AND  n_samples >= 65536                  (16 x the 4096-sample buffers of the reference's tests and benchmarks)
//     AND  the time-split path applies         (aligned layout, filter memory << call length, see iir_segment.cu)
// -> time-split.  Both conditions are properties of the bank and of the call length only; a caller that needs
// bit-identical re-blocking at such lengths asks for SDSP_B200_IIR_SEQUENTIAL (or sets SDSP_B200_IIR_AUTO_SPLIT=0).
#include <cstdio>
#include <cstdlib>

#include "iir_internal.h"

namespace sdsp_b200
{
// Both sequential kernels produce identical bits (same iir_section() calls in the same order), so which
// one runs may depend on the call: the TMA kernel needs a 16-byte aligned base and pitch and pays off
// once a warp has a few stages of work.
static bool iir_use_tma(const IirBank &b, const void *data, size_t n_samples, size_t stride)
{
    static const bool disabled = getenv("SDSP_B200_NO_TMA") != nullptr;
    if (disabled || !iir_tma_built_for(b.sections))
        return false;
    if (n_samples < 256)
        return false;
    return iir_tma_applicable(b, data, n_samples, stride);
}

static bool iir_auto_time_split(IirBank &b, const void *data, size_t n_samples, size_t stride)
{
    static const bool enabled = []() {
        const char *e = getenv("SDSP_B200_IIR_AUTO_SPLIT");
        return !e || atoi(e) != 0;
    }();
    if (!enabled || n_samples < 65536 || (b.n_channels + 31) / 32 > (size_t)2 * b.sm_count)
        return false;
    return iir_segment_applicable(b, data, n_samples, stride);
}

int iir_dispatch(IirBank &b, void *data, size_t n_samples, size_t stride, int path, cudaStream_t stream)
{
    // time-parallel: the time-split path when the filter's memory fits a segment of this call, else the look-back scan where
    // the layout allows it, else (call too short to split, or an unsuitable layout) the sequential kernels -- the request is
    // for speed, the result is the same stream either way
    if (path == SDSP_B200_IIR_SCAN) {
        if (iir_segment_applicable(b, data, n_samples, stride))
            return iir_launch_segmented(b, data, n_samples, stride, stream);
        if (iir_scan_applicable(b, data, n_samples, stride))
            return iir_launch_scan(b, data, n_samples, stride, stream);
        path = SDSP_B200_IIR_AUTO;
    }
    if (path == SDSP_B200_IIR_SCAN_LOOKBACK)
        return iir_launch_scan(b, data, n_samples, stride, stream);
    if (path == SDSP_B200_IIR_SCAN_SPLIT) {
        if (!iir_segment_applicable(b, data, n_samples, stride))
            return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir time-split path needs a 16-byte aligned layout, a filter whose natural response "
                                                        "vanishes within a segment, and fewer channels than the GPU has lanes");
        return iir_launch_segmented(b, data, n_samples, stride, stream);
    }
    if (path == SDSP_B200_IIR_AUTO && iir_auto_time_split(b, data, n_samples, stride))
        return iir_launch_segmented(b, data, n_samples, stride, stream);
    if (iir_chain_applicable(b, n_samples)) // one channel, a few thousand samples: sections spread over warps (same bits)
        return iir_launch_chain(b, data, n_samples, stream);
    if (iir_use_tma(b, data, n_samples, stride))
        return iir_launch_tma(b, data, n_samples, stride, stream);
    return iir_launch_sequential(b, data, n_samples, stride, stream);
}

int iir_describe(IirBank &b, size_t n_samples, size_t stride, int path, char *buf, size_t buf_len)
{
    // alignment of the (unknown here) base pointer is assumed: describe() reports the kernel a 16-byte aligned call gets
    bool timepar = path == SDSP_B200_IIR_SCAN || path == SDSP_B200_IIR_SCAN_LOOKBACK || path == SDSP_B200_IIR_SCAN_SPLIT;
    const bool auto_split = path == SDSP_B200_IIR_AUTO && b.h_gain.size() == b.n_channels && iir_use_tma(b, nullptr, n_samples, stride) &&
                            iir_auto_time_split(b, nullptr, n_samples, stride);
    if (auto_split) {
        timepar = true;
        path = SDSP_B200_IIR_SCAN;
    }
    const bool chain = !timepar && iir_chain_applicable(b, n_samples);
    const bool tma = !timepar && !chain && iir_use_tma(b, nullptr, n_samples, stride);
    char how[512];
    snprintf(how, sizeof how, "%s", chain ? "sequential/chain (one channel in shared memory, one warp per section)" :
                                    tma   ? "sequential/tma (warp per 32 channels, skewed sections, TMA ring per warp through the blocked view)" :
                                            "sequential/generic");
    if (timepar) {
        const bool split = path != SDSP_B200_IIR_SCAN_LOOKBACK && b.h_gain.size() == b.n_channels && iir_use_tma(b, nullptr, n_samples, stride) &&
                           iir_segment_describe(b, n_samples, how, sizeof how) == 0;
        const bool lookback = !split && (path == SDSP_B200_IIR_SCAN_LOOKBACK || (stride % iir_scan_chunk(b.precision) == 0 || b.n_channels == 1));
        if (!split && lookback)
            snprintf(how, sizeof how, "look-back scan (warp per 32 chunks of %d samples, Kogge-Stone over lanes, decoupled look-back between tiles)",
                     iir_scan_chunk(b.precision));
        else if (!split)
            snprintf(how, sizeof how, "%s (no time-parallel kernel applies to this call)",
                     iir_use_tma(b, nullptr, n_samples, stride) ? "sequential/tma" : "sequential/generic");
    }
    snprintf(buf, buf_len, "iir bank: %zu channels x %d sections %s numerator=%d; n_samples=%zu stride=%zu path=%s -> %s", b.n_channels,
             b.sections, b.precision == SDSP_B200_F32 ? "f32" : "f64", b.numerator, n_samples, stride,
             auto_split ? "auto (few channels, long call)" : timepar ? "time-parallel" : path == SDSP_B200_IIR_SEQUENTIAL ? "sequential" : "auto", how);
    return SDSP_B200_OK;
}

void iir_bank_release_aux(IirBank &b)
{
    if (b.d_scan_tables)
        cudaFree(b.d_scan_tables);
    if (b.d_scan_flags)
        cudaFree(b.d_scan_flags);
    if (b.d_state_alt)
        cudaFree(b.d_state_alt);
    if (b.d_seg_state)
        cudaFree(b.d_seg_state);
    b.d_scan_tables = b.d_scan_flags = b.d_state_alt = b.d_seg_state = nullptr;
    b.seg_state_bytes = 0;
}
} // namespace sdsp_b200

