// iir_dispatch.cu -- chooses how a bank walks the time axis.
//
// The choice is a pure function of the bank configuration and the memory layout of the call
// (channel count, alignment), never of the call length alone: reference test/testIIR.cpp:61-75 cuts a
// stream into 32-sample calls and demands bit-identical output, and every sequential kernel evaluates
// iir_step() in the same order, so SDSP_B200_IIR_AUTO only ever resolves to a sequential kernel unless
// the caller opts into the (reassociating) scan with SDSP_B200_IIR_SCAN.
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
    if (iir_use_tma(b, data, n_samples, stride))
        return iir_launch_tma(b, data, n_samples, stride, stream);
    return iir_launch_sequential(b, data, n_samples, stride, stream);
}

int iir_describe(IirBank &b, size_t n_samples, size_t stride, int path, char *buf, size_t buf_len)
{
    // alignment of the (unknown here) base pointer is assumed: describe() reports the kernel a 16-byte aligned call gets
    const bool timepar = path == SDSP_B200_IIR_SCAN || path == SDSP_B200_IIR_SCAN_LOOKBACK || path == SDSP_B200_IIR_SCAN_SPLIT;
    const bool tma = !timepar && iir_use_tma(b, nullptr, n_samples, stride);
    char how[512];
    snprintf(how, sizeof how, "%s", tma ? "sequential/tma (warp per 32 channels, skewed sections, 6-stage TMA ring per warp)" : "sequential/generic");
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
             timepar ? "time-parallel" : path == SDSP_B200_IIR_SEQUENTIAL ? "sequential" : "auto", how);
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

