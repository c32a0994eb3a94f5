// iir_scan.cu -- precision dispatch of the look-back scan path (iir_scan_impl.cuh) and its host emulation entry.
#include <cstdint>

#include "iir_core.cuh"
#include "iir_internal.h"
#include "tma_ptx.cuh"

namespace sdsp_b200
{
int iir_scan_chunk_f32();
int iir_scan_chunk_f64();
int iir_launch_scan_tiles_f32(IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream, size_t *done);
int iir_launch_scan_tiles_f64(IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream, size_t *done);
int iir_emulate_scan_f32(int sections, int numerator, double gain, const double *b, const double *a, double *mem, void *data,
                         size_t n_samples, int chunk, bool force_general);
int iir_emulate_scan_f64(int sections, int numerator, double gain, const double *b, const double *a, double *mem, void *data,
                         size_t n_samples, int chunk, bool force_general);

int iir_scan_chunk(int precision)
{
    return precision == SDSP_B200_F32 ? iir_scan_chunk_f32() : iir_scan_chunk_f64();
}

bool iir_scan_applicable(const IirBank &b, const void *data, size_t n_samples, size_t stride)
{
    const int L = iir_scan_chunk(b.precision);
    if (reinterpret_cast<uintptr_t>(data) % 16 != 0 || (b.n_channels > 1 && stride % L != 0) || !get_encode_fn())
        return false;
    if (b.sections != 2 && b.sections != 4 && b.sections != 6 && b.sections != 8)
        return false;
    return n_samples >= (size_t)32 * L && b.h_gain.size() == b.n_channels;
}

// whole tiles through the scan kernel, the ragged remainder through the sequential kernel (which picks the
// bank history up where the scan left it)
int iir_launch_scan(IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    const size_t es = b.precision == SDSP_B200_F32 ? 4 : 8;
    const int L = iir_scan_chunk(b.precision);
    if (reinterpret_cast<uintptr_t>(data) % 16 != 0 || (b.n_channels > 1 && stride % L != 0) || !get_encode_fn())
        return set_error(SDSP_B200_ERR_UNSUPPORTED,
                         "iir scan path needs a 16-byte aligned base and (for several channels) a channel stride that is a multiple of %d samples", L);
    size_t done = 0;
    int rc = b.precision == SDSP_B200_F32 ? iir_launch_scan_tiles_f32(b, data, n_samples, stride, stream, &done) :
                                            iir_launch_scan_tiles_f64(b, data, n_samples, stride, stream, &done);
    if (rc)
        return rc;
    if (done < n_samples)
        rc = iir_launch_sequential(b, static_cast<char *>(data) + done * es, n_samples - done, stride, stream);
    return rc;
}
} // namespace sdsp_b200

using namespace sdsp_b200;

extern "C" int sdsp_b200_debug_emulate_iir_scan(int sections, int numerator, int precision, double gain, const double *b, const double *a,
                                                double *mem, void *data, size_t n_samples, int chunk, int force_general)
{
    if (!a || !mem || !data || (numerator == NUM_GENERIC && !b) || numerator < 0 || numerator > 3)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "emulate_iir_scan: bad argument");
    if (chunk < 8 || chunk > 4096)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "emulate_iir_scan: chunk must be in [8, 4096]");
    if (precision == SDSP_B200_F32)
        return iir_emulate_scan_f32(sections, numerator, gain, b, a, mem, data, n_samples, chunk, force_general != 0);
    if (precision == SDSP_B200_F64)
        return iir_emulate_scan_f64(sections, numerator, gain, b, a, mem, data, n_samples, chunk, force_general != 0);
    return set_error(SDSP_B200_ERR_INVALID_ARG, "emulate_iir_scan: bad precision");
}
