// iir_scan.cu -- the time-parallel IIR path (chunked state-space scan), see iir_scan_core.cuh.
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "iir_internal.h"
#include "iir_scan_core.cuh"

namespace sdsp_b200
{
// =================================================================================================
// host emulation
template <typename T, int M, int KIND>
static int emulate_scan(double gain, const double *b, const double *a, double *mem, void *data, size_t n, int L, bool force_general)
{
    IirCoef<T, M> c;
    IirState<T, M> s;
    iir_pack_coef<T, M>(c, gain, b, a);
    for (int r = 0; r <= M; r++) {
        s.h[r][0] = (T)mem[2 * r];
        s.h[r][1] = (T)mem[2 * r + 1];
    }
    std::vector<double> tab;
    int reach = 0;
    if (scan_build_tables(M, KIND, gain, b, a, L, tab, reach) != 0)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "scan: sections=%d not built", M);
    T *d = static_cast<T *>(data);
    const size_t done = scan_emulate_channel<T, M, KIND>(c, s, tab, reach, L, d, n, force_general);
    for (size_t i = done; i < n; i++)
        d[i] = iir_step<T, M, KIND>(d[i], c, s);
    for (int r = 0; r <= M; r++) {
        mem[2 * r] = (double)s.h[r][0];
        mem[2 * r + 1] = (double)s.h[r][1];
    }
    return SDSP_B200_OK;
}

template <typename T, int M>
static int emulate_scan_kind(int kind, double gain, const double *b, const double *a, double *mem, void *data, size_t n, int L, bool fg)
{
    switch (kind) {
    case NUM_GENERIC: return emulate_scan<T, M, NUM_GENERIC>(gain, b, a, mem, data, n, L, fg);
    case NUM_LP: return emulate_scan<T, M, NUM_LP>(gain, b, a, mem, data, n, L, fg);
    case NUM_HP: return emulate_scan<T, M, NUM_HP>(gain, b, a, mem, data, n, L, fg);
    default: return emulate_scan<T, M, NUM_BP>(gain, b, a, mem, data, n, L, fg);
    }
}

template <typename T>
static int emulate_scan_sections(int m, int kind, double gain, const double *b, const double *a, double *mem, void *data, size_t n, int L,
                                 bool fg)
{
    switch (m) {
    case 2: return emulate_scan_kind<T, 2>(kind, gain, b, a, mem, data, n, L, fg);
    case 4: return emulate_scan_kind<T, 4>(kind, gain, b, a, mem, data, n, L, fg);
    case 6: return emulate_scan_kind<T, 6>(kind, gain, b, a, mem, data, n, L, fg);
    case 8: return emulate_scan_kind<T, 8>(kind, gain, b, a, mem, data, n, L, fg);
    default: return set_error(SDSP_B200_ERR_UNSUPPORTED, "scan: sections=%d not built (2, 4, 6, 8)", m);
    }
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
        return emulate_scan_sections<float>(sections, numerator, gain, b, a, mem, data, n_samples, chunk, force_general != 0);
    if (precision == SDSP_B200_F64)
        return emulate_scan_sections<double>(sections, numerator, gain, b, a, mem, data, n_samples, chunk, force_general != 0);
    return set_error(SDSP_B200_ERR_INVALID_ARG, "emulate_iir_scan: bad precision");
}
