// iir_tma.cu -- applicability and precision dispatch of the TMA-fed IIR kernels (iir_tma_impl.cuh).
#include <cstdint>

#include "iir_internal.h"
#include "tma_ptx.cuh"

namespace sdsp_b200
{
int iir_launch_tma_f32(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream);
int iir_launch_tma_f64(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream);
int iir_launch_tma_rows_f32(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state, size_t n_samples,
                            bool accumulate, cudaStream_t stream, int *slots);
int iir_launch_tma_rows_f64(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state, size_t n_samples,
                            bool accumulate, cudaStream_t stream, int *slots);

bool iir_tma_applicable(const IirBank &b, const void *data, size_t n_samples, size_t stride)
{
    const size_t es = b.precision == SDSP_B200_F32 ? 4 : 8;
    if (reinterpret_cast<uintptr_t>(data) % 16 != 0)
        return false;
    if ((stride * es) % 16 != 0 && b.n_channels > 1)
        return false;
    if (n_samples >= (1ull << 31) || b.n_channels >= (1ull << 31))
        return false;
    return get_encode_fn() != nullptr;
}

bool iir_tma_built_for(int sections)
{
    return sections == 2 || sections == 4 || sections == 6 || sections == 8;
}

int iir_launch_tma(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    return b.precision == SDSP_B200_F32 ? iir_launch_tma_f32(b, data, n_samples, stride, stream) :
                                          iir_launch_tma_f64(b, data, n_samples, stride, stream);
}

// rows = segments: row (c, s) covers data[c*stride + s*seg_len + (0 .. n_samples-1)], n_samples <= seg_len; history of
// row r at row_state[k][r] (k < 2(m+1), rows = channels*segs wide).  accumulate: see ROWS_SEG_ACC in iir_tma_impl.cuh.
int iir_launch_tma_rows(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state, size_t n_samples,
                        bool accumulate, cudaStream_t stream)
{
    return b.precision == SDSP_B200_F32 ? iir_launch_tma_rows_f32(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, nullptr) :
                                          iir_launch_tma_rows_f64(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, nullptr);
}

// resident row-warps per SM of the row pass for this bank (what the planner sizes its waves by)
int iir_tma_rows_slots_per_sm(const IirBank &b)
{
    int slots = 0;
    const int rc = b.precision == SDSP_B200_F32 ? iir_launch_tma_rows_f32(b, nullptr, 0, 0, 0, nullptr, 0, false, nullptr, &slots) :
                                                  iir_launch_tma_rows_f64(b, nullptr, 0, 0, 0, nullptr, 0, false, nullptr, &slots);
    return rc == SDSP_B200_OK && slots > 0 ? slots : 4;
}
} // namespace sdsp_b200
