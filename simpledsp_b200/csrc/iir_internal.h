// iir_internal.h -- bank object and the launchers shared by the IIR translation units (internal).
#pragma once
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "common.h"
#include "host_stage.h"

namespace sdsp_b200
{
struct IirBank {
    int sections = 0, precision = 0, numerator = 0, device = 0, sm_count = 0;
    size_t n_channels = 0;
    void *d_coef = nullptr;  // [1 + 4m][n_channels]   gain, b1[m], b2[m], -a1[m], -a2[m]
    void *d_state = nullptr; // [2(m + 1) (+ m)][n_channels] rows 2r, 2r+1: x[n-1], x[n-2] of history row r; fp32: + m rows of running differences
    void *d_state_alt = nullptr; // scan path: the launch reads d_state and writes here, then the two are swapped
    // host copy of the coefficients in double (scan tables are derived from it)
    std::vector<double> h_gain, h_b, h_a;
    unsigned long coef_version = 0;
    bool b2_all_one = false; // generic bank, every b2 == 1: the kernels skip that multiply (same bits), see NUM_GENERIC_B2ONE
    // scan path: per-channel propagation tables on the device, rebuilt when the coefficients change
    void *d_scan_tables = nullptr;
    size_t scan_tables_bytes = 0;
    unsigned long scan_tables_version = ~0ul;
    int scan_chunk = 0;
    int scan_reach_max = 0;
    unsigned scan_epoch = 0;
    void *d_scan_flags = nullptr;
    size_t scan_flags_bytes = 0;
    // time-split path (iir_segment.cu): filter memory in samples (0 = never decays) and per-segment history
    unsigned long long decay_len = 0;
    unsigned long decay_version = ~0ul;
    bool decay_len_valid = false;
    void *d_seg_state = nullptr;
    size_t seg_state_bytes = 0;
    struct SegPlanMemo {
        bool valid = false, ok = false, first_round = false;
        unsigned long coef_version = 0;
        size_t n_samples = 0, segs = 0, seg_len = 0, corr = 0;
    } seg_plan_memo[2];
    // host staging
    void *d_stage = nullptr;
    size_t stage_bytes = 0;
    HostStage host;
    std::mutex mu;
};

// rows of the device-side state array (iir_core.cuh: the fp32 delta form carries one running difference per section)
inline int iir_bank_state_rows(const IirBank &b)
{
    return 2 * (b.sections + 1) + (b.precision == SDSP_B200_F32 ? b.sections : 0);
}

// fft.cu
void fft_release_l2_persist(); // sdsp_b200_shutdown()
// iir.cu
void iir_release_process_once_cache(); // sdsp_b200_shutdown()
int iir_launch_sequential(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream);
constexpr size_t IIR_CHAIN_MAX_BYTES = 160 * 1024; // samples of the one channel held in shared memory
bool iir_chain_applicable(const IirBank &b, size_t n_samples);
int iir_launch_chain(const IirBank &b, void *data, size_t n_samples, cudaStream_t stream);
// iir_tma.cu -- the fast sequential path (needs 16-byte aligned base and pitch, even section count)
bool iir_tma_applicable(const IirBank &b, const void *data, size_t n_samples, size_t stride);
bool iir_tma_built_for(int sections);
int iir_launch_tma(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream);
// rows = segments of channels (see iir_segment.cu)
int iir_launch_tma_rows(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state, size_t n_samples,
                        bool accumulate, cudaStream_t stream);
int iir_tma_rows_slots_per_sm(const IirBank &b);
// iir_segment.cu -- time-parallel path for filters whose memory is shorter than a segment
bool iir_segment_applicable(IirBank &b, const void *data, size_t n_samples, size_t stride);
int iir_launch_segmented(IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream);
int iir_segment_describe(IirBank &b, size_t n_samples, char *buf, size_t buf_len);
unsigned long long iir_decay_length(IirBank &b);
// iir_scan.cu -- time-parallel path, general (look-back carry of the propagation term)
int iir_launch_scan(IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream);
bool iir_scan_applicable(const IirBank &b, const void *data, size_t n_samples, size_t stride);
int iir_scan_chunk(int precision);
// iir_dispatch.cu
int iir_dispatch(IirBank &b, void *data, size_t n_samples, size_t stride, int path, cudaStream_t stream);
int iir_describe(IirBank &b, size_t n_samples, size_t stride, int path, char *buf, size_t buf_len);
void iir_bank_release_aux(IirBank &b);
} // namespace sdsp_b200
