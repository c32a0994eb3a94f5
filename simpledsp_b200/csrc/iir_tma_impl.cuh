// iir_tma_impl.cuh -- the fast sequential IIR path (compiled once per precision, see iir_tma_f32.cu / iir_tma_f64.cu): TMA-fed, one warp per 32 channels, sections software-skewed.
//
// Replaces casc_2o_iir<m_t>::process (reference include/sdsp/casc_2o_iir.h:36-80) for large banks.
//
// Data layout in HBM is the reference's: one contiguous range per channel, data[channel*stride + n].
// A lane-per-channel kernel therefore wants a [32 channels x TS samples] patch transposed on chip.  The
// TMA engine does that for free: a 2-D tensor map over (samples, channels) with boxes of 8 rows x 128
// bytes (four per 32 channels) and SWIZZLE_128B lands each channel's 128-byte run in its own shared-memory row, XOR-swizzled so
// that lane r reading 16-byte chunk c of row r (LDS.128 at r*128 + ((c ^ (r&7))<<4)) is conflict free.
// A second, "blocked" view of the same memory -- (32-sample block interior, channel, block index), the block index being
// the OUTERMOST tensor dimension with a stride of 128 bytes -- lets one TMA instruction fetch SUB consecutive 128-byte
// pieces of 8 channels ([SUB][8 rows][128 B] in shared memory: every 1 KB unit is one swizzle atom, so the lane
// addressing does not change).  A warp that has an SM sub-partition to itself pays for every instruction it issues,
// TMA set-up included: the blocked view halves (SUB = 2) the TMA instructions per stage.  Whole stages use it; the
// ragged last stage goes box by box through the plain view, whose extent clips at n_samples.
// Every warp runs its own ring of stages with its own mbarriers -- no block-wide synchronisation -- and
// writes results back with TMA stores from the same buffers (the filter runs in place, as in the
// reference).  Out-of-range rows / samples are zero-filled on load and clipped on store by the TMA unit,
// so ragged channel counts and lengths need no special addressing.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda.h>
#include <cuda_runtime.h>

#include "iir_core.cuh"
#include "iir_internal.h"
#include "tma_ptx.cuh"

namespace sdsp_b200
{
// TSB samples per box (one 128-byte row), SUB boxes per stage, NST stages per warp, PF stages of
// prefetch distance; WARPS independent warps per CTA.
//
// MODE (the time-split path, iir_segment.cu, reuses this kernel on "virtual channels" = segments of a channel):
//   ROWS_PLAIN   a row is a channel; 2-D tensor map (samples, channels)
//   ROWS_SEG     a row is segment (row % seg_per_ch) of channel (row / seg_per_ch); 3-D tensor map
//                (samples, segments, channels); coefficients are the channel's, history is the row's
//   ROWS_SEG_ACC same rows, but the cascade runs on zero input from the row's history and its output is
//                ADDED to the samples already there (the natural-response correction of a segment)
enum : int { ROWS_PLAIN = 0, ROWS_SEG = 1, ROWS_SEG_ACC = 2 };

template <typename T, int M, int KIND, int SUB, int CSUB, int NST, int PF, int WARPS, int RG, int MODE, bool PACK = true>
__global__ void __launch_bounds__(WARPS * 32)
    iir_tma_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap bmap, int n_samples,
                   const T *__restrict__ coef, T *__restrict__ state, size_t n_channels, size_t n_coef_channels, unsigned seg_per_ch,
                   int use_blocked)
{
    constexpr int TSB = 128 / (int)sizeof(T);
    constexpr int TS = TSB * SUB;   // samples per stage (what one mbarrier phase delivers)
    constexpr int CTS = TSB * CSUB; // samples per skewed compute tile
    static_assert(SUB % CSUB == 0 && RG == 8, "stage = whole compute tiles; boxes of 8 rows (one 1 KB swizzle atom)");
    // shared-memory layout of a stage: [row group g = 0..3][block u = 0..SUB-1][8 rows][128 B]
    constexpr int UNIT_BYTES = RG * 128;          // one (g, u) unit = one TMA box of the plain view
    constexpr int GROUP_BYTES = SUB * UNIT_BYTES; // one box of the blocked view
    constexpr int STAGE_BYTES = 32 * 128 * SUB;
    constexpr int VN = Vec16<T>::N;
    using V = typename Vec16<T>::type;
    static_assert(PF >= 1 && PF < NST, "prefetch distance must leave room for stores in flight");
    static_assert(MODE == ROWS_PLAIN || RG == 8, "segment rows are fetched in groups of 8 (segments per channel is a multiple of 8)");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bars[WARPS][NST];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t group = (size_t)blockIdx.x * WARPS + warp;
    const size_t ch0 = group * 32;
    if (ch0 >= n_channels)
        return; // warps never synchronise with one another
    const size_t ch = ch0 + lane;
    const bool active = ch < n_channels;
    unsigned char *ring = smem_raw + (size_t)warp * NST * STAGE_BYTES;
    uint64_t *bar = bars[warp];

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; s++)
            mbar_init(&bar[s], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();

    // coefficients and history of this lane's channel
    IirCoef<T, M> c;
    IirState<T, M> st;
    if (active) {
        const size_t cc = MODE == ROWS_PLAIN ? ch : ch / seg_per_ch; // whose coefficients
        iir_load_coef<T, M>(c, coef, n_coef_channels, cc);
        iir_load_state<T, M>(st, state, n_channels, ch);
    } else {
        iir_zero_coef<T, M>(c);
        iir_zero_state<T, M>(st);
    }

    const int n_stages = (n_samples + TS - 1) / TS;
    const int y0 = (int)ch0;
    // segment rows: (segment, channel) coordinates of this warp's four row groups (lane 0 issues the copies)
    int seg_of[MODE == ROWS_PLAIN ? 1 : 32 / RG], chn_of[MODE == ROWS_PLAIN ? 1 : 32 / RG];
    if constexpr (MODE != ROWS_PLAIN) {
#pragma unroll
        for (int g = 0; g < 32 / RG; g++) {
            const size_t v = ch0 + (size_t)g * RG;
            seg_of[g] = (int)(v % seg_per_ch);
            chn_of[g] = (int)(v / seg_per_ch); // beyond the last channel: zero-filled on load, clipped on store
        }
    }

    auto issue_load = [&](int k) { // lane 0 only
        uint64_t *b = &bar[k % NST];
        unsigned char *dst = ring + (size_t)(k % NST) * STAGE_BYTES;
        mbar_expect_tx(b, STAGE_BYTES);
        if (use_blocked && (k + 1) * TS <= n_samples) { // whole stage: one box of the blocked view per row group
#pragma unroll
            for (int g = 0; g < 32 / RG; g++)
                if constexpr (MODE == ROWS_PLAIN)
                    tma_load_3d(dst + g * GROUP_BYTES, &bmap, 0, y0 + g * RG, k * SUB, b);
                else
                    tma_load_4d(dst + g * GROUP_BYTES, &bmap, 0, seg_of[g], chn_of[g], k * SUB, b);
            return;
        }
        // ragged last stage: the SUB boxes of one row group issued back to back, so that the requests for consecutive
        // 128-byte pieces of a channel reach the memory system together (DRAM page locality)
#pragma unroll
        for (int g = 0; g < 32 / RG; g++)
#pragma unroll
            for (int u = 0; u < SUB; u++)
                if constexpr (MODE == ROWS_PLAIN)
                    tma_load_2d(dst + g * GROUP_BYTES + u * UNIT_BYTES, &map, k * TS + u * TSB, y0 + g * RG, b);
                else
                    tma_load_3d(dst + g * GROUP_BYTES + u * UNIT_BYTES, &map, k * TS + u * TSB, seg_of[g], chn_of[g], b);
    };

    if (lane == 0) {
        for (int k = 0; k < PF && k < n_stages; k++)
            issue_load(k);
    }

    // this lane's row inside its row group, and the XOR that un-swizzles 16-byte chunks
    const uint32_t row_off = (uint32_t)(lane >> 3) * (uint32_t)GROUP_BYTES + (uint32_t)(lane & 7) * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);

    for (int k = 0; k < n_stages; k++) {
        if (lane == 0 && k + PF < n_stages) {
            // the buffer about to be refilled was handed to a TMA store NST - PF steps ago: at most
            // NST - PF - 1 younger store groups may still be reading shared memory
            tma_wait_read<NST - PF - 1>();
            issue_load(k + PF);
        }
        mbar_wait(&bar[k % NST], (uint32_t)((k / NST) & 1));
        unsigned char *buf = ring + (size_t)(k % NST) * STAGE_BYTES;
        const int remaining = n_samples - k * TS;

#pragma unroll 1
        for (int ct = 0; ct < SUB / CSUB; ct++) {
            unsigned char *cbuf = buf + ct * CSUB * UNIT_BYTES;
            const int rem = remaining - ct * CTS;
            if (rem >= CTS) {
                V vin, vout;
                iir_tile_dispatch<T, M, KIND, CTS, PACK>(
                    c, st,
                    [&](int i) -> T {
                        if constexpr (MODE == ROWS_SEG_ACC)
                            return (T)0;
                        if (i % VN == 0) {
                            const int box = i / TSB, chunk = (i % TSB) / VN;
                            vin = *reinterpret_cast<const V *>(cbuf + box * UNIT_BYTES + row_off + (((uint32_t)chunk ^ sw) << 4));
                        }
                        return vget(vin, i % VN);
                    },
                    [&](int i, T y) {
                        if constexpr (MODE == ROWS_SEG_ACC) {
                            if (i % VN == 0) {
                                const int box = i / TSB, chunk = (i % TSB) / VN;
                                vout = *reinterpret_cast<const V *>(cbuf + box * UNIT_BYTES + row_off + (((uint32_t)chunk ^ sw) << 4));
                            }
                            y += vget(vout, i % VN);
                        }
                        vset(vout, i % VN, y);
                        if (i % VN == VN - 1) {
                            const int box = i / TSB, chunk = (i % TSB) / VN;
                            *reinterpret_cast<V *>(cbuf + box * UNIT_BYTES + row_off + (((uint32_t)chunk ^ sw) << 4)) = vout;
                        }
                    });
            } else if (rem > 0) {
                // ragged tail: plain sample-by-sample order (same arithmetic, see iir_core.cuh)
                for (int i = 0; i < rem; i++) {
                    const int box = i / TSB, chunk = (i % TSB) / VN, e = i % VN;
                    T *p = reinterpret_cast<T *>(cbuf + box * UNIT_BYTES + row_off + (((uint32_t)chunk ^ sw) << 4)) + e;
                    if constexpr (MODE == ROWS_SEG_ACC)
                        *p += iir_step<T, M, KIND>((T)0, c, st);
                    else
                        *p = iir_step<T, M, KIND>(*p, c, st);
                }
            }
        }
        fence_proxy_async(); // generic-proxy writes above -> visible to the TMA store below
        __syncwarp();
        if (lane == 0) {
            if (use_blocked && remaining >= TS) {
#pragma unroll
                for (int g = 0; g < 32 / RG; g++)
                    if constexpr (MODE == ROWS_PLAIN)
                        tma_store_3d(&bmap, 0, y0 + g * RG, k * SUB, buf + g * GROUP_BYTES);
                    else
                        tma_store_4d(&bmap, 0, seg_of[g], chn_of[g], k * SUB, buf + g * GROUP_BYTES);
            } else {
#pragma unroll
                for (int g = 0; g < 32 / RG; g++)
#pragma unroll
                    for (int u = 0; u < SUB; u++)
                        if constexpr (MODE == ROWS_PLAIN)
                            tma_store_2d(&map, k * TS + u * TSB, y0 + g * RG, buf + g * GROUP_BYTES + u * UNIT_BYTES);
                        else
                            tma_store_3d(&map, k * TS + u * TSB, seg_of[g], chn_of[g], buf + g * GROUP_BYTES + u * UNIT_BYTES);
            }
            tma_commit();
        }
    }
    if (lane == 0)
        tma_wait_all();
    if (active && MODE != ROWS_SEG_ACC) // (the correction pass ends in a history that is negligible by construction)
        iir_store_state<T, M>(st, state, n_channels, ch);
}

// SDSP_B200_IIR_BLOCKED=0 (comparison aid): every stage box by box through the plain view
static int iir_tma_blocked_view()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SDSP_B200_IIR_BLOCKED");
        v = e ? atoi(e) != 0 : 1;
    }
    return v;
}

template <typename T, int M, int KIND, int SUB, int CSUB, int NST, int PF, int WARPS, int RG, bool PACK = true>
static int launch_tma_cfg(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream, int promo)
{
    constexpr int TSB = 128 / (int)sizeof(T);
    CUtensorMap map;
    const cuuint64_t gdim[2] = { (cuuint64_t)n_samples, (cuuint64_t)b.n_channels };
    // a single channel has no second row: any 16-byte-multiple pitch is acceptable to the encoder
    const cuuint64_t pitch = b.n_channels > 1 ? (cuuint64_t)stride * sizeof(T) : (((cuuint64_t)n_samples * sizeof(T) + 15) / 16) * 16;
    const cuuint64_t gstride[1] = { pitch };
    const cuuint32_t box[2] = { (cuuint32_t)TSB, (cuuint32_t)RG };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUtensorMapL2promotion pr = promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B :
                                      promo == 64  ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
                                      promo == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE :
                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult r = get_encode_fn()(&map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, data, gdim,
                                 gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(SDSP_B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (n_samples=%zu channels=%zu pitch=%llu)", (int)r, n_samples,
                         b.n_channels, (unsigned long long)pitch);
    // the blocked view: (sample inside a 128-byte block, channel, block index); whole blocks only
    CUtensorMap bmap;
    {
        const cuuint64_t blocks = n_samples / TSB ? n_samples / TSB : 1;
        const cuuint64_t bdim[3] = { (cuuint64_t)TSB, (cuuint64_t)b.n_channels, blocks };
        const cuuint64_t bstride[2] = { pitch, 128 };
        const cuuint32_t bbox[3] = { (cuuint32_t)TSB, (cuuint32_t)RG, (cuuint32_t)SUB };
        const cuuint32_t bestr[3] = { 1, 1, 1 };
        r = get_encode_fn()(&bmap, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, data, bdim,
                            bstride, bbox, bestr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return set_error(SDSP_B200_ERR_CUDA, "cuTensorMapEncodeTiled (blocked view) failed with %d (n_samples=%zu channels=%zu pitch=%llu)",
                             (int)r, n_samples, b.n_channels, (unsigned long long)pitch);
    }
    auto kern = iir_tma_kernel<T, M, KIND, SUB, CSUB, NST, PF, WARPS, RG, ROWS_PLAIN, PACK>;
    constexpr size_t smem = (size_t)WARPS * NST * SUB * 32 * 128;
    static bool configured_dev[64] = {};
    bool &configured = configured_dev[b.device & 63]; // (the attribute is per device; a process may hold banks on several)
    if (!configured) {
        SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const size_t groups = (b.n_channels + 31) / 32;
    const unsigned grid = (unsigned)((groups + WARPS - 1) / WARPS);
    kern<<<grid, WARPS * 32, smem, stream>>>(map, bmap, (int)n_samples, static_cast<const T *>(b.d_coef), static_cast<T *>(b.d_state),
                                             b.n_channels, b.n_channels, 1u, iir_tma_blocked_view());
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

// ---- segment rows (time-split path): `segs` segments of seg_len samples per channel, each a row of its own
static int rows_tune()
{
    static int c = -1;
    if (c < 0) {
        const char *e = getenv("SDSP_B200_SEG_TUNE"); // kernel-tuning aid: ring shape of the row passes
        c = e ? atoi(e) : 0;
    }
    return c;
}

// slots: out-parameter query -- resident row-warps per SM for this configuration (no launch when data == nullptr)
template <typename T, int M, int KIND, int MODE, int SUB, int CSUB, int NST, int PF>
static int launch_rows_cfg(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state, size_t n_samples,
                           cudaStream_t stream, int *slots)
{
    constexpr int WARPS = 1, RG = 8;
    constexpr int TSB = 128 / (int)sizeof(T);
    auto kern = iir_tma_kernel<T, M, KIND, SUB, CSUB, NST, PF, WARPS, RG, MODE>;
    constexpr size_t smem = (size_t)WARPS * NST * SUB * 32 * 128;
    static bool configured_dev[64] = {};
    static int occ_dev[64] = {};
    bool &configured = configured_dev[b.device & 63]; // (the attribute is per device; a process may hold banks on several)
    int &occ = occ_dev[b.device & 63];
    if (!configured) {
        SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SDSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
        configured = true;
    }
    if (slots) {
        *slots = occ * WARPS;
        return SDSP_B200_OK;
    }
    CUtensorMap map;
    const cuuint64_t gdim[3] = { (cuuint64_t)seg_len, (cuuint64_t)segs, (cuuint64_t)b.n_channels };
    const cuuint64_t ch_pitch = b.n_channels > 1 ? (cuuint64_t)stride * sizeof(T) : (cuuint64_t)segs * seg_len * sizeof(T);
    const cuuint64_t gstride[2] = { (cuuint64_t)seg_len * sizeof(T), ch_pitch };
    const cuuint32_t box[3] = { (cuuint32_t)TSB, (cuuint32_t)RG, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = get_encode_fn()(&map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, data, gdim,
                                 gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(SDSP_B200_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with %d (seg_len=%zu segs=%zu channels=%zu pitch=%llu)", (int)r,
                         seg_len, segs, b.n_channels, (unsigned long long)ch_pitch);
    // blocked view: (sample inside a 128-byte block, segment, channel, block index inside the segment)
    CUtensorMap bmap;
    {
        const cuuint64_t blocks = seg_len / TSB ? seg_len / TSB : 1;
        const cuuint64_t bdim[4] = { (cuuint64_t)TSB, (cuuint64_t)segs, (cuuint64_t)b.n_channels, blocks };
        const cuuint64_t bstride[3] = { (cuuint64_t)seg_len * sizeof(T), ch_pitch, 128 };
        const cuuint32_t bbox[4] = { (cuuint32_t)TSB, (cuuint32_t)RG, 1, (cuuint32_t)SUB };
        const cuuint32_t bestr[4] = { 1, 1, 1, 1 };
        r = get_encode_fn()(&bmap, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, data, bdim,
                            bstride, bbox, bestr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return set_error(SDSP_B200_ERR_CUDA, "cuTensorMapEncodeTiled (blocked 4-D) failed with %d (seg_len=%zu segs=%zu channels=%zu)", (int)r,
                             seg_len, segs, b.n_channels);
    }
    const size_t rows = b.n_channels * segs;
    const size_t groups = (rows + 31) / 32;
    const unsigned grid = (unsigned)((groups + WARPS - 1) / WARPS);
    kern<<<grid, WARPS * 32, smem, stream>>>(map, bmap, (int)n_samples, static_cast<const T *>(b.d_coef), static_cast<T *>(row_state), rows,
                                             b.n_channels, (unsigned)segs, iir_tma_blocked_view());
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

template <typename T, int M, int KIND, int MODE>
static int launch_rows_ring(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state, size_t n_samples,
                            cudaStream_t stream, int *slots)
{
    if constexpr (M == 4 && KIND == NUM_GENERIC) { // alternatives exist for the headline instantiations only
        switch (rows_tune()) { //                                         SUB CSUB NST PF   ring per warp
        case 1: return launch_rows_cfg<T, M, KIND, MODE, 2, 2, 4, 2>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots); // 32 KB
        case 2: return launch_rows_cfg<T, M, KIND, MODE, 4, 2, 2, 1>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots); // 32 KB
        case 3: return launch_rows_cfg<T, M, KIND, MODE, 4, 2, 3, 1>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots); // 48 KB
        case 4: return launch_rows_cfg<T, M, KIND, MODE, 4, 2, 3, 2>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots); // 48 KB
        case 5: return launch_rows_cfg<T, M, KIND, MODE, 2, 2, 6, 3>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots); // 48 KB
        case 6: return launch_rows_cfg<T, M, KIND, MODE, 2, 2, 2, 1>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots); // 16 KB
        case 7: return launch_rows_cfg<T, M, KIND, MODE, 2, 2, 3, 1>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots); // 24 KB
        default: break;
        }
    }
    // defaults (ring sweeps with the blocked view: profiles/r02_iir_rows_ring_sweep.txt; round 1, before it: r01_iir_split_ring_sweep.txt):
    // fp32: three 8 KB stages per row-warp (9 row-warps per SM), fp64: six with three in flight (4 per SM)
    if constexpr (sizeof(T) == 4)
        return launch_rows_cfg<T, M, KIND, MODE, 2, 2, 3, 1>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots); // 24 KB
    else
        return launch_rows_cfg<T, M, KIND, MODE, 2, 2, 6, 3>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots); // 48 KB
}

template <typename T, int M, int KIND>
static int launch_rows_mode(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state, size_t n_samples,
                            bool accumulate, cudaStream_t stream, int *slots)
{
    if (accumulate)
        return launch_rows_ring<T, M, KIND, ROWS_SEG_ACC>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots);
    return launch_rows_ring<T, M, KIND, ROWS_SEG>(b, data, seg_len, segs, stride, row_state, n_samples, stream, slots);
}

template <typename T, int M>
static int launch_rows_kind(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state, size_t n_samples,
                            bool accumulate, cudaStream_t stream, int *slots)
{
    // (the b2 == 1 specialisation of the plain pass is not used here: measured neutral when several row-warps share a
    // scheduler, profiles/r02_iir_b2one_ab.txt)
    switch (b.numerator) {
    case NUM_GENERIC: return launch_rows_mode<T, M, NUM_GENERIC>(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, slots);
    case NUM_LP: return launch_rows_mode<T, M, NUM_LP>(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, slots);
    case NUM_HP: return launch_rows_mode<T, M, NUM_HP>(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, slots);
    default: return launch_rows_mode<T, M, NUM_BP>(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, slots);
    }
}

template <typename T>
static int launch_rows_sections(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state, size_t n_samples,
                                bool accumulate, cudaStream_t stream, int *slots)
{
    switch (b.sections) {
    case 2: return launch_rows_kind<T, 2>(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, slots);
    case 4: return launch_rows_kind<T, 4>(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, slots);
    case 6: return launch_rows_kind<T, 6>(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, slots);
    case 8: return launch_rows_kind<T, 8>(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, slots);
    default: return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir tma path: sections=%d not built", b.sections);
    }
}

template <typename T, int M, int KIND>
static int launch_tma(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    const int promo = 128; // TMA L2 promotion bytes (64 / 256 measured equal, profiles/r01_iir_tma_config_sweep.txt)
    // (A section-pipelined variant -- M/2 compute warps per 32 channels handing stages to one another -- was measured in
    // round 1 and dropped: the bound is the issue rate of the recurrence, not its length; profiles/r01_iir_pipe_vs_single.txt.)
    // single-warp CTAs, 8-row boxes; ring of three 128-byte x 4 stages, two in flight (profiles/r02_iir_tma_ring_sweep.txt;
    // round 1's sweep, before the blocked view: profiles/r01_iir_tma_config_sweep.txt).  fp32: scalar arithmetic while the bank
    // leaves schedulers to spare, packed once every SM holds its four warps; SDSP_B200_IIR_PACK=0|1 pins the choice
    // (comparison aid).  Same bits either way.
    bool pack = false;
    if constexpr (sizeof(T) == 4) {
        static int pin = -2;
        if (pin == -2) {
            const char *e = getenv("SDSP_B200_IIR_PACK");
            pin = e ? atoi(e) : -1;
        }
        pack = pin >= 0 ? pin != 0 : (b.n_channels + 31) / 32 >= (size_t)b.sm_count * 4;
    }
    if constexpr (M == 4 && KIND == NUM_GENERIC) { // SDSP_B200_TMA_TUNE (kernel-tuning aid): ring shape of the plain pass
        static int tune = -1;
        if (tune < 0) {
            const char *e = getenv("SDSP_B200_TMA_TUNE");
            tune = e ? atoi(e) : 0;
        }
        switch (tune) { //                                  SUB CSUB NST PF
        case 1: return launch_tma_cfg<T, M, KIND, 4, 2, 3, 1, 1, 8, false>(b, data, n_samples, stride, stream, promo);
        case 2: return launch_tma_cfg<T, M, KIND, 2, 2, 6, 3, 1, 8, false>(b, data, n_samples, stride, stream, promo);
        case 3: return launch_tma_cfg<T, M, KIND, 4, 2, 3, 2, 1, 8, false>(b, data, n_samples, stride, stream, promo);
        case 4: return launch_tma_cfg<T, M, KIND, 2, 2, 4, 1, 1, 8, false>(b, data, n_samples, stride, stream, promo);
        case 5: return launch_tma_cfg<T, M, KIND, 2, 2, 4, 2, 1, 8, false>(b, data, n_samples, stride, stream, promo);
        case 6: return launch_tma_cfg<T, M, KIND, 2, 2, 6, 2, 1, 8, false>(b, data, n_samples, stride, stream, promo);
        case 7: return launch_tma_cfg<T, M, KIND, 4, 2, 2, 1, 1, 8, false>(b, data, n_samples, stride, stream, promo);
        default: break;
        }
    }
    if (!pack)
        return launch_tma_cfg<T, M, KIND, 4, 2, 3, 2, 1, 8, false>(b, data, n_samples, stride, stream, promo);
    return launch_tma_cfg<T, M, KIND, 2, 2, 6, 3, 1, 8, true>(b, data, n_samples, stride, stream, promo);
}

template <typename T, int M>
static int launch_tma_kind(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    if constexpr (sizeof(T) == 4) { // fp32 only: +3.8 % on config 3; in fp64 an add costs the FP64 pipe what an fma does and the kernel got slower
        if (b.numerator == NUM_GENERIC && b.b2_all_one)
            return launch_tma<T, M, NUM_GENERIC_B2ONE>(b, data, n_samples, stride, stream);
    }
    switch (b.numerator) {
    case NUM_GENERIC: return launch_tma<T, M, NUM_GENERIC>(b, data, n_samples, stride, stream);
    case NUM_LP: return launch_tma<T, M, NUM_LP>(b, data, n_samples, stride, stream);
    case NUM_HP: return launch_tma<T, M, NUM_HP>(b, data, n_samples, stride, stream);
    default: return launch_tma<T, M, NUM_BP>(b, data, n_samples, stride, stream);
    }
}

template <typename T>
static int launch_tma_sections(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    switch (b.sections) {
    case 2: return launch_tma_kind<T, 2>(b, data, n_samples, stride, stream);
    case 4: return launch_tma_kind<T, 4>(b, data, n_samples, stride, stream);
    case 6: return launch_tma_kind<T, 6>(b, data, n_samples, stride, stream);
    case 8: return launch_tma_kind<T, 8>(b, data, n_samples, stride, stream);
    default: return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir tma path: sections=%d not built", b.sections);
    }
}

// ---- the entry points of this translation unit (one per precision: iir_tma_f32.cu / iir_tma_f64.cu)
#define SDSP_TMA_CAT2(a, b) a##b
#define SDSP_TMA_CAT(a, b) SDSP_TMA_CAT2(a, b)
#define SDSP_TMA_FN(name) SDSP_TMA_CAT(name, SDSP_TMA_SUFFIX)

int SDSP_TMA_FN(iir_launch_tma_)(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    return launch_tma_sections<SDSP_TMA_TYPE>(b, data, n_samples, stride, stream);
}
int SDSP_TMA_FN(iir_launch_tma_rows_)(const IirBank &b, void *data, size_t seg_len, size_t segs, size_t stride, void *row_state,
                                      size_t n_samples, bool accumulate, cudaStream_t stream, int *slots)
{
    return launch_rows_sections<SDSP_TMA_TYPE>(b, data, seg_len, segs, stride, row_state, n_samples, accumulate, stream, slots);
}
} // namespace sdsp_b200
