// iir_tma.cu -- the fast sequential IIR path: TMA-fed, one warp per 32 channels, sections software-skewed.
//
// Replaces casc_2o_iir<m_t>::process (reference include/sdsp/casc_2o_iir.h:36-80) for large banks.
//
// Data layout in HBM is the reference's: one contiguous range per channel, data[channel*stride + n].
// A lane-per-channel kernel therefore wants a [32 channels x TS samples] patch transposed on chip.  The
// TMA engine does that for free: a 2-D tensor map over (samples, channels) with a box of 32 rows x 128
// bytes and SWIZZLE_128B lands each channel's 128-byte run in its own shared-memory row, XOR-swizzled so
// that lane r reading 16-byte chunk c of row r (LDS.128 at r*128 + ((c ^ (r&7))<<4)) is conflict free.
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
template <typename T, int M, int KIND, int SUB, int CSUB, int NST, int PF, int WARPS, int RG>
__global__ void __launch_bounds__(WARPS * 32)
    iir_tma_kernel(const __grid_constant__ CUtensorMap map, int n_samples, const T *__restrict__ coef, T *__restrict__ state,
                   size_t n_channels)
{
    constexpr int TSB = 128 / (int)sizeof(T);
    constexpr int TS = TSB * SUB;   // samples per stage (what one mbarrier phase delivers)
    constexpr int CTS = TSB * CSUB; // samples per skewed compute tile
    static_assert(SUB % CSUB == 0 && (RG == 32 || RG == 8), "stage = whole compute tiles; boxes of 32 or 8 rows");
    constexpr int BOX_BYTES = 32 * 128;
    constexpr int STAGE_BYTES = BOX_BYTES * SUB;
    constexpr int VN = Vec16<T>::N;
    using V = typename Vec16<T>::type;
    static_assert(PF >= 1 && PF < NST, "prefetch distance must leave room for stores in flight");

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
        c.gain = coef[ch];
#pragma unroll
        for (int j = 0; j < M; j++) {
            c.b1[j] = coef[(size_t)(1 + j) * n_channels + ch];
            c.b2[j] = coef[(size_t)(1 + M + j) * n_channels + ch];
            c.na1[j] = coef[(size_t)(1 + 2 * M + j) * n_channels + ch];
            c.na2[j] = coef[(size_t)(1 + 3 * M + j) * n_channels + ch];
        }
#pragma unroll
        for (int r = 0; r <= M; r++) {
            st.h[r][0] = state[(size_t)(2 * r) * n_channels + ch];
            st.h[r][1] = state[(size_t)(2 * r + 1) * n_channels + ch];
        }
    } else {
        c.gain = 0;
#pragma unroll
        for (int j = 0; j < M; j++)
            c.b1[j] = c.b2[j] = c.na1[j] = c.na2[j] = 0;
#pragma unroll
        for (int r = 0; r <= M; r++)
            st.h[r][0] = st.h[r][1] = 0;
    }

    const int n_stages = (n_samples + TS - 1) / TS;
    const int y0 = (int)ch0;

    auto issue_load = [&](int k) { // lane 0 only
        uint64_t *b = &bar[k % NST];
        unsigned char *dst = ring + (size_t)(k % NST) * STAGE_BYTES;
        mbar_expect_tx(b, STAGE_BYTES);
        // RG = 8: boxes of 8 rows, the SUB boxes of one row group issued back to back, so that the requests
        // for consecutive 128-byte pieces of a channel reach the memory system together (DRAM page locality)
#pragma unroll
        for (int g = 0; g < 32 / RG; g++)
#pragma unroll
            for (int u = 0; u < SUB; u++)
                tma_load_2d(dst + u * BOX_BYTES + g * RG * 128, &map, k * TS + u * TSB, y0 + g * RG, b);
    };

    if (lane == 0) {
        for (int k = 0; k < PF && k < n_stages; k++)
            issue_load(k);
    }

    // this lane's row inside a box, and the XOR that un-swizzles 16-byte chunks
    const uint32_t row_off = (uint32_t)lane * 128u;
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
            unsigned char *cbuf = buf + ct * CSUB * BOX_BYTES;
            const int rem = remaining - ct * CTS;
            if (rem >= CTS) {
                V vin, vout;
                iir_tile_dispatch<T, M, KIND, CTS>(
                    c, st,
                    [&](int i) -> T {
                        if (i % VN == 0) {
                            const int box = i / TSB, chunk = (i % TSB) / VN;
                            vin = *reinterpret_cast<const V *>(cbuf + box * BOX_BYTES + row_off + (((uint32_t)chunk ^ sw) << 4));
                        }
                        return vget(vin, i % VN);
                    },
                    [&](int i, T y) {
                        vset(vout, i % VN, y);
                        if (i % VN == VN - 1) {
                            const int box = i / TSB, chunk = (i % TSB) / VN;
                            *reinterpret_cast<V *>(cbuf + box * BOX_BYTES + row_off + (((uint32_t)chunk ^ sw) << 4)) = vout;
                        }
                    });
            } else if (rem > 0) {
                // ragged tail: plain sample-by-sample order (same arithmetic, see iir_core.cuh)
                for (int i = 0; i < rem; i++) {
                    const int box = i / TSB, chunk = (i % TSB) / VN, e = i % VN;
                    T *p = reinterpret_cast<T *>(cbuf + box * BOX_BYTES + row_off + (((uint32_t)chunk ^ sw) << 4)) + e;
                    *p = iir_step<T, M, KIND>(*p, c, st);
                }
            }
        }
        fence_proxy_async(); // generic-proxy writes above -> visible to the TMA store below
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int g = 0; g < 32 / RG; g++)
#pragma unroll
                for (int u = 0; u < SUB; u++)
                    tma_store_2d(&map, k * TS + u * TSB, y0 + g * RG, buf + u * BOX_BYTES + g * RG * 128);
            tma_commit();
        }
    }
    if (lane == 0)
        tma_wait_all();
    if (active) {
#pragma unroll
        for (int r = 0; r <= M; r++) {
            state[(size_t)(2 * r) * n_channels + ch] = st.h[r][0];
            state[(size_t)(2 * r + 1) * n_channels + ch] = st.h[r][1];
        }
    }
}

// =================================================================================================
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

template <typename T, int M, int KIND, int SUB, int CSUB, int NST, int PF, int WARPS, int RG>
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
    auto kern = iir_tma_kernel<T, M, KIND, SUB, CSUB, NST, PF, WARPS, RG>;
    constexpr size_t smem = (size_t)WARPS * NST * SUB * 32 * 128;
    static bool configured = false;
    if (!configured) {
        SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const size_t groups = (b.n_channels + 31) / 32;
    const unsigned grid = (unsigned)((groups + WARPS - 1) / WARPS);
    kern<<<grid, WARPS * 32, smem, stream>>>(map, (int)n_samples, static_cast<const T *>(b.d_coef), static_cast<T *>(b.d_state), b.n_channels);
    SDSP_CUDA(cudaGetLastError());
    return SDSP_B200_OK;
}

// SDSP_B200_IIR_TUNE="<config>,<l2 promotion bytes>": kernel-tuning aid; the extra configurations exist only for
// the headline instantiation (fp32, 4 sections, generic numerator)
static void tma_tune(int &cfg, int &promo)
{
    static int c = -1, p = 128;
    if (c < 0) {
        c = 0;
        if (const char *e = getenv("SDSP_B200_IIR_TUNE"))
            sscanf(e, "%d,%d", &c, &p);
    }
    cfg = c;
    promo = p;
}

template <typename T, int M, int KIND>
static int launch_tma(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    int cfg, promo;
    tma_tune(cfg, promo);
    if constexpr (sizeof(T) == 4 && M == 4 && KIND == NUM_GENERIC) {
        switch (cfg) { //                                       SUB CSUB NST PF WARPS RG
        case 1: return launch_tma_cfg<T, M, KIND, 2, 2, 6, 3, 4, 8>(b, data, n_samples, stride, stream, promo);
        case 2: return launch_tma_cfg<T, M, KIND, 4, 2, 3, 2, 4, 8>(b, data, n_samples, stride, stream, promo);
        case 3: return launch_tma_cfg<T, M, KIND, 4, 2, 4, 2, 1, 8>(b, data, n_samples, stride, stream, promo);
        case 4: return launch_tma_cfg<T, M, KIND, 4, 2, 3, 2, 1, 8>(b, data, n_samples, stride, stream, promo);
        case 5: return launch_tma_cfg<T, M, KIND, 8, 2, 3, 2, 1, 8>(b, data, n_samples, stride, stream, promo);
        case 6: return launch_tma_cfg<T, M, KIND, 4, 2, 3, 2, 1, 32>(b, data, n_samples, stride, stream, promo);
        case 7: return launch_tma_cfg<T, M, KIND, 2, 2, 6, 3, 4, 32>(b, data, n_samples, stride, stream, promo);
        case 8: return launch_tma_cfg<T, M, KIND, 8, 2, 2, 1, 1, 8>(b, data, n_samples, stride, stream, promo);
        default: break;
        }
    }
    // default (measured best of the sweep in profiles/r01_iir_tma_config_sweep.txt): single-warp CTAs, 8-row boxes
    return launch_tma_cfg<T, M, KIND, 2, 2, 6, 3, 1, 8>(b, data, n_samples, stride, stream, promo);
}

template <typename T, int M>
static int launch_tma_kind(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
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

bool iir_tma_built_for(int sections)
{
    return sections == 2 || sections == 4 || sections == 6 || sections == 8;
}

int iir_launch_tma(const IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream)
{
    if (b.precision == SDSP_B200_F32)
        return launch_tma_sections<float>(b, data, n_samples, stride, stream);
    return launch_tma_sections<double>(b, data, n_samples, stride, stream);
}
} // namespace sdsp_b200
