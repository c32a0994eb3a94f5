// iir_scan_impl.cuh -- the look-back scan IIR path (chunked state-space scan), see iir_scan_core.cuh.
// Compiled once per precision (iir_scan_f32.cu / iir_scan_f64.cu).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "iir_internal.h"
#include "iir_scan_core.cuh"
#include "tma_ptx.cuh"

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
    iir_state_from_mem<T, M>(s, mem);
    std::vector<double> tab;
    int reach = 0;
    if (scan_build_tables(M, KIND, gain, b, a, L, scan_negligible<T>(), IirDelta<T>::value, tab, reach) != 0)
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "scan: sections=%d not built", M);
    T *d = static_cast<T *>(data);
    const size_t done = scan_emulate_channel<T, M, KIND>(c, s, tab, reach, L, d, n, force_general);
    for (size_t i = done; i < n; i++)
        d[i] = iir_step<T, M, KIND>(d[i], c, s);
    iir_state_to_mem<T, M>(s, mem);
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

// =================================================================================================
// device side
template <typename T, int SD>
struct alignas(16) ScanRec { // one per (channel, tile); written once per launch, tagged with the launch epoch
    T uh[2];     // scaled-input history leaving the tile (known as soon as the tile is loaded)
    T agg[SD];   // state leaving the tile if it had been entered with zero state
    T incl[SD];  // true state leaving the tile
    unsigned flag_x, flag_a, flag_i, pad;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// lane 0 polls, the warp re-converges; afterwards plain loads that bypass L1 see the published data
__device__ __forceinline__ void warp_wait_flag(const unsigned *flag, unsigned epoch, int lane)
{
    if (lane == 0) {
        while (ld_relaxed_u32(flag) != epoch)
            __nanosleep(20);
        (void)ld_acquire_u32(flag); // one acquire after the flag has been seen (orders the data reads below)
    }
    __syncwarp();
}
template <typename T>
__device__ __forceinline__ T ld_cg(const T *p)
{
    return __ldcg(p);
}

// L samples per lane, 32 lanes per tile, WARPS independent warps per CTA, RG rows per TMA box
// ONE_CH: the bank has a single channel, so one copy of its tables serves the whole CTA (more warps fit).
template <typename T, int M, int KIND, int L, int WARPS, int RG, bool ONE_CH>
__global__ void __launch_bounds__(WARPS * 32)
    iir_scan_kernel(const __grid_constant__ CUtensorMap map, const T *__restrict__ coef, const T *__restrict__ state, T *__restrict__ state_out,
                    size_t n_channels,
                    const T *__restrict__ tables, const int *__restrict__ reach_of, ScanRec<T, 2 * M> *__restrict__ recs,
                    unsigned *__restrict__ ticket, unsigned epoch, unsigned n_tiles, unsigned rows_per_channel)
{
    constexpr int SD = 2 * M;
    constexpr int TSB = 128 / (int)sizeof(T); // samples per 128-byte box row
    constexpr int NBOX = L / TSB;
    constexpr int CTS = 2 * TSB; // skewed compute tile
    static_assert(L % CTS == 0, "chunk must hold whole compute tiles");
    constexpr int BOX_BYTES = 32 * 128;
    constexpr int TILE_BYTES = NBOX * BOX_BYTES;
    constexpr int VN = Vec16<T>::N;
    constexpr int TAB = scan_table_count(M, L);
    using V = typename Vec16<T>::type;

    constexpr int TABLE_BYTES = (TAB * (int)sizeof(T) + 15) / 16 * 16;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bars[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *buf = smem_raw + (size_t)warp * TILE_BYTES;
    // this warp's copy of its channel's tables (H then the six matrices), refilled when the channel changes
    T *tab = reinterpret_cast<T *>(smem_raw + (size_t)WARPS * TILE_BYTES + (size_t)(ONE_CH ? 0 : warp) * TABLE_BYTES);
    unsigned tab_ch = 0xffffffffu;
    IirCoef<T, M> c;
    iir_zero_coef<T, M>(c);
    int reach = 0;
    uint64_t *bar = &bars[warp];
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    auto elem = [&](int row, int i) -> T * { // address of sample i of chunk `row` inside the swizzled tile
        const int box = i / TSB, chunk = (i % TSB) / VN, e = i % VN;
        return reinterpret_cast<T *>(buf + box * BOX_BYTES + row * 128 + ((((uint32_t)chunk) ^ (uint32_t)(row & 7)) << 4)) + e;
    };
    // coefficients into registers, tables into shared memory.  H is re-laid for the correction loop: fp32 keeps
    // sample pairs together, [i/2][k][2], so that one 16-byte load feeds two FFMA2; fp64 stays [i][k]
    auto load_channel = [&](unsigned ch, int first, int step) {
        iir_load_coef<T, M>(c, coef, n_channels, ch);
        reach = reach_of[ch];
        const T *src = tables + (size_t)ch * TAB;
        for (int idx = first; idx < L * SD; idx += step) {
            const int i = idx / SD, k = idx % SD;
            const int dst = sizeof(T) == 4 ? (((i >> 1) * SD + k) * 2 + (i & 1)) : idx;
            tab[dst] = __ldg(src + idx);
        }
        for (int idx = L * SD + first; idx < TAB; idx += step)
            tab[idx] = __ldg(src + idx);
    };
    if (ONE_CH) {
        load_channel(0, threadIdx.x, WARPS * 32);
        __syncthreads();
    }
    const unsigned total = (unsigned)n_channels * n_tiles;
    unsigned phase = 0;

    for (;;) {
        unsigned w = 0;
        if (lane == 0)
            w = atomicAdd(ticket, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total)
            break;
        const unsigned ch = w / n_tiles, t = w % n_tiles;
        const int row0 = (int)(ch * rows_per_channel + t * 32u);
        if (lane == 0) {
            mbar_expect_tx(bar, TILE_BYTES);
#pragma unroll
            for (int g = 0; g < 32 / RG; g++)
#pragma unroll
                for (int u = 0; u < NBOX; u++)
                    tma_load_2d(buf + u * BOX_BYTES + g * RG * 128, &map, u * TSB, row0 + g * RG, bar);
        }
        if (!ONE_CH && ch != tab_ch) { // while the tile is in flight: this channel's coefficients and tables
            load_channel(ch, lane, 32);
            tab_ch = ch;
            __syncwarp();
        }
        ScanRec<T, SD> *rec = recs + (size_t)ch * n_tiles + t;

        mbar_wait(bar, phase);
        phase ^= 1u;

        // ---- scaled-input history entering each chunk, read before anything is overwritten
        IirState<T, M> z;
        iir_zero_state<T, M>(z);
        if (lane > 0) {
            z.h[0][0] = *elem(lane - 1, L - 1) * c.gain;
            z.h[0][1] = *elem(lane - 1, L - 2) * c.gain;
        }
        const T out_u1 = *elem(31, L - 1) * c.gain, out_u2 = *elem(31, L - 2) * c.gain; // leaves the tile
        if (lane == 31) {
            rec->uh[0] = out_u1;
            rec->uh[1] = out_u2;
            st_release_u32(&rec->flag_x, epoch);
        }
        if (t == 0) {
            if (lane == 0) {
                z.h[0][0] = state[(size_t)0 * n_channels + ch];
                z.h[0][1] = state[(size_t)1 * n_channels + ch];
            }
        } else {
            warp_wait_flag(&rec[-1].flag_x, epoch, lane);
            if (lane == 0) {
                z.h[0][0] = ld_cg(&rec[-1].uh[0]);
                z.h[0][1] = ld_cg(&rec[-1].uh[1]);
            }
        }
        __syncwarp();

        // ---- zero-state pass over this lane's chunk, in place
#pragma unroll 1
        for (int ct = 0; ct < L / CTS; ct++) {
            unsigned char *cbuf = buf + ct * 2 * BOX_BYTES;
            V vin, vout;
            iir_tile_dispatch<T, M, KIND, CTS>(
                c, z,
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
        }
        T P[SD];
        scan_state_to_vec<T, M>(z, P);

        // ---- Kogge-Stone over the 32 chunks: P = state leaving chunk `lane` had the tile been entered with zero state
#pragma unroll
        for (int j = 0; j < SCAN_KS_STEPS; j++) {
            T recv[SD];
#pragma unroll
            for (int k = 0; k < SD; k++)
                recv[k] = __shfl_up_sync(0xffffffffu, P[k], 1 << j);
            if (lane >= (1 << j))
                tri_matvec_acc<T, SD>(tab + scan_off_A(M, L, j), recv, P);
        }
        T agg[SD], Pprev[SD];
#pragma unroll
        for (int k = 0; k < SD; k++) {
            agg[k] = __shfl_sync(0xffffffffu, P[k], 31);
            Pprev[k] = __shfl_up_sync(0xffffffffu, P[k], 1);
            if (lane == 0)
                Pprev[k] = 0;
        }
        if (lane == 31) {
#pragma unroll
            for (int k = 0; k < SD; k++)
                rec->agg[k] = agg[k];
            st_release_u32(&rec->flag_a, epoch);
        }

        // ---- state entering the tile (identical in every lane)
        const T *Mt = tab + scan_off_A(M, L, SCAN_KS_STEPS);
        // the bank's section history as a carried vector: (v1, v2) rows 2 + k; delta form (v1, d): d in rows 2(M+1) + j
        auto bank_vec = [&](int k) -> T {
            const int row = (IirDelta<T>::value && (k & 1)) ? 2 * (M + 1) + k / 2 : 2 + k;
            return state[(size_t)row * n_channels + ch];
        };
        T cin[SD];
        if (t == 0) {
#pragma unroll
            for (int k = 0; k < SD; k++)
                cin[k] = bank_vec(k);
        } else if (reach <= SCAN_MAX_REACH) {
            const unsigned K = (unsigned)reach < t ? (unsigned)reach : t;
#pragma unroll
            for (int k = 0; k < SD; k++)
                cin[k] = (K == t) ? bank_vec(k) : (T)0;
            for (unsigned kk = K; kk >= 1; kk--) {
                const ScanRec<T, SD> *pr = rec - kk;
                warp_wait_flag(&pr->flag_a, epoch, lane);
                T nxt[SD];
#pragma unroll
                for (int k = 0; k < SD; k++)
                    nxt[k] = ld_cg(&pr->agg[k]);
                tri_matvec_acc<T, SD>(Mt, cin, nxt);
#pragma unroll
                for (int k = 0; k < SD; k++)
                    cin[k] = nxt[k];
            }
        } else {
            warp_wait_flag(&rec[-1].flag_i, epoch, lane);
#pragma unroll
            for (int k = 0; k < SD; k++)
                cin[k] = ld_cg(&rec[-1].incl[k]);
        }
        T incl[SD];
#pragma unroll
        for (int k = 0; k < SD; k++)
            incl[k] = agg[k];
        tri_matvec_acc<T, SD>(Mt, cin, incl);
        if (lane == 31) {
            if (reach > SCAN_MAX_REACH) { // only the wait-for-predecessor path reads inclusive states
#pragma unroll
                for (int k = 0; k < SD; k++)
                    rec->incl[k] = incl[k];
                st_release_u32(&rec->flag_i, epoch);
            }
            if (t + 1 == n_tiles) { // the bank's history after the last whole tile.  It goes to a second buffer: the
                                    // first tiles of this launch may not have read the incoming history yet
                IirState<T, M> fin;
                fin.h[0][0] = out_u1;
                fin.h[0][1] = out_u2;
                scan_vec_to_state<T, M>(incl, fin);
                iir_store_state<T, M>(fin, state_out, n_channels, ch);
            }
        }

        // ---- A_L^lane * cin by the binary digits of the lane index, plus the zero-state prefix
        T ci[SD];
#pragma unroll
        for (int k = 0; k < SD; k++)
            ci[k] = cin[k];
#pragma unroll
        for (int j = 0; j < SCAN_KS_STEPS; j++) {
            T nq[SD];
#pragma unroll
            for (int k = 0; k < SD; k++)
                nq[k] = 0;
            tri_matvec_acc<T, SD>(tab + scan_off_A(M, L, j), ci, nq);
            if (lane & (1 << j)) {
#pragma unroll
                for (int k = 0; k < SD; k++)
                    ci[k] = nq[k];
            }
        }
#pragma unroll
        for (int k = 0; k < SD; k++)
            ci[k] += Pprev[k];

        // ---- natural-response correction of this lane's chunk: y[i] += sum_k H[i][k] ci[k]
        if constexpr (sizeof(T) == 4) {
            f32x2 cib[SD];
#pragma unroll
            for (int k = 0; k < SD; k++)
                cib[k] = mk2(ci[k], ci[k]);
#pragma unroll 2
            for (int q = 0; q < L / 4; q++) {
                const int box = (q * 4) / TSB, chunk = ((q * 4) % TSB) / 4;
                float4 *p = reinterpret_cast<float4 *>(buf + box * BOX_BYTES + row_off + (((uint32_t)chunk ^ sw) << 4));
                float4 v = *p;
                f32x2 y01 = mk2(v.x, v.y), y23 = mk2(v.z, v.w);
                const float4 *h01 = reinterpret_cast<const float4 *>(tab + (size_t)(2 * q) * SD * 2);
                const float4 *h23 = h01 + SD / 2;
#pragma unroll
                for (int kk = 0; kk < SD / 2; kk++) {
                    const float4 ha = h01[kk], hb = h23[kk];
                    y01 = fma2(mk2(ha.x, ha.y), cib[2 * kk], y01);
                    y23 = fma2(mk2(hb.x, hb.y), cib[2 * kk], y23);
                    y01 = fma2(mk2(ha.z, ha.w), cib[2 * kk + 1], y01);
                    y23 = fma2(mk2(hb.z, hb.w), cib[2 * kk + 1], y23);
                }
                *p = make_float4(y01.x, y01.y, y23.x, y23.y);
            }
        } else {
#pragma unroll 4
            for (int q = 0; q < L / 2; q++) {
                const int box = (q * 2) / TSB, chunk = ((q * 2) % TSB) / 2;
                double2 *p = reinterpret_cast<double2 *>(buf + box * BOX_BYTES + row_off + (((uint32_t)chunk ^ sw) << 4));
                double2 v = *p;
                const double2 *h0 = reinterpret_cast<const double2 *>(tab + (size_t)(2 * q) * SD);
                const double2 *h1 = h0 + SD / 2;
#pragma unroll
                for (int kk = 0; kk < SD / 2; kk++) {
                    const double2 ha = h0[kk], hb = h1[kk];
                    v.x = fma_t((T)ha.x, ci[2 * kk], (T)v.x);
                    v.y = fma_t((T)hb.x, ci[2 * kk], (T)v.y);
                    v.x = fma_t((T)ha.y, ci[2 * kk + 1], (T)v.x);
                    v.y = fma_t((T)hb.y, ci[2 * kk + 1], (T)v.y);
                }
                *p = v;
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int g = 0; g < 32 / RG; g++)
#pragma unroll
                for (int u = 0; u < NBOX; u++)
                    tma_store_2d(&map, u * TSB, row0 + g * RG, buf + u * BOX_BYTES + g * RG * 128);
            tma_commit();
            tma_wait_read<0>(); // single buffer: the next tile's load must not overtake this store's reads
        }
        __syncwarp();
    }
    if (lane == 0)
        tma_wait_all();
}

// =================================================================================================
// host side
template <typename T>
struct ScanChunk; // samples per lane: 32 KiB tiles
template <>
struct ScanChunk<float> {
    static constexpr int L = 128;
};
template <>
struct ScanChunk<double> {
    static constexpr int L = 64;
};

template <typename T, int M, int KIND, int L, int WARPS, bool ONE_CH>
static int launch_scan_cfg(IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream, size_t *done)
{
    constexpr int SD = 2 * M, RG = 8;
    constexpr int TSB = 128 / (int)sizeof(T);
    constexpr int TAB = scan_table_count(M, L);
    const size_t tile = (size_t)32 * L;
    const size_t n_tiles = n_samples / tile;
    *done = 0;
    if (n_tiles == 0)
        return SDSP_B200_OK;
    if (n_tiles * b.n_channels >= (1ull << 31))
        return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir scan: too many tiles");
    if (b.h_gain.size() != b.n_channels)
        return set_error(SDSP_B200_ERR_INVALID_ARG, "iir scan: coefficients have not been set");

    // per-channel propagation tables, rebuilt when the coefficients change
    if (b.scan_tables_version != b.coef_version || b.scan_chunk != L) {
        std::vector<T> tabs((size_t)b.n_channels * TAB);
        std::vector<int> reach(b.n_channels);
        std::vector<double> one;
        for (size_t ch = 0; ch < b.n_channels; ch++) {
            int r = 0;
            scan_build_tables(M, KIND, b.h_gain[ch], &b.h_b[ch * 3 * M], &b.h_a[ch * 3 * M], L, scan_negligible<T>(), IirDelta<T>::value, one, r);
            reach[ch] = r;
            for (int i = 0; i < TAB; i++)
                tabs[ch * TAB + i] = (T)one[i];
        }
        const size_t bytes = tabs.size() * sizeof(T) + reach.size() * sizeof(int);
        if (b.scan_tables_bytes < bytes) {
            if (b.d_scan_tables)
                cudaFree(b.d_scan_tables);
            b.d_scan_tables = nullptr;
            b.scan_tables_bytes = 0;
            if (cudaMalloc(&b.d_scan_tables, bytes) != cudaSuccess) {
                cudaGetLastError();
                return set_error(SDSP_B200_ERR_OOM, "iir scan: cannot allocate %zu bytes of tables", bytes);
            }
            b.scan_tables_bytes = bytes;
        }
        SDSP_CUDA(cudaMemcpyAsync(b.d_scan_tables, tabs.data(), tabs.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        SDSP_CUDA(cudaMemcpyAsync(static_cast<char *>(b.d_scan_tables) + tabs.size() * sizeof(T), reach.data(), reach.size() * sizeof(int),
                                  cudaMemcpyHostToDevice, stream));
        SDSP_CUDA(cudaStreamSynchronize(stream)); // host vectors go out of scope
        b.scan_tables_version = b.coef_version;
        b.scan_chunk = L;
        b.scan_reach_max = 0;
        for (int r : reach)
            b.scan_reach_max = r > b.scan_reach_max ? r : b.scan_reach_max;
    }
    // carry records + ticket
    const size_t rec_bytes = sizeof(ScanRec<T, SD>) * b.n_channels * n_tiles + 256;
    if (b.scan_flags_bytes < rec_bytes) {
        if (b.d_scan_flags)
            cudaFree(b.d_scan_flags);
        b.d_scan_flags = nullptr;
        b.scan_flags_bytes = 0;
        if (cudaMalloc(&b.d_scan_flags, rec_bytes) != cudaSuccess) {
            cudaGetLastError();
            return set_error(SDSP_B200_ERR_OOM, "iir scan: cannot allocate %zu bytes of carry records", rec_bytes);
        }
        b.scan_flags_bytes = rec_bytes;
        SDSP_CUDA(cudaMemsetAsync(b.d_scan_flags, 0, rec_bytes, stream));
        b.scan_epoch = 0;
    }
    b.scan_epoch++;
    unsigned *ticket = static_cast<unsigned *>(b.d_scan_flags);
    auto *recs = reinterpret_cast<ScanRec<T, SD> *>(static_cast<char *>(b.d_scan_flags) + 256);
    SDSP_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), stream));

    // 2-D view: rows of L samples; channel c starts at row c * (stride / L)
    const size_t rows_per_channel = b.n_channels > 1 ? stride / L : n_tiles * 32;
    const size_t total_rows = (b.n_channels - 1) * rows_per_channel + n_tiles * 32;
    CUtensorMap map;
    const cuuint64_t gdim[2] = { (cuuint64_t)L, (cuuint64_t)total_rows };
    const cuuint64_t gstride[1] = { (cuuint64_t)L * sizeof(T) };
    const cuuint32_t box[2] = { (cuuint32_t)TSB, (cuuint32_t)RG };
    const cuuint32_t estr[2] = { 1, 1 };
    CUresult r = get_encode_fn()(&map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, data, gdim,
                                 gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(SDSP_B200_ERR_CUDA, "iir scan: cuTensorMapEncodeTiled failed with %d", (int)r);

    auto kern = iir_scan_kernel<T, M, KIND, L, WARPS, RG, ONE_CH>;
    constexpr size_t smem = (size_t)WARPS * (size_t)(L / TSB) * 32 * 128 + (size_t)(ONE_CH ? 1 : WARPS) * (((size_t)TAB * sizeof(T) + 15) / 16 * 16);
    static bool configured_dev[64] = {};
    static int occ_dev[64] = {};
    bool &configured = configured_dev[b.device & 63]; // (the attribute is per device; a process may hold banks on several)
    int &occ = occ_dev[b.device & 63];
    if (!configured) {
        SDSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SDSP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
        if (occ < 1)
            occ = 1;
        configured = true;
    }
    const size_t total_tiles = b.n_channels * n_tiles;
    size_t grid = (size_t)b.sm_count * occ;
    if (grid * WARPS > total_tiles)
        grid = (total_tiles + WARPS - 1) / WARPS;
    // outgoing history is written to the bank's second state buffer, which then becomes the current one
    const size_t state_bytes = (size_t)iir_bank_state_rows(b) * b.n_channels * sizeof(T);
    if (!b.d_state_alt) {
        if (cudaMalloc(&b.d_state_alt, state_bytes) != cudaSuccess) {
            cudaGetLastError();
            return set_error(SDSP_B200_ERR_OOM, "iir scan: cannot allocate %zu bytes of state", state_bytes);
        }
    }
    kern<<<(unsigned)grid, WARPS * 32, smem, stream>>>(map, static_cast<const T *>(b.d_coef), static_cast<const T *>(b.d_state),
                                                      static_cast<T *>(b.d_state_alt), b.n_channels,
                                                      static_cast<const T *>(b.d_scan_tables),
                                                      reinterpret_cast<const int *>(static_cast<char *>(b.d_scan_tables) +
                                                                                    (size_t)b.n_channels * TAB * sizeof(T)),
                                                      recs, ticket, b.scan_epoch, (unsigned)n_tiles, (unsigned)rows_per_channel);
    SDSP_CUDA(cudaGetLastError());
    std::swap(b.d_state, b.d_state_alt);
    *done = n_tiles * tile;
    return SDSP_B200_OK;
}

// chunk length: ScanChunk<T>::L (16 KiB tiles); warps per CTA = what 227 KB of shared memory holds next to the table copies
template <typename T, int M, int KIND>
static int launch_scan(IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream, size_t *done)
{
    constexpr int L = ScanChunk<T>::L;
    constexpr int TABB = (scan_table_count(M, L) * (int)sizeof(T) + 15) / 16 * 16;
    constexpr int W_ONE = (220 * 1024 - TABB) / (16 * 1024) > 12 ? 12 : (220 * 1024 - TABB) / (16 * 1024);
    constexpr int W_MANY = (220 * 1024) / (16 * 1024 + TABB) > 12 ? 12 : (220 * 1024) / (16 * 1024 + TABB);
    if (b.n_channels == 1)
        return launch_scan_cfg<T, M, KIND, L, W_ONE, true>(b, data, n_samples, stride, stream, done);
    return launch_scan_cfg<T, M, KIND, L, W_MANY, false>(b, data, n_samples, stride, stream, done);
}

template <typename T, int M>
static int launch_scan_kind(IirBank &b, void *data, size_t n, size_t stride, cudaStream_t s, size_t *done)
{
    switch (b.numerator) {
    case NUM_GENERIC: return launch_scan<T, M, NUM_GENERIC>(b, data, n, stride, s, done);
    case NUM_LP: return launch_scan<T, M, NUM_LP>(b, data, n, stride, s, done);
    case NUM_HP: return launch_scan<T, M, NUM_HP>(b, data, n, stride, s, done);
    default: return launch_scan<T, M, NUM_BP>(b, data, n, stride, s, done);
    }
}

template <typename T>
static int launch_scan_sections(IirBank &b, void *data, size_t n, size_t stride, cudaStream_t s, size_t *done)
{
    switch (b.sections) {
    case 2: return launch_scan_kind<T, 2>(b, data, n, stride, s, done);
    case 4: return launch_scan_kind<T, 4>(b, data, n, stride, s, done);
    case 6: return launch_scan_kind<T, 6>(b, data, n, stride, s, done);
    case 8: return launch_scan_kind<T, 8>(b, data, n, stride, s, done);
    default: return set_error(SDSP_B200_ERR_UNSUPPORTED, "iir scan path: sections=%d not built (2, 4, 6, 8)", b.sections);
    }
}

// ---- the entry points of this translation unit (one per precision: iir_scan_f32.cu / iir_scan_f64.cu)
#define SDSP_SCAN_CAT2(a, b) a##b
#define SDSP_SCAN_CAT(a, b) SDSP_SCAN_CAT2(a, b)
#define SDSP_SCAN_FN(name) SDSP_SCAN_CAT(name, SDSP_SCAN_SUFFIX)

int SDSP_SCAN_FN(iir_scan_chunk_)()
{
    return ScanChunk<SDSP_SCAN_TYPE>::L;
}
int SDSP_SCAN_FN(iir_launch_scan_tiles_)(IirBank &b, void *data, size_t n_samples, size_t stride, cudaStream_t stream, size_t *done)
{
    return launch_scan_sections<SDSP_SCAN_TYPE>(b, data, n_samples, stride, stream, done);
}
int SDSP_SCAN_FN(iir_emulate_scan_)(int sections, int numerator, double gain, const double *b, const double *a, double *mem, void *data,
                                    size_t n_samples, int chunk, bool force_general)
{
    return emulate_scan_sections<SDSP_SCAN_TYPE>(sections, numerator, gain, b, a, mem, data, n_samples, chunk, force_general);
}
} // namespace sdsp_b200
