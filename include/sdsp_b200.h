/*
 * sdsp_b200.h -- C ABI of libsdsp_b200.so: the B200 (sm_100a) implementation of simpledsp's two
 * data-parallel hot paths, batched.
 *
 * simpledsp itself has no FFI: its hot path is a header-only C++ template API
 * (namespace sdsp, include/sdsp/fft.h and include/sdsp/casc_2o_iir.h in the reference).  This file is
 * the one process/device boundary the B200 build introduces.  The drop-in headers shipped beside it
 * (include/sdsp/fft.h, include/sdsp/casc_2o_iir.h, include/sdsp/filter_type.h) keep the reference's
 * names and signatures and forward to the entry points below; any other language binds the same
 * symbols (see INTEGRATION.md for the ctypes / cgo / JNI stubs).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns an int status (0 = ok) and records a
 *     message retrievable with sdsp_b200_last_error() (thread local).
 *   - complex data is interleaved (re, im), exactly std::complex<T>[n] / sdsp::complex_array<N>
 *     (reference include/sdsp/fft.h:51-52); frames are contiguous: data[frame][n].
 *   - IIR data is planar channel-major: data[channel * channel_stride + sample]; one contiguous
 *     range per filter object, as in reference casc_2o_iir.h:36-80 (process(begin, end)).
 *   - all transforms and filters run IN PLACE (reference fft.h:290-291, 342-345; casc_2o_iir.h:71).
 *   - ptr_kind says whether `data` is a host pointer (the library stages it through device memory,
 *     synchronously) or a device pointer on the handle's device (asynchronous on `stream`).
 *   - `stream` is a cudaStream_t passed as void*; NULL = the legacy default stream.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *     SDSP_B200_ERR_NO_DEVICE / SDSP_B200_ERR_CUDA.
 */
#ifndef SDSP_B200_H
#define SDSP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDSP_B200_VERSION 200 /* 0.2.0 */
/* largest frame sdsp_b200_fft_plan_create accepts (the reference takes any power of two its compiler can fold tables for) */
#define SDSP_B200_FFT_MAX_N_F32 (1u << 18)
#define SDSP_B200_FFT_MAX_N_F64 (1u << 17)
/* section counts: a device-resident bank holds 1..8 sections per channel; the single-object entry points (designers, preload,
 * sdsp_b200_iir_process_once) chain groups of eight and take up to this many */
#define SDSP_B200_IIR_MAX_SECTIONS_ONCE 64

enum sdsp_b200_status {
    SDSP_B200_OK = 0,
    SDSP_B200_ERR_INVALID_ARG = 1,
    SDSP_B200_ERR_UNSUPPORTED = 2, /* size / radix / section count outside what is built */
    SDSP_B200_ERR_CUDA = 3,
    SDSP_B200_ERR_OOM = 4,
    SDSP_B200_ERR_NO_DEVICE = 5
};

enum sdsp_b200_precision { SDSP_B200_F32 = 0, SDSP_B200_F64 = 1 };
/* reference fft.h:121-146: forward_fft (Sign +1, no scaling) / reverse_fft (Sign -1, times 1.0/N) */
enum sdsp_b200_direction { SDSP_B200_FORWARD = 0, SDSP_B200_REVERSE = 1 };
enum sdsp_b200_ptr_kind { SDSP_B200_PTR_HOST = 0, SDSP_B200_PTR_DEVICE = 1 };
/* reference filter_type.h:6 -- same values */
enum sdsp_b200_filter_type { SDSP_B200_FILTER_NONE = 0, SDSP_B200_LOW_PASS = 1, SDSP_B200_HIGH_PASS = 2, SDSP_B200_BAND_PASS = 3 };
/* which class of the reference a bank mirrors: casc_2o_iir (runtime numerator, casc_2o_iir.h:8-215)
 * or casc_2o_iir_lp / _hp / _bp (numerators {1,2,1} {1,-2,1} {1,0,-1} hard-wired, :266-468) */
enum sdsp_b200_numerator { SDSP_B200_NUM_GENERIC = 0, SDSP_B200_NUM_LP = 1, SDSP_B200_NUM_HP = 2, SDSP_B200_NUM_BP = 3 };
/* how a bank walks the time axis */
enum sdsp_b200_iir_path {
    SDSP_B200_IIR_AUTO = 0,       /* a sequential kernel (bit-identical however a stream is cut into calls), EXCEPT for banks too small to
                                     fill the GPU on calls far longer than block streaming uses: ceil(channels/32) <= 2 x SM count (9472
                                     channels on a B200) AND n_samples >= 65536 AND the time-split path applies -> time-split (reassociates
                                     within the parity tolerance).  SDSP_B200_IIR_SEQUENTIAL, or SDSP_B200_IIR_AUTO_SPLIT=0 in the
                                     environment, keeps such calls on the sequential kernels. */
    SDSP_B200_IIR_SEQUENTIAL = 1, /* lane per channel, samples in order; bit-identical however a stream is cut into calls */
    SDSP_B200_IIR_SCAN = 2,       /* time axis split into chunks that carry boundary state (reassociates; fp64 error ~1e-13 of peak):
                                     the time-split kernel when the filter's memory fits a segment, else the look-back scan; where neither
                                     applies (call too short for the filter, unsuitable layout) the sequential kernels */
    SDSP_B200_IIR_SCAN_LOOKBACK = 3, /* force the look-back scan kernel (any stable filter) */
    SDSP_B200_IIR_SCAN_SPLIT = 4     /* force the time-split kernel (error if the filter's memory is too long for the call) */
};

typedef struct sdsp_b200_fft_plan_s *sdsp_b200_fft_plan;
typedef struct sdsp_b200_iir_bank_s *sdsp_b200_iir_bank;

/* ---------------------------------------------------------------- runtime */
int sdsp_b200_version(void);
const char *sdsp_b200_last_error(void);
int sdsp_b200_device_count(int *count);
/* creates the context on `device` and checks it is sm_100; optional (plans do it lazily) */
int sdsp_b200_init(int device);
int sdsp_b200_shutdown(void);
/* pinned host buffers for callers that want full-speed staging of host data */
int sdsp_b200_host_alloc(void **ptr, size_t bytes);
int sdsp_b200_host_free(void *ptr);
int sdsp_b200_device_alloc(void **ptr, size_t bytes, int device);
int sdsp_b200_device_free(void *ptr, int device);
int sdsp_b200_memcpy(void *dst, const void *src, size_t bytes, int device); /* any direction, synchronous */
int sdsp_b200_device_synchronize(int device);

/* Multi-GPU sharding (host only, no device needed).  Frames and channels are independent objects in the reference (one
 * complex_array per call, fft.h:258-360; one casc_2o_iir object per channel, casc_2o_iir.h:8-20), so device `rank` of `world`
 * owns the contiguous block [*first, *first + *count) of `total` units -- the first total % world ranks hold one unit more --
 * together with its share of the coefficient / history bank; there is no collective on the data path.  A process drives
 * several GPUs by creating one plan / bank per device (the `device` argument) over these ranges. */
int sdsp_b200_shard_range(size_t total, int rank, int world, size_t *first, size_t *count);

/* ---------------------------------------------------------------- FFT
 * Replaces sdsp::fft_radix2<T,N>(complex_array<N>&) (reference include/sdsp/fft.h:258-299) and
 * sdsp::fft_radix4<T,N>(complex_array<N>&) (:301-360), batched over n_frames frames.
 *   radix 2: n must be a power of 2 (static_assert at fft.h:261)
 *   radix 4: n must be a power of 4 (static_assert at fft.h:304)
 * Both produce the same transform: natural-order in, natural-order out, unnormalised forward,
 * 1/N-scaled reverse.  radix selects the argument check, not the arithmetic: on the device the
 * frame is factored into register-resident radix-16/8/4/2 passes. */
int sdsp_b200_fft_plan_create(sdsp_b200_fft_plan *plan, uint32_t n, int radix, int precision, int direction, int device);
int sdsp_b200_fft_plan_destroy(sdsp_b200_fft_plan plan);
/* Device-pointer calls on one plan must be ordered with respect to each other (plans for n >= 32768 own scratch memory);
 * use one plan per concurrent stream. */
int sdsp_b200_fft_exec(sdsp_b200_fft_plan plan, void *data, size_t n_frames, int ptr_kind, void *stream);
/* Real frames in (n scalars each), spectra out (n complex each), out of place.  The reference has no real-input
 * entry point; its callers place real signals in the real part of a complex_array and leave the imaginary part
 * zero (test/testFFT.cpp:24, :86).  This does that placement inside the first load of the transform, so the input
 * costs half the bytes. */
int sdsp_b200_fft_exec_real(sdsp_b200_fft_plan plan, const void *real_in, void *spectrum_out, size_t n_frames, int ptr_kind, void *stream);
/* Real frames in (n scalars each), HALF spectra out: the bins 0 .. n/2 of each frame (n/2 + 1 complex values, frames n/2 + 1
 * values apart), out of place, forward plans only.  The bins above n/2 are the conjugates of those below, X[n - k] = conj X[k],
 * so nothing is lost against the reference's convention (real part filled, imaginary part zero, full spectrum back:
 * test/testFFT.cpp:24, :86) while the call moves 4 + 4 bytes per sample instead of 8 + 8 (fp32).  On the device the frame
 * is read as n/2 complex numbers, transformed by the n/2-point kernel and separated in one more exchange
 * (n = 4 .. 65536 in f32, 4 .. 16384 in f64; larger frames go through the plan's full-length transform and a temporary of full
 * spectra: complete, not fast).  Device buffers aligned to one complex element. */
int sdsp_b200_fft_exec_r2c(sdsp_b200_fft_plan plan, const void *real_in, void *half_spectrum_out, size_t n_frames, int ptr_kind, void *stream);
/* The way back, for REVERSE plans: half spectra in (n/2 + 1 bins per frame), real frames out (n scalars each), 1/n included as in
 * reverse_fft::ScaleValues (fft.h:128-132); the imaginary parts of bins 0 and n/2 are ignored, as the mirror symmetry demands.
 * exec_c2r(exec_r2c(x)) == x to rounding.  Direct kernel for n = 4 .. 32768 (f32) / 4 .. 16384 (f64); larger frames through the
 * plan's full-length transform and a temporary. */
int sdsp_b200_fft_exec_c2r(sdsp_b200_fft_plan plan, const void *half_spectrum_in, void *real_out, size_t n_frames, int ptr_kind, void *stream);
/* human-readable description of the factorisation / launch geometry the plan chose */
int sdsp_b200_fft_plan_describe(sdsp_b200_fft_plan plan, char *buf, size_t buf_len);
/* number of kernel launches one exec of n_frames device-resident frames issues */
int sdsp_b200_fft_plan_launches(sdsp_b200_fft_plan plan, size_t n_frames, int *launches);

/* The reference's twiddle table calc_wCoeffs<N,T>() (fft.h:197-214): out[log2(n)][n][2] doubles,
 * W[i][j] = exp(-i * Sign * 2 pi j / 2^(i+1)).  Host only (no device needed); produced by the same
 * octant-symmetric long-double generator that fills the device tables. */
int sdsp_b200_twiddle_table(uint32_t n, int direction, double *out);

/* Digit-reversal tables, computed on the device (reference fft.h:217-236 digit_reverse<N,base>,
 * :238-256 calc_swap_lookup<N,base>).  base 2 or 4.  half_table = 0: out[i] = rev(i);
 * half_table = 1: the reference's swap table (the higher index of every pair maps to itself).
 * `out` is a host array of n uint32.  Parity requirement: bit-exact. */
int sdsp_b200_digit_reverse_table(uint32_t n, uint32_t base, int half_table, uint32_t *out, int device);
/* Apply out[rev(i)] = in[i] to n_frames frames in place on the device (the permutation stage of
 * fft.h:269-273 / 351-355 as a stand-alone operator). */
int sdsp_b200_digit_reverse_permute(void *data, uint32_t n, uint32_t base, int precision, size_t n_frames,
                                    int ptr_kind, int device, void *stream);

/* ---------------------------------------------------------------- cascaded biquad IIR
 * A bank is n_channels independent filter objects of one class and one section count
 * (reference casc_2o_iir<m_t>, casc_2o_iir.h:8-215).  Coefficients and history live on the device.
 * sections = m_t (1..8; the reference insists on even m_t, casc_2o_iir.h:25 -- the drop-in header keeps
 * that static_assert, the ABI does not need it). */
int sdsp_b200_iir_bank_create(sdsp_b200_iir_bank *bank, int sections, size_t n_channels, int precision,
                              int numerator, int device);
int sdsp_b200_iir_bank_destroy(sdsp_b200_iir_bank bank);
/* Upload coefficients for channels [first, first+count).  Host arrays of doubles laid out like the
 * reference's members: gain[count] (m_gain), b[count][sections][3] (m_b_coeff), a[count][sections][3]
 * (m_a_coeff).  b[..][0] and a[..][0] are ignored (b0 == 1 implicitly, a0 never read:
 * casc_2o_iir.h:64-69).  b may be NULL for the fixed-numerator banks.  Does not touch history
 * (= copy_coeff_from, casc_2o_iir.h:28-34). */
int sdsp_b200_iir_bank_set_coeffs(sdsp_b200_iir_bank bank, size_t first, size_t count, const double *gain,
                                  const double *b, const double *a);
/* History for channels [first, first+count): mem[count][sections+1][2] = for every row of the
 * reference's m_mem (row 0 = scaled input, row j = output of section j-1) the two most recent
 * values {x[n-1], x[n-2]}.  (The reference's 3-slot ring + m_pos, casc_2o_iir.h:11,15, is private;
 * only this information is observable.) */
int sdsp_b200_iir_bank_set_state(sdsp_b200_iir_bank bank, size_t first, size_t count, const double *mem);
int sdsp_b200_iir_bank_get_state(sdsp_b200_iir_bank bank, size_t first, size_t count, double *mem);
int sdsp_b200_iir_bank_reset_state(sdsp_b200_iir_bank bank);
/* fp32 banks evaluate the recurrence in difference form (d_j[n] = v_j[n] - v_j[n-1] is carried next to v_j: this is what
 * keeps an fp32 narrow-band filter within 1e-4 of the fp64 reference, see simpledsp_b200/csrc/iir_core.cuh) and so hold one
 * more number per section than the reference's m_mem: diff[count][sections].  set_state() starts it at mem[j+1][0] - mem[j+1][1]
 * (rounded to fp32), which is right up to the rounding of one addition; a caller that checkpoints an fp32 bank and wants the
 * resumed stream to be BIT-identical to the uninterrupted one saves and restores diff as well (set_state first, then
 * set_state_diff).  fp64 banks have no such state: get gives zeros, set is a no-op. */
int sdsp_b200_iir_bank_set_state_diff(sdsp_b200_iir_bank bank, size_t first, size_t count, const double *diff);
int sdsp_b200_iir_bank_get_state_diff(sdsp_b200_iir_bank bank, size_t first, size_t count, double *diff);
/* casc_2o_iir<m_t>::process(begin, end) (casc_2o_iir.h:36-80) over every channel of the bank:
 * data[c * channel_stride + i], i < n_samples, filtered in place; history carries to the next call.
 * path: sdsp_b200_iir_path. */
int sdsp_b200_iir_bank_process(sdsp_b200_iir_bank bank, void *data, size_t n_samples, size_t channel_stride,
                               int ptr_kind, int path, void *stream);
int sdsp_b200_iir_bank_describe(sdsp_b200_iir_bank bank, size_t n_samples, size_t channel_stride, int path,
                                char *buf, size_t buf_len);

/* Butterworth designers, host scalar code (reference casc_2o_iir.h:168-194 set_lp_coeff,
 * :140-166 set_hp_coeff, :82-138 set_bp_coeff).  Outputs: *gain, b[sections][3], a[sections][3]. */
int sdsp_b200_iir_design_lp(int sections, double f0, double fs, double gain_in, double *gain, double *b, double *a);
int sdsp_b200_iir_design_hp(int sections, double f0, double fs, double gain_in, double *gain, double *b, double *a);
int sdsp_b200_iir_design_bp(int sections, double f0, double fs, double q, double gain_in, double *gain, double *b,
                            double *a);
/* preload_filter(value) (casc_2o_iir.h:196-214): steady-state history mem[sections+1][2] */
int sdsp_b200_iir_preload_state(int sections, int filter_type, double gain, const double *b, const double *a,
                                double value, double *mem);

/* One-shot convenience used by the drop-in header for a single filter object held on the host:
 * uploads {gain,b,a,mem}, filters data[0..n) in place on the device (sequential path), downloads the
 * new history.  precision is that of `data` ONLY: the arithmetic is fp64 either way, as in the reference, whose
 * object computes and keeps m_mem in double for any sample type (casc_2o_iir.h:13-18, 45-71); float samples are
 * widened on the way in and rounded once on the way out.  (fp32 arithmetic: sdsp_b200_iir_bank_* with SDSP_B200_F32.)
 * sections: 1 .. SDSP_B200_IIR_MAX_SECTIONS_ONCE; more than eight run as a chain of groups of eight over the block. */
int sdsp_b200_iir_process_once(int sections, int numerator, int precision, double gain, const double *b,
                               const double *a, double *mem, void *data, size_t n_samples, int device);

/* ---------------------------------------------------------------- debugging / verification aids
 * Host-side execution of the very code the kernels run per thread (same templates, compiled for the
 * host), so index arithmetic and rounding behaviour can be checked where no GPU exists.  NOT a
 * fallback: nothing in the library calls these. */
int sdsp_b200_debug_emulate_fft(uint32_t n, int precision, int direction, void *data, size_t n_frames);
/* the same for the half-spectrum kernels: back = 0 emulates fft_exec_r2c (real frames in, n/2 + 1 bins out), back = 1 fft_exec_c2r;
 * n = 4 .. 32768 (the sizes with a direct kernel), host buffers */
int sdsp_b200_debug_emulate_r2c(uint32_t n, int precision, int back, const void *in, void *out, size_t n_frames);
int sdsp_b200_debug_emulate_iir(int sections, int numerator, int precision, double gain, const double *b,
                                const double *a, double *mem, void *data, size_t n_samples);
/* the same with the fp32 running differences in / out (diff[sections], may be NULL = as sdsp_b200_iir_bank_set_state) */
int sdsp_b200_debug_emulate_iir_diff(int sections, int numerator, int precision, double gain, const double *b,
                                     const double *a, double *mem, double *diff, void *data, size_t n_samples);
/* the scan path's algorithm (iir_scan_core.cuh) played on the host: whole tiles of 32*chunk samples go
 * through the chunked scan, the remainder through the sequential loop.  force_general != 0 takes the
 * wait-for-predecessor carry path whatever the filter's reach. */
int sdsp_b200_debug_emulate_iir_scan(int sections, int numerator, int precision, double gain, const double *b,
                                     const double *a, double *mem, void *data, size_t n_samples,
                                     int chunk, int force_general);
/* Host only.  The memory of one cascade as the time-split IIR path sees it: the number of samples after which every entry
 * of the n-th power of its one-step transition matrix (zero input) is below 2^-32 (f32) / 2^-62 (f64); 0 = it never decays
 * (unstable or marginal filter).  Segments of a time-split call are at least this long. */
int sdsp_b200_debug_iir_decay_length(int sections, int numerator, int precision, const double *b, const double *a,
                                     unsigned long long *samples);
/* Host only.  The order in which the work-queue FFT kernels (frames larger than one CTA, n = n1 x 256) hand out their items:
 * geom = { tiles per frame and phase, columns per column tile, lag in frames, scratch-ring frames }; item = { 1 = column tile /
 * 0 = row tile, tile index, frame } for ticket q (frames past the batch are empty slots).  A row tile waits for the column tiles
 * of its frame, a column tile for the row tiles of frame - ring: both must hold smaller tickets. */
int sdsp_b200_debug_fft_queue_item(unsigned n, int precision, unsigned long long q, int *geom, int *item);
/* the queue of the real-input 65536-point kernel: geom = {column tiles per frame (8), row tiles per frame (9), lag, ring} */
int sdsp_b200_debug_fft_real_queue_item(int half_spectrum, unsigned long long q, int *geom, int *item);

#ifdef __cplusplus
}
#endif
#endif /* SDSP_B200_H */
