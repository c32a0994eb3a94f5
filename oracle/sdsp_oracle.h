/*
 * sdsp_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU-only restatement of the two simpledsp hot paths (radix-2 / radix-4 complex FFT
 * and the cascaded second-order-section IIR).  It exists only so the CUDA path can be checked
 * against the reference's arithmetic on identical inputs.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (simpledsp_b200/, include/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the reference's nine golden impulse responses (test_data/impulse_response/ *.csv,
 *       committed as tests/golden/impulse_response.npz),
 *   (2) the analytic known-answer vectors of the reference's FFT tests (test/testFFT.cpp), and
 *   (3) outputs of the unmodified reference headers compiled here (oracle/_ref/libsdsp_ref.so,
 *       vectors committed as tests/golden/ref_vectors.npz).
 *
 * Every function cites the reference lines (relative to /root/reference) it follows.
 */
#ifndef SDSP_ORACLE_H
#define SDSP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/sdsp/fft.h:12-43 */
uint32_t sdsp_oracle_log2(uint32_t num);
uint32_t sdsp_oracle_log4(uint32_t num);
int sdsp_oracle_is_pow2(uint32_t num);
int sdsp_oracle_is_pow4(uint32_t num);

/* include/sdsp/fft.h:148-194 -- one trig table, out[log2(n)][n]; which: 0 = cosine, 1 = sine */
int sdsp_oracle_calc_trigs(uint32_t n, int which, double *out);
/* include/sdsp/fft.h:197-214 -- W table, out[log2(n)][n][2] (re,im); inverse: 0 forward, 1 reverse */
int sdsp_oracle_calc_wcoeffs(uint32_t n, int inverse, double *out);

/* include/sdsp/fft.h:217-236 */
uint32_t sdsp_oracle_digit_reverse(uint32_t n, uint32_t base, uint32_t idx);
/* include/sdsp/fft.h:238-256 -- the "half" swap table the reference sweeps linearly */
int sdsp_oracle_swap_lookup(uint32_t n, uint32_t base, uint32_t *out);

/* include/sdsp/fft.h:258-299 / 301-360.  data = n interleaved (re,im) doubles, in place.
 * Returns 0, or -1 if n is not a power of 2 (radix2) / power of 4 (radix4). */
int sdsp_oracle_fft_radix2(double *data, uint32_t n, int inverse);
int sdsp_oracle_fft_radix4(double *data, uint32_t n, int inverse);
/* frames contiguous frames of n points each, single thread */
int sdsp_oracle_fft_batch(double *data, uint32_t n, size_t frames, int radix, int inverse);

/* ---- cascaded 2nd-order IIR: include/sdsp/casc_2o_iir.h ---- */
#define SDSP_ORACLE_MAX_SECTIONS 16

/* numerator kinds: 0 = runtime b coefficients (casc_2o_iir), 1/2/3 = fixed lp/hp/bp numerators
 * (casc_2o_iir_lp / _hp / _bp, casc_2o_iir.h:266-468) */
typedef struct sdsp_oracle_iir {
    int sections;                                 /* m_t */
    int kind;                                     /* numerator kind */
    int pos;                                      /* m_pos          casc_2o_iir.h:11 */
    int ftype;                                    /* filter_type    casc_2o_iir.h:20 */
    double gain;                                  /* m_gain         casc_2o_iir.h:13 */
    double mem[SDSP_ORACLE_MAX_SECTIONS + 1][3];  /* m_mem          casc_2o_iir.h:15 */
    double b[SDSP_ORACLE_MAX_SECTIONS][3];        /* m_b_coeff      casc_2o_iir.h:17 */
    double a[SDSP_ORACLE_MAX_SECTIONS][3];        /* m_a_coeff      casc_2o_iir.h:18 */
} sdsp_oracle_iir;

int sdsp_oracle_iir_init(sdsp_oracle_iir *f, int sections, int kind);
void sdsp_oracle_iir_copy_coeff_from(sdsp_oracle_iir *f, const sdsp_oracle_iir *other);
int sdsp_oracle_iir_set_lp(sdsp_oracle_iir *f, double f0, double fs, double gain);
int sdsp_oracle_iir_set_hp(sdsp_oracle_iir *f, double f0, double fs, double gain);
int sdsp_oracle_iir_set_bp(sdsp_oracle_iir *f, double f0, double fs, double q, double gain);
void sdsp_oracle_iir_preload(sdsp_oracle_iir *f, double value);
void sdsp_oracle_iir_process(sdsp_oracle_iir *f, double *data, size_t n);
size_t sdsp_oracle_iir_sizeof(void);

#ifdef __cplusplus
}
#endif
#endif
