/*
 * sdsp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see sdsp_oracle.h).
 *
 * CPU restatement, in plain C, of the arithmetic simpledsp performs on its two hot paths.
 * Operation order follows the reference so that, built with -ffp-contract=off, results agree
 * with the compiled reference to the last bit wherever libm agrees with GCC's compile-time
 * folding of sin/cos (see oracle/README.md).  Citations are relative to /root/reference.
 */
#include "sdsp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ integer helpers */

/* include/sdsp/fft.h:12-19 : position of the highest set bit */
uint32_t sdsp_oracle_log2(uint32_t num)
{
    uint32_t r = 0;
    while ((num >>= 1) > 0u)
        r++;
    return r;
}

/* include/sdsp/fft.h:21-28 */
uint32_t sdsp_oracle_log4(uint32_t num)
{
    uint32_t r = 0;
    while ((num >>= 2) > 0u)
        r++;
    return r;
}

/* include/sdsp/fft.h:31-37 */
int sdsp_oracle_is_pow2(uint32_t num)
{
    return num != 0 && (num & (num - 1)) == 0;
}

/* include/sdsp/fft.h:40-43 */
int sdsp_oracle_is_pow4(uint32_t num)
{
    return sdsp_oracle_is_pow2(num) && (sdsp_oracle_log2(num) % 2 == 0);
}

/* ------------------------------------------------------------------ twiddle tables */

/*
 * include/sdsp/fft.h:67-119 (the two calculator policies) and 148-194 (calc_trigs).
 * Row i holds trig(2*pi*j / 2^(i+1)) for j < n.  Row 0 alternates +-Value0.  For the other
 * rows only the first quarter wave is evaluated with libm; the value at 90 degrees is exact and
 * the rest is produced by walking an index back and forth over [0, quarter] with sign flips.
 */
int sdsp_oracle_calc_trigs(uint32_t n, int which, double *out)
{
    if (!sdsp_oracle_is_pow2(n) || n < 2)
        return -1;
    const uint32_t rows = sdsp_oracle_log2(n);
    const double value0 = which ? 0.0 : 1.0;   /* fft.h:69-72 / 96-99   */
    const double value90 = which ? 1.0 : 0.0;  /* fft.h:73-76 / 100-103 */
    const double sym0 = which ? -1.0 : 1.0;    /* fft.h:84-87 / 111-114 */
    const double sym90 = which ? 1.0 : -1.0;   /* fft.h:88-91 / 115-118 */

    for (uint32_t i = 0; i < rows; i++) {
        double *row = out + (size_t)i * n;
        const uint32_t pow2 = 1u << (i + 1u);
        row[0] = value0;
        if (i == 0) {
            for (uint32_t j = 1; j < n; j++)
                row[j] = row[j - 1] * -1.0;
            continue;
        }
        const uint32_t quarter = 1u << (i - 1u);
        for (uint32_t j = 1; j < quarter; j++) {
            const double rad = 2 * M_PI * j / pow2; /* fft.h:169 : ((2*pi)*j)/pow2 in double */
            row[j] = which ? sin(rad) : cos(rad);
        }
        row[quarter] = value90;

        int dir = -1;
        double sign = sym90;
        uint32_t bouncy = quarter;
        for (uint32_t j = quarter + 1; j < n; j++) {
            bouncy = (uint32_t)((int)bouncy + dir);
            row[j] = row[bouncy] * sign;
            if (bouncy == 0) {
                dir = 1;
                sign *= sym0;
            } else if (bouncy == quarter) {
                dir = -1;
                sign *= sym90;
            }
        }
    }
    return 0;
}

/* include/sdsp/fft.h:197-214 : W[i][j] = (cos, Sign * -1 * sin), Sign = +1 forward / -1 reverse */
int sdsp_oracle_calc_wcoeffs(uint32_t n, int inverse, double *out)
{
    if (!sdsp_oracle_is_pow2(n) || n < 2)
        return -1;
    const uint32_t rows = sdsp_oracle_log2(n);
    const size_t count = (size_t)rows * n;
    double *c = (double *)malloc(count * sizeof(double));
    double *s = (double *)malloc(count * sizeof(double));
    if (!c || !s) {
        free(c);
        free(s);
        return -2;
    }
    sdsp_oracle_calc_trigs(n, 0, c);
    sdsp_oracle_calc_trigs(n, 1, s);
    const double sign = inverse ? -1.0 : 1.0; /* fft.h:123-126 / 137-140 */
    for (size_t k = 0; k < count; k++) {
        out[2 * k] = c[k];
        out[2 * k + 1] = sign * -1.0 * s[k];
    }
    free(c);
    free(s);
    return 0;
}

/* ------------------------------------------------------------------ digit reversal */

/*
 * include/sdsp/fft.h:217-236.  Swap the outermost base-`base` digits pairwise, moving inward;
 * an unpaired middle digit stays where it is.
 */
uint32_t sdsp_oracle_digit_reverse(uint32_t n, uint32_t base, uint32_t idx)
{
    const uint32_t bits = sdsp_oracle_log2(base);
    uint32_t shift = sdsp_oracle_log2(n) - bits;
    uint32_t upper = (base - 1) << shift;
    uint32_t lower = base - 1;
    uint32_t r = 0;
    while (upper > lower) {
        r |= (idx & upper) >> shift;
        r |= (idx & lower) << shift;
        upper >>= bits;
        lower <<= bits;
        shift -= bits * 2;
    }
    if (upper == lower)
        r |= idx & upper;
    return r;
}

/* include/sdsp/fft.h:238-256 */
int sdsp_oracle_swap_lookup(uint32_t n, uint32_t base, uint32_t *out)
{
    if (base == 2 ? !sdsp_oracle_is_pow2(n) : !sdsp_oracle_is_pow4(n))
        return -1;
    for (uint32_t i = 0; i < n; i++)
        out[i] = sdsp_oracle_digit_reverse(n, base, i);
    for (uint32_t i = 1; i + 1 < n; i++) {
        const uint32_t i2 = out[i];
        if (i2 != i)
            out[i2] = i2;
    }
    return 0;
}

/* ------------------------------------------------------------------ FFT */

typedef struct {
    double re, im;
} cplx;

/* std::complex<double> product as libstdc++/libgcc evaluate it for finite operands */
static inline cplx cmul(cplx x, cplx w)
{
    cplx r;
    r.re = x.re * w.re - x.im * w.im;
    r.im = x.re * w.im + x.im * w.re;
    return r;
}
static inline cplx cadd(cplx a, cplx b)
{
    cplx r = { a.re + b.re, a.im + b.im };
    return r;
}
static inline cplx csub(cplx a, cplx b)
{
    cplx r = { a.re - b.re, a.im - b.im };
    return r;
}

/* the linear swap sweep both transforms use: fft.h:269-273 and 351-355 */
static void swap_sweep(cplx *d, const uint32_t *lut, uint32_t n)
{
    for (uint32_t i = 1; i + 1 < n; i++) {
        const uint32_t i2 = lut[i];
        if (i2 != i) {
            cplx t = d[i];
            d[i] = d[i2];
            d[i2] = t;
        }
    }
}

/* fft.h:128-132 : reverse transform multiplies every element by the double (1.0 / N) */
static void scale_inverse(cplx *d, uint32_t n)
{
    const double s = 1.0 / n;
    for (uint32_t i = 0; i < n; i++) {
        d[i].re *= s;
        d[i].im *= s;
    }
}

/* plan cache so repeated calls (batches, benchmarks) do not rebuild the tables, mirroring the
 * reference's function-local constexpr statics (fft.h:264-265, 307-309) */
typedef struct {
    uint32_t n;
    int inverse;
    uint32_t base;
    cplx *w;       /* [log2 n][n] */
    uint32_t *lut; /* [n] */
} plan_t;

#define MAX_PLANS 64
static plan_t g_plans[MAX_PLANS];
static int g_nplans = 0;

static const plan_t *get_plan(uint32_t n, int inverse, uint32_t base)
{
    for (int i = 0; i < g_nplans; i++)
        if (g_plans[i].n == n && g_plans[i].inverse == inverse && g_plans[i].base == base)
            return &g_plans[i];
    if (g_nplans == MAX_PLANS)
        return NULL;
    plan_t p;
    p.n = n;
    p.inverse = inverse;
    p.base = base;
    p.w = (cplx *)malloc((size_t)sdsp_oracle_log2(n) * n * sizeof(cplx));
    p.lut = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    if (!p.w || !p.lut)
        return NULL;
    sdsp_oracle_calc_wcoeffs(n, inverse, (double *)p.w);
    sdsp_oracle_swap_lookup(n, base, p.lut);
    g_plans[g_nplans] = p;
    return &g_plans[g_nplans++];
}

/*
 * include/sdsp/fft.h:258-299.  Decimation in time: bit-reversal swap, then log2(n) butterfly
 * stages with half-span h = 2^i, t = d[j+k+h] * W[i][j+k], (a + t, a - t).
 */
int sdsp_oracle_fft_radix2(double *data, uint32_t n, int inverse)
{
    if (!sdsp_oracle_is_pow2(n) || n < 2)
        return -1;
    const plan_t *p = get_plan(n, inverse, 2);
    if (!p)
        return -2;
    cplx *d = (cplx *)data;
    swap_sweep(d, p->lut, n);

    const uint32_t stages = sdsp_oracle_log2(n);
    for (uint32_t i = 0; i < stages; i++) {
        const uint32_t h = 1u << i;
        const cplx *w = p->w + (size_t)i * n;
        for (uint32_t j = 0; j < n; j += (h << 1)) {
            for (uint32_t k = 0; k < h; k++) {
                const uint32_t lo = j + k, hi = j + k + h;
                const cplx t = cmul(d[hi], w[lo]);
                const cplx a = d[lo];
                d[lo] = cadd(a, t);
                d[hi] = csub(a, t);
            }
        }
    }
    if (inverse)
        scale_inverse(d, n);
    return 0;
}

/*
 * include/sdsp/fft.h:301-360.  Decimation in frequency, radix 4.  The output twiddles of stage
 * i-1 are applied at the *input* of stage i: every operand is pre-multiplied by
 * W_n^(offset * 4^(i-1) * (group mod 4)) unless that exponent is zero, then the 4-point DFT is
 * formed; a base-4 digit-reversal swap follows the last stage.
 */
int sdsp_oracle_fft_radix4(double *data, uint32_t n, int inverse)
{
    if (!sdsp_oracle_is_pow4(n) || n < 4)
        return -1;
    const plan_t *p = get_plan(n, inverse, 4);
    if (!p)
        return -2;
    cplx *d = (cplx *)data;
    const cplx *w = p->w + (size_t)(sdsp_oracle_log2(n) - 1) * n; /* fft.h:309 finest row only */
    const double sign = inverse ? -1.0 : 1.0;
    const uint32_t stages = sdsp_oracle_log4(n);

    for (uint32_t i = 0; i < stages; i++) {
        const uint32_t g = n / (4u << (2u * i));              /* fft.h:312 */
        const uint32_t s = i > 0 ? 1u << ((i - 1u) * 2u) : 0; /* fft.h:313 */
        uint32_t group = 0;
        for (uint32_t j = 0; j < n; j += (g << 2)) {
            const uint32_t branch = group % 4;
            for (uint32_t k = 0; k < g; k++) {
                uint32_t idx[4];
                cplx t[4];
                for (int m = 0; m < 4; m++) {
                    idx[m] = k + (uint32_t)m * g + j;
                    const uint32_t e = (idx[m] - j) * s * branch; /* fft.h:322-325 */
                    t[m] = e > 0 ? cmul(d[idx[m]], w[e]) : d[idx[m]];
                }
                /* fft.h:339-340 : Sign * complex(-im, re) */
                const cplx t1i = { sign * -t[1].im, sign * t[1].re };
                const cplx t3i = { sign * -t[3].im, sign * t[3].re };
                /* fft.h:342-345, left-to-right association */
                d[idx[0]] = cadd(cadd(cadd(t[0], t[1]), t[2]), t[3]);
                d[idx[1]] = cadd(csub(csub(t[0], t1i), t[2]), t3i);
                d[idx[2]] = csub(cadd(csub(t[0], t[1]), t[2]), t[3]);
                d[idx[3]] = csub(csub(cadd(t[0], t1i), t[2]), t3i);
            }
            group++;
        }
    }
    swap_sweep(d, p->lut, n);
    if (inverse)
        scale_inverse(d, n);
    return 0;
}

int sdsp_oracle_fft_batch(double *data, uint32_t n, size_t frames, int radix, int inverse)
{
    for (size_t f = 0; f < frames; f++) {
        double *frame = data + f * 2 * (size_t)n;
        const int rc = radix == 4 ? sdsp_oracle_fft_radix4(frame, n, inverse) :
                                    sdsp_oracle_fft_radix2(frame, n, inverse);
        if (rc)
            return rc;
    }
    return 0;
}

/* ------------------------------------------------------------------ cascaded biquad IIR */

size_t sdsp_oracle_iir_sizeof(void)
{
    return sizeof(sdsp_oracle_iir);
}

/* casc_2o_iir.h:10-26 (and 219-226, 269-272 for the fixed-numerator classes) */
int sdsp_oracle_iir_init(sdsp_oracle_iir *f, int sections, int kind)
{
    if (sections < 1 || sections > SDSP_ORACLE_MAX_SECTIONS || kind < 0 || kind > 3)
        return -1;
    memset(f, 0, sizeof(*f));
    f->sections = sections;
    f->kind = kind;
    f->gain = 1.0;
    return 0;
}

/* casc_2o_iir.h:28-34 : coefficients and type, never the history */
void sdsp_oracle_iir_copy_coeff_from(sdsp_oracle_iir *f, const sdsp_oracle_iir *o)
{
    f->gain = o->gain;
    memcpy(f->b, o->b, sizeof(f->b));
    memcpy(f->a, o->a, sizeof(f->a));
    f->ftype = o->ftype;
}

/* casc_2o_iir.h:168-194 (lp) and 140-166 (hp); the fixed-numerator classes repeat the same
 * design at 297-321 / 355-379 without storing b */
static void design_lp_hp(sdsp_oracle_iir *f, double f0, double fs, double gain_in, int highpass)
{
    const int m = f->sections;
    f->gain = gain_in;
    f->ftype = highpass ? 2 : 1;
    const double e0 = 2 * M_PI * f0 / fs;
    for (int k = 0; k < m; k++) {
        const double dk = 2 * sin((2 * k + 1) * M_PI / (4.0 * m));
        const double t = dk * sin(e0) / 2;
        const double dnm = 1 + t;
        const double beta1 = (1 - t) / dnm / 2;
        const double gamma1 = (0.5 + beta1) * cos(e0);
        const double alpha1 = highpass ? (0.5 + beta1 + gamma1) / 4 : (0.5 + beta1 - gamma1) / 4;
        f->gain *= 2 * alpha1;
        f->b[k][0] = 1.0;
        f->b[k][1] = highpass ? -2.0 : 2.0;
        f->b[k][2] = 1.0;
        f->a[k][0] = 1;
        f->a[k][1] = -2 * gamma1;
        f->a[k][2] = 2 * beta1;
    }
}

int sdsp_oracle_iir_set_lp(sdsp_oracle_iir *f, double f0, double fs, double gain)
{
    if (f->kind != 0 && f->kind != 1)
        return -1;
    design_lp_hp(f, f0, fs, gain, 0);
    return 0;
}

int sdsp_oracle_iir_set_hp(sdsp_oracle_iir *f, double f0, double fs, double gain)
{
    if (f->kind != 0 && f->kind != 2)
        return -1;
    design_lp_hp(f, f0, fs, gain, 1);
    return 0;
}

/* casc_2o_iir.h:82-138 (and 413-467): m/2 pole pairs, sections 2k and 2k+1 */
int sdsp_oracle_iir_set_bp(sdsp_oracle_iir *f, double f0, double fs, double q, double gain_in)
{
    if (f->kind != 0 && f->kind != 3)
        return -1;
    const int m = f->sections;
    f->gain = gain_in;
    const double q2 = 2 * q;
    f->ftype = 3;
    const double e0 = 2 * M_PI * f0 / fs;
    double dnm = sin(e0);
    const double de = 2 * tan(e0 / q2) / dnm;
    for (int k = 0; k < m / 2; k++) {
        const double d = 2 * sin((2 * k + 1) * M_PI / (2.0 * m));
        const double a = (1 + de * de / 4.0) * 2 / d / de;
        const double dk = sqrt(de * d / (a + sqrt(a * a - 1)));
        const double b = d * de / dk / 2.0;
        const double w = b + sqrt(b * b - 1);
        double t = tan(e0 / 2.0);
        const double e1 = 2.0 * atan(t / w);
        const double e2 = 2.0 * atan(w * t);

        t = dk * sin(e1) / 2.0;
        dnm = (1 + t);
        const double beta1 = (1 - t) / dnm / 2.0;
        t = dk * sin(e2) / 2.0;
        dnm = (1 + t);
        const double beta2 = (1 - t) / dnm / 2.0;

        const double gamma1 = (0.5 + beta1) * cos(e1);
        const double gamma2 = (0.5 + beta2) * cos(e2);

        t = sqrt(1 + (w - 1 / w) / dk * (w - 1 / w) / dk);
        const double alpha1 = (0.5 - beta1) * t / 2.0;
        const double alpha2 = (0.5 - beta2) * t / 2.0;

        f->gain *= 4 * alpha1 * alpha2;
        for (int h = 0; h < 2; h++) {
            f->b[2 * k + h][0] = 1.0;
            f->b[2 * k + h][1] = 0;
            f->b[2 * k + h][2] = -1.0;
            f->a[2 * k + h][0] = 1;
        }
        f->a[2 * k][1] = -2 * gamma1;
        f->a[2 * k + 1][1] = -2 * gamma2;
        f->a[2 * k][2] = 2 * beta1;
        f->a[2 * k + 1][2] = 2 * beta2;
    }
    return 0;
}

/* casc_2o_iir.h:196-214 : DC steady state; only the low-pass case propagates past row 0 */
void sdsp_oracle_iir_preload(sdsp_oracle_iir *f, double value)
{
    double v = value * f->gain;
    double mem[SDSP_ORACLE_MAX_SECTIONS + 1][3];
    memset(mem, 0, sizeof(mem));
    for (int i = 0; i < 3; i++)
        mem[0][i] = v;
    if (f->ftype == 1) {
        for (int j = 1; j < f->sections + 1; j++) {
            v /= 1 + f->a[j - 1][1] + f->a[j - 1][2];
            v *= f->b[j - 1][0] + f->b[j - 1][1] + f->b[j - 1][2];
            for (int i = 0; i < 3; i++)
                mem[j][i] = v;
        }
    }
    memcpy(f->mem, mem, sizeof(mem));
}

/*
 * casc_2o_iir.h:36-80 (generic) and 228-263 + 286-295 / 344-353 / 402-411 (fixed numerators).
 * Three-slot circular history per row; row j is the output history of section j-1 and the input
 * history of section j.  b0 is implicitly one, a0 is never read.
 */
void sdsp_oracle_iir_process(sdsp_oracle_iir *f, double *data, size_t n)
{
    const int m = f->sections;
    int p = f->pos;
    double(*y)[3] = f->mem;
    for (size_t i = 0; i < n; i++) {
        y[0][p] = data[i] * f->gain;
        int d1 = p - 1;
        if (d1 < 0)
            d1 += 3;
        int d2 = p - 2;
        if (d2 < 0)
            d2 += 3;
        for (int j = 0; j < m; j++) {
            y[j + 1][p] = y[j][p];
            switch (f->kind) {
            case 0:
                y[j + 1][p] += y[j][d1] * f->b[j][1] - y[j + 1][d1] * f->a[j][1];
                y[j + 1][p] += y[j][d2] * f->b[j][2] - y[j + 1][d2] * f->a[j][2];
                break;
            case 1: /* casc_2o_iir.h:292-293 */
                y[j + 1][p] += y[j][d1] + y[j][d1] - y[j + 1][d1] * f->a[j][1];
                y[j + 1][p] += y[j][d2] - y[j + 1][d2] * f->a[j][2];
                break;
            case 2: /* casc_2o_iir.h:350-351 */
                y[j + 1][p] += -y[j][d1] - y[j][d1] - y[j + 1][d1] * f->a[j][1];
                y[j + 1][p] += y[j][d2] - y[j + 1][d2] * f->a[j][2];
                break;
            default: /* casc_2o_iir.h:408-409 */
                y[j + 1][p] += -y[j + 1][d1] * f->a[j][1];
                y[j + 1][p] += -y[j][d2] - y[j + 1][d2] * f->a[j][2];
                break;
            }
        }
        data[i] = y[m][p];
        p++;
        if (p > 2)
            p = 0;
    }
    f->pos = p;
}
