/*
 * liquid_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH).  See liquid_oracle.h.
 *
 * PARITY UNPINNED (no liquid-dsp sources/binary and no reference tests exist here); the known answers that do
 * exist are checked in tests/test_oracle_known_answers.py.
 *
 * Every function names the liquid-dsp v1.3.2 source file whose published algorithm it restates and the
 * reference (composable-sdr) call site that uses it.  Arithmetic is float32 and strictly sequential, with the
 * same implicit double promotions the C expressions in liquid have.  Build with -ffp-contract=off so results
 * do not depend on the host's FMA support (oracle/Makefile).
 */
#include "liquid_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static int g_opt[ORC_OPT_COUNT] = { 0, 1, 0 };
void orc_set_option(int opt, int value) { if (opt >= 0 && opt < ORC_OPT_COUNT) g_opt[opt] = value; }
int  orc_get_option(int opt) { return (opt >= 0 && opt < ORC_OPT_COUNT) ? g_opt[opt] : -1; }

static inline orc_cf32 cmk(float re, float im) { orc_cf32 z; z.re = re; z.im = im; return z; }

/* ============================================================================================
 * Filter design helpers
 * ========================================================================================== */

/* liquid src/filter/src/firdes.c kaiser_beta_As() */
float orc_kaiser_beta_As(float As)
{
    As = fabsf(As);
    float beta;
    if (As > 50.0f)      beta = 0.1102f * (As - 8.7f);
    else if (As > 21.0f) beta = 0.5842f * powf(As - 21.0f, 0.4f) + 0.07886f * (As - 21.0f);
    else                 beta = 0.0f;
    return beta;
}

/* liquid src/filter/src/firdes.c estimate_req_filter_len() -> Kaiser's formula, truncated to unsigned */
unsigned orc_estimate_req_filter_len(float df, float As)
{
    float n = (As - 7.95f) / (14.26f * df);
    return (unsigned)n;
}

/* Modified Bessel function of the first kind, order 0.  liquid (src/math/src/math.bessel.c besseli0f) sums a
 * 32-term float series; the oracle evaluates the same series in double (difference ~1e-6 relative, inside the
 * tolerance budget). */
static double besseli0(double z)
{
    double y = 1.0, t = 1.0;
    for (int k = 1; k < 64; k++) {
        t *= (0.5 * z) / (double)k;
        y += t * t;
        if (t * t < 1e-20 * y) break;
    }
    return y;
}

/* liquid src/math/src/windows.c kaiser(): r = 2t/(N-1)  (pinned by the firpfbch print in images/ex1_5.gif) */
static double kaiser_win(unsigned i, unsigned N, double beta, double mu)
{
    double t = (double)i - (double)(N - 1) / 2.0 + mu;
    double r = 2.0 * t / (double)(N - 1);
    double a = 1.0 - r * r;
    if (a < 0.0) a = 0.0;
    return besseli0(beta * sqrt(a)) / besseli0(beta);
}

static double sinc_d(double x)
{
    if (fabs(x) < 1e-9) return 1.0;
    return sin(M_PI * x) / (M_PI * x);
}

/* liquid src/filter/src/firdes.c liquid_firdes_kaiser(): h[i] = sinc(2 fc t) * kaiser(i), no normalisation */
void orc_firdes_kaiser(unsigned n, float fc, float As, float mu, float *h)
{
    double beta = (double)orc_kaiser_beta_As(As);
    for (unsigned i = 0; i < n; i++) {
        double t = (double)i - (double)(n - 1) / 2.0 + (double)mu;
        double h1 = sinc_d(2.0 * (double)fc * t);
        double h2 = kaiser_win(i, n, beta, (double)mu);
        h[i] = (float)(h1 * h2);
    }
}

/* liquid dotprod_crcf: real taps x complex samples.  liquid's SIMD variants sum lanes in a different order;
 * the oracle sums left to right. */
static inline orc_cf32 dot_rc(const float *h, const orc_cf32 *x, unsigned n)
{
    float re = 0.0f, im = 0.0f;
    for (unsigned i = 0; i < n; i++) {
        re += h[i] * x[i].re;
        im += h[i] * x[i].im;
    }
    return cmk(re, im);
}

/* liquid windowcf: sliding buffer, oldest first, newest last */
typedef struct { orc_cf32 *v; unsigned n; } win_cf;
static void win_init(win_cf *w, unsigned n) { w->n = n; w->v = (orc_cf32 *)calloc(n ? n : 1, sizeof(orc_cf32)); }
static void win_free(win_cf *w) { free(w->v); w->v = NULL; }
static inline void win_push(win_cf *w, orc_cf32 x)
{
    memmove(w->v, w->v + 1, (w->n - 1) * sizeof(orc_cf32));
    w->v[w->n - 1] = x;
}

/* ============================================================================================
 * nco_crcf -- liquid src/nco/src/nco.c (v1.3.x: uint32 phase accumulator, 1024-entry sine table)
 * Reference: ncoCreate/ncoMixDown/ncoMixUp Liquid.chs:782-809; pre-rotation Liquid.chs:811-821,847
 * ========================================================================================== */
struct orc_nco_s { int type; uint32_t theta, d_theta; float alpha, beta; };

static float g_sintab[1024];
static int   g_sintab_ready = 0;
const float *orc_nco_sintab(void)
{
    if (!g_sintab_ready) {
        /* NCO(_create): q->sintab[i] = SIN(2.0f*M_PI*(float)(i)/1024.0f)  (argument evaluated in double) */
        for (unsigned i = 0; i < 1024; i++)
            g_sintab[i] = sinf((float)(2.0f * M_PI * (float)i / 1024.0f));
        g_sintab_ready = 1;
    }
    return g_sintab;
}

/* NCO(_constrain): float32 fraction of a turn -> uint32.  0.525 -> 0x86666600 (known answer #2) */
static uint32_t nco_constrain(float theta)
{
    float p = (float)(theta * 0.159154943091895);
    float fpart = p - (float)((long)p);
    if (fpart < 0.0f) fpart = (float)(fpart + 1.0);
    float scaled = fpart * 4294967296.0f;            /* (float)0xffffffff rounds to 2^32 */
    if (scaled >= 4294967296.0f) return 0u;          /* C UB in liquid; wraps to 0 on x86 */
    return (uint32_t)scaled;
}

orc_nco orc_nco_crcf_create(int type)
{
    orc_nco q = (orc_nco)calloc(1, sizeof(*q));
    q->type = type;
    q->alpha = 0.1f; q->beta = sqrtf(0.1f);          /* NCO_PLL_BANDWIDTH_DEFAULT */
    (void)orc_nco_sintab();
    return q;
}
void orc_nco_crcf_destroy(orc_nco q) { free(q); }
void orc_nco_crcf_set_frequency(orc_nco q, float dtheta) { q->d_theta = nco_constrain(dtheta); }
void orc_nco_crcf_set_phase(orc_nco q, float theta) { q->theta = nco_constrain(theta); }
uint32_t orc_nco_crcf_get_phase_word(orc_nco q) { return q->theta; }
uint32_t orc_nco_crcf_get_freq_word(orc_nco q) { return q->d_theta; }
void orc_nco_crcf_step(orc_nco q) { q->theta += q->d_theta; }

static inline void nco_sincos(const struct orc_nco_s *q, float *s, float *c)
{
    if (q->type == 1 && g_opt[ORC_OPT_VCO_DIRECT]) {
        float phi = (float)(2.0 * M_PI * (double)q->theta / 4294967296.0);
        *s = sinf(phi); *c = cosf(phi);
    } else {
        /* NCO(_index): round to nearest of 1024 table entries */
        unsigned idx = ((q->theta + (1u << 21)) >> 22) & 0x3ff;
        *s = g_sintab[idx];
        *c = g_sintab[(idx + 256) & 0x3ff];
    }
}
static void nco_pll_set_bandwidth(orc_nco q, float bw) { q->alpha = bw; q->beta = sqrtf(bw); }
static void nco_pll_step(orc_nco q, float dphi)
{
    q->d_theta += nco_constrain(dphi * q->alpha);
    q->theta   += nco_constrain(dphi * q->beta);
}
/* the scalar members the reference's pilot PLL drives per sample (pllCreate / pllStep, Liquid.chs:959-988) */
void orc_nco_crcf_pll_set_bandwidth(orc_nco q, float bw) { nco_pll_set_bandwidth(q, bw); }
void orc_nco_crcf_pll_step(orc_nco q, float dphi) { nco_pll_step(q, dphi); }
/* NCO(_get_phase): 2.0f*M_PI*(float)theta / (float)(1LLU<<32), evaluated in double, returned as float */
float orc_nco_crcf_get_phase(orc_nco q) { return (float)(2.0 * M_PI * (double)(float)q->theta / 4294967296.0); }
void orc_nco_crcf_cexpf(orc_nco q, orc_cf32 *y) { float s, c; nco_sincos(q, &s, &c); y->re = c; y->im = s; }
static inline orc_cf32 nco_mix_down1(const struct orc_nco_s *q, orc_cf32 x)
{
    float s, c; nco_sincos(q, &s, &c);
    /* x * conj(c + j s) */
    return cmk(x.re * c + x.im * s, x.im * c - x.re * s);
}
void orc_nco_crcf_mix_block_down(orc_nco q, const orc_cf32 *x, orc_cf32 *y, unsigned n)
{
    for (unsigned i = 0; i < n; i++) { y[i] = nco_mix_down1(q, x[i]); q->theta += q->d_theta; }
}
void orc_nco_crcf_mix_block_up(orc_nco q, const orc_cf32 *x, orc_cf32 *y, unsigned n)
{
    for (unsigned i = 0; i < n; i++) {
        float s, c; nco_sincos(q, &s, &c);
        y[i] = cmk(x[i].re * c - x[i].im * s, x[i].im * c + x[i].re * s);
        q->theta += q->d_theta;
    }
}

/* ============================================================================================
 * resamp2_crcf half-band -- liquid src/filter/src/resamp2.c
 * ========================================================================================== */
typedef struct { unsigned m; float *h1; win_cf w0, w1; } resamp2_t;

static void resamp2_init(resamp2_t *q, unsigned m, float f0, float As)
{
    (void)f0; /* msresamp always passes 0 */
    q->m = m;
    unsigned h_len = 4 * m + 1;
    double *h = (double *)malloc(h_len * sizeof(double));
    double beta = (double)orc_kaiser_beta_As(As);
    for (unsigned i = 0; i < h_len; i++) {
        double t = (double)i - (double)(h_len - 1) / 2.0;
        h[i] = sinc_d(t / 2.0) * kaiser_win(i, h_len, beta, 0.0);
    }
    q->h1 = (float *)malloc(2 * m * sizeof(float));
    unsigned j = 0;
    for (unsigned i = 1; i < h_len; i += 2) q->h1[j++] = (float)h[h_len - i - 1];
    free(h);
    win_init(&q->w0, 2 * m);
    win_init(&q->w1, 2 * m);
}
static void resamp2_free(resamp2_t *q) { free(q->h1); win_free(&q->w0); win_free(&q->w1); }

/* RESAMP2(_decim_execute): x[0] -> filter branch, x[1] -> delay branch, y = y0 + y1 (gain 2, removed by zeta) */
static inline orc_cf32 resamp2_decim(resamp2_t *q, const orc_cf32 *x)
{
    win_push(&q->w1, x[0]);
    orc_cf32 y1 = dot_rc(q->h1, q->w1.v, 2 * q->m);
    win_push(&q->w0, x[1]);
    orc_cf32 y0 = q->w0.v[q->m - 1];
    return cmk(y0.re + y1.re, y0.im + y1.im);
}
/* RESAMP2(_interp_execute): y[0] = delay branch, y[1] = filter branch */
static inline void resamp2_interp(resamp2_t *q, orc_cf32 x, orc_cf32 *y)
{
    win_push(&q->w0, x);
    y[0] = q->w0.v[q->m - 1];
    win_push(&q->w1, x);
    y[1] = dot_rc(q->h1, q->w1.v, 2 * q->m);
}

/* ============================================================================================
 * msresamp2_crcf -- liquid src/filter/src/msresamp2.c, created by msresamp as (type, S, 0.4, 0, As)
 * ========================================================================================== */
typedef struct {
    int interp; unsigned S, M; float zeta;
    unsigned *m_stage; resamp2_t *st; orc_cf32 *b0, *b1;
} msresamp2_t;

static void msresamp2_init(msresamp2_t *q, int interp, unsigned S, float fc, float f0, float As)
{
    q->interp = interp; q->S = S; q->M = 1u << S; q->zeta = 1.0f / (float)q->M;
    q->m_stage = (unsigned *)calloc(S ? S : 1, sizeof(unsigned));
    q->st = (resamp2_t *)calloc(S ? S : 1, sizeof(resamp2_t));
    q->b0 = (orc_cf32 *)calloc(q->M, sizeof(orc_cf32));
    q->b1 = (orc_cf32 *)calloc(q->M, sizeof(orc_cf32));
    float As_stage = As + 5.0f;
    for (unsigned i = 0; i < S; i++) {
        fc = (i == 1) ? (float)((0.5 - fc) / 2.0f) : 0.5f * fc;
        f0 = 0.5f * f0;
        float ft = 2 * (0.25f - fc);
        unsigned h_len = orc_estimate_req_filter_len(ft, As_stage);
        unsigned m = (unsigned)ceilf((float)(h_len - 1) / 4.0f);
        q->m_stage[i] = m < 3 ? 3 : m;
        resamp2_init(&q->st[i], q->m_stage[i], f0, As_stage);
    }
}
static void msresamp2_free(msresamp2_t *q)
{
    for (unsigned i = 0; i < q->S; i++) resamp2_free(&q->st[i]);
    free(q->m_stage); free(q->st); free(q->b0); free(q->b1);
}
/* MSRESAMP2(_decim_execute): 2^S inputs -> 1 output; highest-rate stage is index S-1 */
static orc_cf32 msresamp2_decim(msresamp2_t *q, const orc_cf32 *x)
{
    if (q->S == 0) return x[0];
    const orc_cf32 *b0 = x; orc_cf32 *b1 = q->b1;
    for (unsigned s = 0; s < q->S; s++) {
        unsigned g = q->S - s - 1, k = 1u << g;
        for (unsigned i = 0; i < k; i++) b1[i] = resamp2_decim(&q->st[g], &b0[2 * i]);
        b0 = (s % 2) == 0 ? q->b1 : q->b0;
        b1 = (s % 2) == 0 ? q->b0 : q->b1;
    }
    return cmk(b0[0].re * q->zeta, b0[0].im * q->zeta);
}
/* MSRESAMP2(_interp_execute): 1 input -> 2^S outputs; stage 0 first */
static void msresamp2_interp(msresamp2_t *q, orc_cf32 x, orc_cf32 *y)
{
    if (q->S == 0) { y[0] = x; return; }
    orc_cf32 *b0 = q->b0, *b1 = q->b1;
    b0[0] = x;
    for (unsigned s = 0; s < q->S; s++) {
        unsigned k = 1u << s;
        if (s == q->S - 1) b1 = y;
        for (unsigned i = 0; i < k; i++) resamp2_interp(&q->st[s], b0[i], &b1[2 * i]);
        b0 = (s % 2) == 0 ? q->b1 : q->b0;
        b1 = (s % 2) == 0 ? q->b0 : q->b1;
    }
}

/* ============================================================================================
 * resamp_crcf arbitrary -- liquid src/filter/src/resamp.fixed.c (fixed-point phase variant, SURVEY A.4 "F")
 * with its firpfb_crcf bank (src/filter/src/firpfb.c)
 * ========================================================================================== */
typedef struct {
    float rate; uint32_t step, phase; unsigned bits, npfb, hsub;
    float *bank;             /* [npfb][hsub]; bank[i][j] multiplies the j-th NEWEST sample: h[i + j*npfb] */
    float *bank_rev;         /* liquid layout: oldest first */
    win_cf w;
} resamp_t;

static void resamp_init(resamp_t *q, float rate, unsigned m, float fc, float As, unsigned npfb)
{
    q->rate = rate;
    q->step = (uint32_t)round((float)(1 << 24) / rate);
    q->phase = 0;
    unsigned bits = 0; while ((1u << bits) < npfb) bits++;     /* liquid_nextpow2 */
    q->bits = bits; q->npfb = 1u << bits;
    unsigned n = 2 * m * q->npfb + 1;
    float *hf = (float *)malloc(n * sizeof(float));
    orc_firdes_kaiser(n, fc / (float)q->npfb, As, 0.0f, hf);
    float gain = 0.0f;
    for (unsigned i = 0; i < n; i++) gain += hf[i];
    gain = (float)q->npfb / gain;
    for (unsigned i = 0; i < n; i++) hf[i] = hf[i] * gain;
    q->hsub = (n - 1) / q->npfb;                               /* = 2m */
    q->bank = (float *)malloc(q->npfb * q->hsub * sizeof(float));
    q->bank_rev = (float *)malloc(q->npfb * q->hsub * sizeof(float));
    for (unsigned i = 0; i < q->npfb; i++)
        for (unsigned j = 0; j < q->hsub; j++) {
            q->bank[i * q->hsub + j] = hf[i + j * q->npfb];
            q->bank_rev[i * q->hsub + (q->hsub - j - 1)] = hf[i + j * q->npfb];
        }
    free(hf);
    win_init(&q->w, q->hsub);
}
static void resamp_free(resamp_t *q) { free(q->bank); free(q->bank_rev); win_free(&q->w); }

/* RESAMP(_execute): push one input, emit while phase < 2^24 */
static inline unsigned resamp_exec(resamp_t *q, orc_cf32 x, orc_cf32 *y)
{
    win_push(&q->w, x);
    unsigned n = 0;
    while (q->phase <= 0x00ffffffu) {
        unsigned idx = q->phase >> (24 - q->bits);
        y[n++] = dot_rc(q->bank_rev + idx * q->hsub, q->w.v, q->hsub);
        q->phase += q->step;
    }
    q->phase -= (1u << 24);
    return n;
}

/* ============================================================================================
 * msresamp_crcf -- liquid src/filter/src/msresamp.c.  Reference: resampler, Liquid.chs:56-117
 * ========================================================================================== */
struct orc_msresamp_s {
    float rate, As; int interp; unsigned S; float rate_arb;
    msresamp2_t hb; resamp_t arb;
    orc_cf32 *buffer; unsigned buffer_index;
};

orc_msresamp orc_msresamp_crcf_create(float r, float As)
{
    if (!(r > 0.0f)) return NULL;
    orc_msresamp q = (orc_msresamp)calloc(1, sizeof(*q));
    q->rate = r; q->As = As;
    q->interp = (r > 1.0f);
    q->rate_arb = r; q->S = 0;
    if (q->interp) { while (q->rate_arb > 2.0f) { q->S++; q->rate_arb *= 0.5f; } }
    else           { while (q->rate_arb < 0.5f) { q->S++; q->rate_arb *= 2.0f; } }
    q->buffer = (orc_cf32 *)calloc(4 + (1u << q->S), sizeof(orc_cf32));
    q->buffer_index = 0;
    msresamp2_init(&q->hb, q->interp, q->S, 0.4f, 0.0f, As);
    if (g_opt[ORC_OPT_RESAMP_FC_OLD]) {
        resamp_init(&q->arb, q->rate_arb, 7, 0.4f, As, 64);
    } else {
        float fc = 0.515f * q->rate_arb; if (fc > 0.49f) fc = 0.49f;
        resamp_init(&q->arb, q->rate_arb, 7, fc, As, 256);
    }
    return q;
}
void orc_msresamp_crcf_destroy(orc_msresamp q)
{
    if (!q) return;
    msresamp2_free(&q->hb); resamp_free(&q->arb); free(q->buffer); free(q);
}
float orc_msresamp_crcf_get_rate(orc_msresamp q) { return q->rate; }

void orc_msresamp_crcf_execute(orc_msresamp q, const orc_cf32 *x, unsigned nx, orc_cf32 *y, unsigned *ny)
{
    unsigned n = 0, M = 1u << q->S;
    if (!q->interp) {
        /* MSRESAMP(_decim_execute): buffer 2^S inputs -> half-band chain -> arbitrary resampler */
        for (unsigned i = 0; i < nx; i++) {
            q->buffer[q->buffer_index++] = x[i];
            if (q->buffer_index == M) {
                orc_cf32 hb = msresamp2_decim(&q->hb, q->buffer);
                n += resamp_exec(&q->arb, hb, &y[n]);
                q->buffer_index = 0;
            }
        }
    } else {
        /* MSRESAMP(_interp_execute): arbitrary resampler -> half-band interpolators */
        orc_cf32 tmp[8];
        for (unsigned i = 0; i < nx; i++) {
            unsigned nw = resamp_exec(&q->arb, x[i], tmp);
            for (unsigned j = 0; j < nw; j++) { msresamp2_interp(&q->hb, tmp[j], &y[n]); n += M; }
        }
    }
    *ny = n;
}
unsigned orc_msresamp_num_stages(orc_msresamp q) { return q->S; }
unsigned orc_msresamp_stage_m(orc_msresamp q, unsigned s) { return q->hb.m_stage[s]; }
const float *orc_msresamp_stage_h1(orc_msresamp q, unsigned s) { return q->hb.st[s].h1; }
float orc_msresamp_rate_arbitrary(orc_msresamp q) { return q->rate_arb; }
uint32_t orc_msresamp_resamp_step(orc_msresamp q) { return q->arb.step; }
unsigned orc_msresamp_resamp_npfb(orc_msresamp q) { return q->arb.npfb; }
const float *orc_msresamp_resamp_bank(orc_msresamp q) { return q->arb.bank; }

/* ============================================================================================
 * iirfilt_crcf dc blocker -- liquid src/filter/src/iirfilt.c ("normal" direct form II; known answer #3)
 * Reference: dcBlocker, Liquid.chs:575-589 (alpha = 0.0005)
 * ========================================================================================== */
struct orc_iirfilt_s { float b[2], a[2]; orc_cf32 v1; };
orc_iirfilt orc_iirfilt_crcf_create_dc_blocker(float alpha)
{
    orc_iirfilt q = (orc_iirfilt)calloc(1, sizeof(*q));
    q->b[0] = 1.0f; q->b[1] = -1.0f;
    q->a[0] = 1.0f; q->a[1] = -1.0f + alpha;
    return q;
}
void orc_iirfilt_crcf_destroy(orc_iirfilt q) { free(q); }
void orc_iirfilt_crcf_coeffs(orc_iirfilt q, float b[2], float a[2]) { b[0]=q->b[0]; b[1]=q->b[1]; a[0]=q->a[0]; a[1]=q->a[1]; }
void orc_iirfilt_crcf_execute_block(orc_iirfilt q, const orc_cf32 *x, unsigned n, orc_cf32 *y)
{
    /* IIRFILT(_execute_norm): v0 = x - a1*v1 ; y = b0*v0 + b1*v1 */
    for (unsigned i = 0; i < n; i++) {
        orc_cf32 v0 = cmk(x[i].re - q->a[1] * q->v1.re, x[i].im - q->a[1] * q->v1.im);
        y[i] = cmk(q->b[0] * v0.re + q->b[1] * q->v1.re, q->b[0] * v0.im + q->b[1] * q->v1.im);
        q->v1 = v0;
    }
}

/* ============================================================================================
 * firpfbch_crcf analyzer -- liquid src/multichannel/src/firpfbch.c
 * Reference: firpfbchCreate (kaiser, m=7, As=80), Liquid.chs:811-821
 * ========================================================================================== */
struct orc_firpfbch_s {
    unsigned M, p, h_len; float *h; float *hsub; /* [M][p], oldest first */
    win_cf *w; unsigned filter_index;
    orc_cf32 *X, *x; float *tw_re, *tw_im; int pow2;
};

static void dft_forward(orc_firpfbch q)
{
    unsigned M = q->M;
    if (q->pow2) {
        /* iterative radix-2 DIT, float32 butterflies, twiddles rounded from double */
        unsigned bits = 0; while ((1u << bits) < M) bits++;
        for (unsigned i = 0; i < M; i++) {
            unsigned r = 0; for (unsigned b = 0; b < bits; b++) if (i & (1u << b)) r |= 1u << (bits - 1 - b);
            q->x[r] = q->X[i];
        }
        for (unsigned len = 2; len <= M; len <<= 1) {
            unsigned half = len >> 1, stride = M / len;
            for (unsigned i = 0; i < M; i += len)
                for (unsigned j = 0; j < half; j++) {
                    float wr = q->tw_re[j * stride], wi = q->tw_im[j * stride];
                    orc_cf32 a = q->x[i + j], b = q->x[i + j + half];
                    float tr = b.re * wr - b.im * wi, ti = b.re * wi + b.im * wr;
                    q->x[i + j]        = cmk(a.re + tr, a.im + ti);
                    q->x[i + j + half] = cmk(a.re - tr, a.im - ti);
                }
        }
    } else {
        for (unsigned k = 0; k < M; k++) {
            double sr = 0.0, si = 0.0;
            for (unsigned n = 0; n < M; n++) {
                unsigned t = (unsigned)(((unsigned long long)k * n) % M);
                double wr = q->tw_re[t], wi = q->tw_im[t];
                sr += q->X[n].re * wr - q->X[n].im * wi;
                si += q->X[n].re * wi + q->X[n].im * wr;
            }
            q->x[k] = cmk((float)sr, (float)si);
        }
    }
}

orc_firpfbch orc_firpfbch_crcf_create_kaiser(int type, unsigned M, unsigned m, float As)
{
    if (type != 0 || M == 0 || m == 0) return NULL;  /* only LIQUID_ANALYZER (0) is on the reference path */
    orc_firpfbch q = (orc_firpfbch)calloc(1, sizeof(*q));
    q->M = M; q->p = 2 * m;
    unsigned h_len = 2 * M * m + 1;
    float *h = (float *)malloc(h_len * sizeof(float));
    orc_firdes_kaiser(h_len, 0.5f / (float)M, fabsf(As), 0.0f, h);
    q->h_len = M * q->p; q->h = h;
    q->hsub = (float *)malloc(M * q->p * sizeof(float));
    q->w = (win_cf *)calloc(M, sizeof(win_cf));
    for (unsigned i = 0; i < M; i++) {
        for (unsigned n = 0; n < q->p; n++) q->hsub[i * q->p + (q->p - n - 1)] = h[i + n * M];
        win_init(&q->w[i], q->p);
    }
    q->filter_index = M - 1;
    q->X = (orc_cf32 *)calloc(M, sizeof(orc_cf32));
    q->x = (orc_cf32 *)calloc(M, sizeof(orc_cf32));
    q->tw_re = (float *)malloc(M * sizeof(float)); q->tw_im = (float *)malloc(M * sizeof(float));
    for (unsigned t = 0; t < M; t++) {
        q->tw_re[t] = (float)cos(-2.0 * M_PI * (double)t / (double)M);
        q->tw_im[t] = (float)sin(-2.0 * M_PI * (double)t / (double)M);
    }
    q->pow2 = (M & (M - 1)) == 0 && M > 1;
    return q;
}
void orc_firpfbch_crcf_destroy(orc_firpfbch q)
{
    if (!q) return;
    for (unsigned i = 0; i < q->M; i++) win_free(&q->w[i]);
    free(q->w); free(q->h); free(q->hsub); free(q->X); free(q->x); free(q->tw_re); free(q->tw_im); free(q);
}
const float *orc_firpfbch_taps(orc_firpfbch q, unsigned *h_len) { if (h_len) *h_len = q->h_len; return q->h; }

void orc_firpfbch_crcf_analyzer_execute(orc_firpfbch q, const orc_cf32 *x, orc_cf32 *y)
{
    unsigned M = q->M;
    /* FIRPFBCH(_analyzer_push): x[0] -> w[M-1], x[1] -> w[M-2], ... */
    for (unsigned i = 0; i < M; i++) {
        win_push(&q->w[q->filter_index], x[i]);
        q->filter_index = (q->filter_index + M - 1) % M;
    }
    /* FIRPFBCH(_analyzer_run)(q, 0, y): X[M-1-i] = dot(h_sub_i, w[i]); forward DFT */
    for (unsigned i = 0; i < M; i++)
        q->X[M - i - 1] = dot_rc(q->hsub + i * q->p, q->w[i].v, q->p);
    dft_forward(q);
    memcpy(y, q->x, M * sizeof(orc_cf32));
}

/* ============================================================================================
 * firpfbch2_crcf analyzer -- liquid src/multichannel/src/firpfbch2.c (2x oversampled: M/2 samples in, M out).
 * NOT called by the reference (which uses firpfbch_crcf, Liquid.chs:730-742); restated because the task's
 * north star names it.  create_kaiser: 2 M m + 1 Kaiser taps at fc = 1/M (analyzer), scaled to
 * sum M; create: branch n = taps h[n + i M] reversed; execute_analyzer: the M/2 new samples go into windows
 * base-1 .. base-M/2 (base = M/2 on even frames, M on odd ones), branch i filters window (offset + i) mod M
 * (offset = 0 / M/2), backward DFT, scale 1/M.   Confidence M.
 * ========================================================================================== */
struct orc_firpfbch2_s { unsigned M, M2, p; float *h, *hsub; win_cf *w; int flag; orc_cf32 *X; };
orc_firpfbch2 orc_firpfbch2_crcf_create_kaiser(int type, unsigned M, unsigned m, float As)
{
    if (type != 0 || M < 2 || (M & 1) || m == 0) return NULL;
    orc_firpfbch2 q = (orc_firpfbch2)calloc(1, sizeof(*q));
    q->M = M; q->M2 = M / 2; q->p = 2 * m;
    unsigned h_len = 2 * M * m + 1;
    q->h = (float *)malloc(h_len * sizeof(float));
    orc_firdes_kaiser(h_len, 1.0f / (float)M, As, 0.0f, q->h);
    float sum = 0.0f;
    for (unsigned i = 0; i < h_len; i++) sum += q->h[i];
    for (unsigned i = 0; i < h_len; i++) q->h[i] = q->h[i] * (float)M / sum;
    q->hsub = (float *)malloc((size_t)M * q->p * sizeof(float));
    q->w = (win_cf *)calloc(M, sizeof(win_cf));
    for (unsigned n = 0; n < M; n++) {
        for (unsigned i = 0; i < q->p; i++) q->hsub[n * q->p + (q->p - i - 1)] = q->h[M * i + n];
        win_init(&q->w[n], q->p);
    }
    q->X = (orc_cf32 *)calloc(M, sizeof(orc_cf32));
    return q;
}
void orc_firpfbch2_crcf_destroy(orc_firpfbch2 q)
{
    if (!q) return;
    for (unsigned i = 0; i < q->M; i++) win_free(&q->w[i]);
    free(q->w); free(q->h); free(q->hsub); free(q->X); free(q);
}
const float *orc_firpfbch2_taps(orc_firpfbch2 q, unsigned *h_len) { if (h_len) *h_len = q->M * q->p; return q->h; }
void orc_firpfbch2_crcf_execute(orc_firpfbch2 q, const orc_cf32 *x, orc_cf32 *y)
{
    const unsigned M = q->M, M2 = q->M2;
    const unsigned base = q->flag ? M : M2, offset = q->flag ? M2 : 0;
    for (unsigned i = 0; i < M2; i++) win_push(&q->w[base - i - 1], x[i]);
    for (unsigned i = 0; i < M; i++) {
        unsigned b = (offset + i) % M;
        q->X[b] = dot_rc(q->hsub + i * q->p, q->w[b].v, q->p);
    }
    for (unsigned k = 0; k < M; k++) {
        double sr = 0.0, si = 0.0;
        for (unsigned n = 0; n < M; n++) {
            double a = 2.0 * M_PI * (double)(((unsigned long long)k * n) % M) / (double)M;
            double wr = cos(a), wi = sin(a);
            sr += q->X[n].re * wr - q->X[n].im * wi;
            si += q->X[n].re * wi + q->X[n].im * wr;
        }
        y[k] = cmk((float)sr / (float)M, (float)si / (float)M);
    }
    q->flag = 1 - q->flag;
}

/* ============================================================================================
 * agc_crcf -- liquid src/agc/src/agc.c.  Reference: agcCreate/agcExecuteBlock, Liquid.chs:693-717
 * ========================================================================================== */
enum { SQ_UNKNOWN = 0, SQ_ENABLED, SQ_RISE, SQ_SIGNALHI, SQ_FALL, SQ_SIGNALLO, SQ_TIMEOUT, SQ_DISABLED };
struct orc_agc_s {
    float g, scale, bandwidth, alpha, y2_prime; int is_locked;
    int squelch_mode; float squelch_threshold; unsigned squelch_timeout, squelch_timer;
};
orc_agc orc_agc_crcf_create(void)
{
    orc_agc q = (orc_agc)calloc(1, sizeof(*q));
    q->bandwidth = 1e-2f; q->alpha = q->bandwidth;
    q->g = 1.0f; q->y2_prime = 1.0f; q->is_locked = 0;
    q->squelch_mode = SQ_DISABLED; q->squelch_threshold = 0.0f; q->squelch_timeout = 100; q->squelch_timer = 0;
    q->scale = 1.0f;
    return q;
}
void orc_agc_crcf_destroy(orc_agc q) { free(q); }
void orc_agc_crcf_set_bandwidth(orc_agc q, float bt) { q->bandwidth = bt; q->alpha = bt; }
void orc_agc_crcf_set_signal_level(orc_agc q, float x2) { q->g = 1.0f / x2; q->y2_prime = 1.0f; }
void orc_agc_crcf_squelch_enable(orc_agc q) { q->squelch_mode = SQ_ENABLED; }
void orc_agc_crcf_squelch_set_threshold(orc_agc q, float t) { q->squelch_threshold = t; }
void orc_agc_crcf_squelch_set_timeout(orc_agc q, unsigned t) { q->squelch_timeout = t; }
float orc_agc_crcf_get_rssi(orc_agc q) { return (float)(-20 * log10((double)q->g)); }
float orc_agc_crcf_get_gain(orc_agc q) { return q->g; }
int orc_agc_crcf_squelch_get_status(orc_agc q) { return q->squelch_mode; }

static void agc_squelch_update(orc_agc q)
{
    int ex = orc_agc_crcf_get_rssi(q) > q->squelch_threshold;
    switch (q->squelch_mode) {
    case SQ_ENABLED:  q->squelch_mode = ex ? SQ_RISE : SQ_ENABLED; break;
    case SQ_RISE:     q->squelch_mode = ex ? SQ_SIGNALHI : SQ_FALL; break;
    case SQ_SIGNALHI: q->squelch_mode = ex ? SQ_SIGNALHI : SQ_FALL; break;
    case SQ_FALL:     q->squelch_mode = ex ? SQ_SIGNALHI : SQ_SIGNALLO; q->squelch_timer = q->squelch_timeout; break;
    case SQ_SIGNALLO:
        q->squelch_timer--;
        if (q->squelch_timer == 0) q->squelch_mode = SQ_TIMEOUT;
        else if (ex)               q->squelch_mode = SQ_SIGNALHI;
        break;
    case SQ_TIMEOUT:  q->squelch_mode = SQ_ENABLED; break;
    default: break;
    }
}
static inline orc_cf32 agc_execute1(orc_agc q, orc_cf32 x)
{
    orc_cf32 y = cmk(x.re * q->g, x.im * q->g);
    float y2 = y.re * y.re + y.im * y.im;
    q->y2_prime = (float)((1.0 - q->alpha) * q->y2_prime + q->alpha * y2);
    if (q->is_locked) return y;
    if (q->y2_prime > 1e-6f) q->g *= expf(-0.5f * q->alpha * logf(q->y2_prime));
    if (q->g > 1e6f) q->g = 1e6f;
    agc_squelch_update(q);
    return cmk(y.re * q->scale, y.im * q->scale);
}
void orc_agc_crcf_execute_block(orc_agc q, const orc_cf32 *x, unsigned n, orc_cf32 *y)
{
    for (unsigned i = 0; i < n; i++) y[i] = agc_execute1(q, x[i]);
}
/* Haskell agcExecuteBlock (Liquid.chs:693-705): execute 1 sample, zero it unless status == SIGNALHI(3) */
void orc_hs_agc_execute_block(orc_agc q, const orc_cf32 *x, unsigned n, orc_cf32 *y)
{
    for (unsigned i = 0; i < n; i++) {
        y[i] = agc_execute1(q, x[i]);
        if (q->squelch_mode != SQ_SIGNALHI) y[i] = cmk(0.0f, 0.0f);
    }
}

/* ============================================================================================
 * freqdem -- liquid src/modem/src/freqdem.c.  Reference: fmDemodulator kf, Liquid.chs:303-334
 * ========================================================================================== */
struct orc_freqdem_s { float kf, ref; orc_cf32 r_prime; };
orc_freqdem orc_freqdem_create(float kf)
{
    if (!(kf > 0.0f)) return NULL;
    orc_freqdem q = (orc_freqdem)calloc(1, sizeof(*q));
    q->kf = kf; q->ref = (float)(1.0f / (2 * M_PI * kf));
    return q;
}
void orc_freqdem_destroy(orc_freqdem q) { free(q); }
void orc_freqdem_demodulate_block(orc_freqdem q, const orc_cf32 *r, unsigned n, float *m)
{
    for (unsigned i = 0; i < n; i++) {
        /* cargf(conjf(r_prime) * r) * ref */
        float re = q->r_prime.re * r[i].re + q->r_prime.im * r[i].im;
        float im = q->r_prime.re * r[i].im - q->r_prime.im * r[i].re;
        m[i] = atan2f(im, re) * q->ref;
        q->r_prime = r[i];
    }
}

/* ============================================================================================
 * ampmodem -- liquid src/modem/src/ampmodem.c (3-argument create => the 2019 redesign), DSB only.
 * Reference: amdemodCreate = ampmodem_create 0.8 0 0, Liquid.chs:452-459.   Confidence LOW:
 * both candidate demodulators are implemented; ORC_OPT_AMPMODEM_PLL selects.
 * ========================================================================================== */
struct orc_ampmodem_s {
    float mod_index; int type, suppressed; unsigned m;
    float *h_dc;  float *w_dc;          /* 2m+1-tap real dc-blocker (firfilt_rrrf_create_dc_blocker(m,20)) */
    float *h_lp;  orc_cf32 *w_lp;       /* 2m+1-tap low-pass for carrier recovery (kaiser 0.01, 40 dB) */
    orc_cf32 *dly;                      /* m-sample delay line */
    orc_nco mixer;
};
orc_ampmodem orc_ampmodem_create(float mod_index, int type, int suppressed)
{
    if (type != 0 || suppressed != 0) return NULL;     /* only DSB with carrier is on the reference path */
    orc_ampmodem q = (orc_ampmodem)calloc(1, sizeof(*q));
    q->mod_index = mod_index; q->type = type; q->suppressed = suppressed; q->m = 25;
    unsigned n = 2 * q->m + 1;
    /* liquid_firdes_notch(m, 0, As): h = -w/sum(w) + delta[m] */
    q->h_dc = (float *)malloc(n * sizeof(float));
    double beta = (double)orc_kaiser_beta_As(20.0f), scale = 0.0;
    double *w = (double *)malloc(n * sizeof(double));
    for (unsigned i = 0; i < n; i++) { w[i] = kaiser_win(i, n, beta, 0.0); scale += w[i]; }
    for (unsigned i = 0; i < n; i++) q->h_dc[i] = (float)(-w[i] / scale);
    q->h_dc[q->m] += 1.0f;
    free(w);
    q->w_dc = (float *)calloc(n, sizeof(float));
    q->h_lp = (float *)malloc(n * sizeof(float));
    orc_firdes_kaiser(n, 0.01f, 40.0f, 0.0f, q->h_lp);
    q->w_lp = (orc_cf32 *)calloc(n, sizeof(orc_cf32));
    q->dly = (orc_cf32 *)calloc(q->m + 1, sizeof(orc_cf32));
    q->mixer = orc_nco_crcf_create(0);
    nco_pll_set_bandwidth(q->mixer, 0.001f);
    return q;
}
void orc_ampmodem_destroy(orc_ampmodem q)
{
    if (!q) return;
    free(q->h_dc); free(q->w_dc); free(q->h_lp); free(q->w_lp); free(q->dly); orc_nco_crcf_destroy(q->mixer); free(q);
}
static inline float am_dcblock(orc_ampmodem q, float v)
{
    unsigned n = 2 * q->m + 1;
    memmove(q->w_dc, q->w_dc + 1, (n - 1) * sizeof(float));
    q->w_dc[n - 1] = v;
    float acc = 0.0f;
    for (unsigned i = 0; i < n; i++) acc += q->h_dc[n - 1 - i] * q->w_dc[i];
    return acc;
}
void orc_ampmodem_demodulate_block(orc_ampmodem q, const orc_cf32 *r, unsigned nn, float *out)
{
    unsigned n = 2 * q->m + 1;
    for (unsigned k = 0; k < nn; k++) {
        if (!g_opt[ORC_OPT_AMPMODEM_PLL]) {
            /* ampmodem_demod_dsb_peak_detect: |x| -> dc block -> /mod_index */
            float t = hypotf(r[k].re, r[k].im);
            out[k] = am_dcblock(q, t) / q->mod_index;
        } else {
            /* ampmodem_demod_dsb_pll_carrier */
            memmove(q->w_lp, q->w_lp + 1, (n - 1) * sizeof(orc_cf32));
            q->w_lp[n - 1] = r[k];
            float xr = 0.0f, xi = 0.0f;
            for (unsigned i = 0; i < n; i++) { xr += q->h_lp[n - 1 - i] * q->w_lp[i].re; xi += q->h_lp[n - 1 - i] * q->w_lp[i].im; }
            memmove(q->dly, q->dly + 1, q->m * sizeof(orc_cf32));
            q->dly[q->m] = r[k];
            orc_cf32 x1 = q->dly[0];
            orc_cf32 v0 = nco_mix_down1(q->mixer, cmk(xr, xi));
            orc_cf32 v1 = nco_mix_down1(q->mixer, x1);
            float phase_error = v0.im;
            nco_pll_step(q->mixer, phase_error);
            orc_nco_crcf_step(q->mixer);
            float m = v1.re / q->mod_index;
            out[k] = am_dcblock(q, m);
        }
    }
}

/* ============================================================================================
 * Haskell firpfbchChan (Liquid.chs:827-862): pre-rotate the chunk, run nf = n/C frames, channel-major out
 * ========================================================================================== */
void orc_hs_firpfbch_chan(orc_firpfbch fb, orc_nco nco, unsigned C, const orc_cf32 *x, unsigned n, orc_cf32 *y)
{
    unsigned nf = n / C;
    orc_cf32 *dx = (orc_cf32 *)malloc((n ? n : 1) * sizeof(orc_cf32));
    orc_cf32 *tmp = (orc_cf32 *)malloc(C * sizeof(orc_cf32));
    orc_nco_crcf_mix_block_down(nco, x, dx, n);
    for (unsigned i = 0; i < nf; i++) {
        orc_firpfbch_crcf_analyzer_execute(fb, dx + (size_t)C * i, tmp);
        for (unsigned j = 0; j < C; j++) y[(size_t)nf * j + i] = tmp[j];
    }
    free(dx); free(tmp);
}

/* ==========================================================================================
 * iirfilt_rrrf from a Butterworth prototype, second-order sections -- liquid src/filter/src/iirdes.c
 * (liquid_iirdes, butter_azpkf, bilinear_zpkf, iirdes_dzpk2sosf), iirfilt.c (create_sos / execute_sos) and
 * iirfiltsos.c (direct form II).  Reference: iirfiltCreate = iirfilt_rrrf_create_prototype 0 0 0 n fc f0 ap as
 * (Liquid.chs:629-633: Butterworth, low-pass, SOS); wbFMDemodulator uses n = 2, fc = 5000/quadRate
 * (Liquid.chs:653-656).  Only that family (ftype 0, btype 0, format 0) is restated.   Confidence M.
 * ========================================================================================== */
typedef struct { float re, im; } cplx;
static cplx c_mk(float re, float im) { cplx z; z.re = re; z.im = im; return z; }
static cplx c_add(cplx a, cplx b) { return c_mk(a.re + b.re, a.im + b.im); }
static cplx c_sub(cplx a, cplx b) { return c_mk(a.re - b.re, a.im - b.im); }
static cplx c_mul(cplx a, cplx b) { return c_mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static cplx c_div(cplx a, cplx b)
{
    float d = b.re * b.re + b.im * b.im;
    return c_mk((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
#define ORC_IIR_MAX_SOS 8
struct orc_iirfilt_rrrf_s { unsigned nsos; float b[ORC_IIR_MAX_SOS][3], a[ORC_IIR_MAX_SOS][3], v[ORC_IIR_MAX_SOS][3]; };

orc_iirfilt_rrrf orc_iirfilt_rrrf_create_prototype(int ftype, int btype, int format, unsigned n, float fc, float f0,
                                                   float Ap, float As)
{
    (void)f0; (void)Ap; (void)As;
    if (ftype != 0 || btype != 0 || format != 0 || n == 0 || n > 2 * ORC_IIR_MAX_SOS || !(fc > 0.0f && fc < 0.5f)) return NULL;
    unsigned r = n % 2, L = (n - r) / 2, i, k = 0;
    /* butter_azpkf: poles on the unit circle of the left half plane, in conjugate pairs; no zeros; gain 1 */
    cplx pa[2 * ORC_IIR_MAX_SOS], pd[2 * ORC_IIR_MAX_SOS], zd[2 * ORC_IIR_MAX_SOS];
    for (i = 0; i < L; i++) {
        float theta = (float)(2 * (i + 1) + n - 1) * (float)M_PI / (float)(2 * n);
        pa[k++] = c_mk(cosf(theta), sinf(theta));
        pa[k++] = c_mk(cosf(theta), -sinf(theta));
    }
    if (r) pa[k++] = c_mk(-1.0f, 0.0f);
    /* iirdes_freqprewarp (low-pass): m = tan(pi fc);  bilinear_zpkf: z = (1 + m s)/(1 - m s), zeros at -1 */
    float m = tanf((float)M_PI * fc);
    cplx G = c_mk(1.0f, 0.0f);
    for (i = 0; i < n; i++) {
        cplx pm = c_mk(pa[i].re * m, pa[i].im * m);
        zd[i] = c_mk(-1.0f, 0.0f);
        pd[i] = c_div(c_add(c_mk(1.0f, 0.0f), pm), c_sub(c_mk(1.0f, 0.0f), pm));
        G = c_mul(G, c_div(c_sub(c_mk(1.0f, 0.0f), pd[i]), c_sub(c_mk(1.0f, 0.0f), zd[i])));
    }
    float kd = G.re;
    /* iirdes_dzpk2sosf: one section per conjugate pair (pairs already adjacent), a real pole last; the gain is
     * spread evenly over the sections' feed-forward coefficients */
    orc_iirfilt_rrrf q = (orc_iirfilt_rrrf)calloc(1, sizeof(*q));
    q->nsos = L + r;
    float kk = powf(kd, 1.0f / (float)(L + r));
    for (i = 0; i < L; i++) {
        cplx p0 = c_mk(-pd[2 * i].re, -pd[2 * i].im), p1 = c_mk(-pd[2 * i + 1].re, -pd[2 * i + 1].im);
        cplx z0 = c_mk(-zd[2 * i].re, -zd[2 * i].im), z1 = c_mk(-zd[2 * i + 1].re, -zd[2 * i + 1].im);
        q->a[i][0] = 1.0f; q->a[i][1] = c_add(p0, p1).re; q->a[i][2] = c_mul(p0, p1).re;
        q->b[i][0] = 1.0f; q->b[i][1] = c_add(z0, z1).re; q->b[i][2] = c_mul(z0, z1).re;
    }
    if (r) {
        q->a[L][0] = 1.0f; q->a[L][1] = -pd[n - 1].re; q->a[L][2] = 0.0f;
        q->b[L][0] = 1.0f; q->b[L][1] = -zd[n - 1].re; q->b[L][2] = 0.0f;
    }
    for (i = 0; i < L + r; i++) { q->b[i][0] *= kk; q->b[i][1] *= kk; q->b[i][2] *= kk; }
    return q;
}
void orc_iirfilt_rrrf_destroy(orc_iirfilt_rrrf q) { free(q); }
unsigned orc_iirfilt_rrrf_coeffs(orc_iirfilt_rrrf q, float *b, float *a)
{
    for (unsigned i = 0; i < q->nsos; i++) for (int j = 0; j < 3; j++) { b[3 * i + j] = q->b[i][j]; a[3 * i + j] = q->a[i][j]; }
    return q->nsos;
}
void orc_iirfilt_rrrf_execute_block(orc_iirfilt_rrrf q, const float *x, unsigned n, float *y)
{
    for (unsigned i = 0; i < n; i++) {
        float t = x[i];
        for (unsigned s = 0; s < q->nsos; s++) {
            /* iirfiltsos_execute_df2 */
            float *v = q->v[s], *b = q->b[s], *a = q->a[s];
            v[2] = v[1]; v[1] = v[0];
            v[0] = t - a[1] * v[1] - a[2] * v[2];
            t = b[0] * v[0] + b[1] * v[1] + b[2] * v[2];
        }
        y[i] = t;
    }
}

/* ==========================================================================================
 * firdecim_rrrf -- liquid src/filter/src/firdecim.c.  create_kaiser(M, m, As): h_len = 2 M m + 1,
 * fc = 0.5/M, Kaiser prototype; execute: push the M samples of a block, the output is the dot product taken
 * right after the FIRST of them went in.  Reference: firdecimCreate m = firdecim_rrrf_create_kaiser m 10 60
 * (Liquid.chs:487-492); firDecim runs length `div` m blocks per array (Liquid.chs:497-500).   Confidence M.
 * ========================================================================================== */
struct orc_firdecim_s { unsigned M, h_len; float *h, *w; };
orc_firdecim orc_firdecim_rrrf_create_kaiser(unsigned M, unsigned m, float As)
{
    if (M < 1 || m < 1) return NULL;
    orc_firdecim q = (orc_firdecim)calloc(1, sizeof(*q));
    q->M = M; q->h_len = 2 * M * m + 1;
    float *hf = (float *)calloc(q->h_len, sizeof(float));
    orc_firdes_kaiser(q->h_len, 0.5f / (float)M, As, 0.0f, hf);
    q->h = (float *)calloc(q->h_len, sizeof(float));
    for (unsigned i = 0; i < q->h_len; i++) q->h[i] = hf[q->h_len - i - 1];   /* firdecim_create: reversed */
    q->w = (float *)calloc(q->h_len, sizeof(float));                            /* window, oldest first */
    free(hf);
    return q;
}
void orc_firdecim_rrrf_destroy(orc_firdecim q) { if (q) { free(q->h); free(q->w); free(q); } }
const float *orc_firdecim_rrrf_taps(orc_firdecim q, unsigned *h_len) { *h_len = q->h_len; return q->h; }
void orc_firdecim_rrrf_execute_block(orc_firdecim q, const float *x, unsigned n, float *y)
{
    for (unsigned k = 0; k < n; k++) {
        for (unsigned i = 0; i < q->M; i++) {
            memmove(q->w, q->w + 1, (q->h_len - 1) * sizeof(float));
            q->w[q->h_len - 1] = x[(size_t)k * q->M + i];
            if (i == 0) {
                float r = 0.0f;
                for (unsigned j = 0; j < q->h_len; j++) r += q->h[j] * q->w[j];
                y[k] = r;
            }
        }
    }
}

/* ============================================================================================
 * The whole chain, apps/SoapySDR.hs:181-283.  Order: offset mix -> resampler -> dcBlocker ->
 * [channelizer ->] per-channel (agc -> demod) [-> mix sum].  `compact` only re-chunks, so the chain is
 * restated as a stream: whole frames of C samples are consumed as they become available.
 * ========================================================================================== */
struct orc_chain_s {
    orc_chain_cfg cfg; unsigned C, nout;
    orc_nco offset; int mix_up;
    orc_msresamp rs;
    orc_iirfilt dc;
    orc_firpfbch fb; orc_nco fb_nco;
    orc_firpfbch2 fb2;                               /* cfg.channelizer == 1 */
    orc_agc *agc; orc_freqdem *fm; orc_ampmodem *am;
    orc_iirfilt_rrrf *deemph; orc_firdecim *dec; float *dec_left; unsigned *dec_fill;   /* DeWBFM tail per channel */
    orc_cf32 *frame_buf; size_t frame_fill;          /* < C leftover samples */
};

orc_chain orc_chain_create(const orc_chain_cfg *cfg)
{
    orc_chain q = (orc_chain)calloc(1, sizeof(*q));
    q->cfg = *cfg; q->C = cfg->channels ? cfg->channels : 1;
    q->nout = (q->C > 1 && !cfg->mix) ? q->C : 1;
    /* f = 2*pi*offset/samplerate :: Float (SoapySDR.hs:205) */
    float f = (float)(2.0f * (float)M_PI * (float)cfg->offset_hz / (float)cfg->samplerate);
    if (f != 0.0f) {
        q->offset = orc_nco_crcf_create(1);
        q->mix_up = f < 0.0f;
        orc_nco_crcf_set_frequency(q->offset, q->mix_up ? -f : f);
    }
    if (cfg->bandwidth_hz != 0.0) q->rs = orc_msresamp_crcf_create((float)(cfg->bandwidth_hz / cfg->samplerate), 60.0f);
    q->dc = orc_iirfilt_crcf_create_dc_blocker(0.0005f);
    if (q->C > 1 && cfg->channelizer == 1) {
        q->fb2 = orc_firpfbch2_crcf_create_kaiser(0, q->C, 7, 80.0f);
        q->frame_buf = (orc_cf32 *)calloc(q->C, sizeof(orc_cf32));
    } else if (q->C > 1) {
        q->fb = orc_firpfbch_crcf_create_kaiser(0, q->C, 7, 80.0f);
        q->fb_nco = orc_nco_crcf_create(1);
        /* offset = -0.5*(n-1)/n*2*pi in Float (Liquid.chs:817) */
        float off = -(0.5f * ((float)q->C - 1.0f) / (float)q->C * 2.0f * (float)M_PI);
        orc_nco_crcf_set_frequency(q->fb_nco, off);
        q->frame_buf = (orc_cf32 *)calloc(q->C, sizeof(orc_cf32));
    }
    q->agc = (orc_agc *)calloc(q->C, sizeof(orc_agc));
    q->fm = (orc_freqdem *)calloc(q->C, sizeof(orc_freqdem));
    q->am = (orc_ampmodem *)calloc(q->C, sizeof(orc_ampmodem));
    q->deemph = (orc_iirfilt_rrrf *)calloc(q->C, sizeof(orc_iirfilt_rrrf));
    q->dec = (orc_firdecim *)calloc(q->C, sizeof(orc_firdecim));
    if (cfg->demod == 3) {
        q->cfg.decim = cfg->decim ? cfg->decim : 1;
        q->dec_left = (float *)calloc((size_t)q->C * q->cfg.decim, sizeof(float));
        q->dec_fill = (unsigned *)calloc(q->C, sizeof(unsigned));
    }
    for (unsigned c = 0; c < q->C; c++) {
        if (cfg->agc_thresh_db != 0.0f) {
            /* agcCreate, Liquid.chs:707-717 */
            q->agc[c] = orc_agc_crcf_create();
            orc_agc_crcf_set_bandwidth(q->agc[c], 0.1f);
            orc_agc_crcf_set_signal_level(q->agc[c], 1e-3f);
            orc_agc_crcf_squelch_enable(q->agc[c]);
            orc_agc_crcf_squelch_set_threshold(q->agc[c], cfg->agc_thresh_db);
            orc_agc_crcf_squelch_set_timeout(q->agc[c], 1000);
        }
        if (cfg->demod == 1) q->fm[c] = orc_freqdem_create(cfg->kf);
        if (cfg->demod == 2) q->am[c] = orc_ampmodem_create(0.8f, 0, 0);
        if (cfg->demod == 3) {
            /* wbFMDemodulator quadRate decim (Liquid.chs:652-656), quadRate = outBW (SoapySDR.hs:257):
             * firDecimator decim . iirFilter 2 (5000/quadRate) 0 10 10 . fmDemodulator 0.6 */
            double quad = cfg->bandwidth_hz != 0.0 ? cfg->bandwidth_hz : cfg->samplerate;
            q->fm[c] = orc_freqdem_create(0.6f);
            q->deemph[c] = orc_iirfilt_rrrf_create_prototype(0, 0, 0, 2, (float)(5000.0 / quad), 0.0f, 10.0f, 10.0f);
            q->dec[c] = orc_firdecim_rrrf_create_kaiser(q->cfg.decim, 10, 60.0f);
        }
    }
    return q;
}
void orc_chain_destroy(orc_chain q)
{
    if (!q) return;
    if (q->offset) orc_nco_crcf_destroy(q->offset);
    if (q->rs) orc_msresamp_crcf_destroy(q->rs);
    orc_iirfilt_crcf_destroy(q->dc);
    if (q->fb) { orc_firpfbch_crcf_destroy(q->fb); orc_nco_crcf_destroy(q->fb_nco); }
    if (q->fb2) orc_firpfbch2_crcf_destroy(q->fb2);
    for (unsigned c = 0; c < q->C; c++) {
        if (q->agc[c]) orc_agc_crcf_destroy(q->agc[c]);
        if (q->fm[c]) orc_freqdem_destroy(q->fm[c]);
        if (q->am[c]) orc_ampmodem_destroy(q->am[c]);
        if (q->deemph[c]) orc_iirfilt_rrrf_destroy(q->deemph[c]);
        if (q->dec[c]) orc_firdecim_rrrf_destroy(q->dec[c]);
    }
    free(q->deemph); free(q->dec); free(q->dec_left); free(q->dec_fill);
    free(q->agc); free(q->fm); free(q->am); free(q->frame_buf); free(q);
}
unsigned orc_chain_num_outputs(orc_chain q) { return q->nout; }

/* per-channel demod = (fm|am|wbfm|id) . agc  (SoapySDR.hs:236-272); returns the number of samples written.
 * DeWBFM restated as a stream: the decimator consumes whole blocks of `decim` samples as they become available
 * (the reference drops the remainder of every array, Liquid.chs:497-500, which ties its output to the chunking). */
static size_t chain_demod(orc_chain q, unsigned c, const orc_cf32 *x, unsigned n, orc_cf32 *tmp, void *out)
{
    const orc_cf32 *s = x;
    if (q->agc[c]) { orc_hs_agc_execute_block(q->agc[c], x, n, tmp); s = tmp; }
    if (q->cfg.demod == 1)      orc_freqdem_demodulate_block(q->fm[c], s, n, (float *)out);
    else if (q->cfg.demod == 2) orc_ampmodem_demodulate_block(q->am[c], s, n, (float *)out);
    else if (q->cfg.demod == 3) {
        const unsigned M = q->cfg.decim, fill = q->dec_fill[c];
        float *buf = (float *)malloc(((size_t)fill + n + 1) * sizeof(float));
        memcpy(buf, q->dec_left + (size_t)c * M, fill * sizeof(float));
        orc_freqdem_demodulate_block(q->fm[c], s, n, buf + fill);
        orc_iirfilt_rrrf_execute_block(q->deemph[c], buf + fill, n, buf + fill);
        const unsigned tot = fill + n, nb = tot / M;
        orc_firdecim_rrrf_execute_block(q->dec[c], buf, nb, (float *)out);
        q->dec_fill[c] = tot - nb * M;
        memcpy(q->dec_left + (size_t)c * M, buf + (size_t)nb * M, q->dec_fill[c] * sizeof(float));
        free(buf);
        return nb;
    }
    else                        memcpy(out, s, n * sizeof(orc_cf32));
    return n;
}

int orc_chain_process(orc_chain q, const orc_cf32 *x, size_t nx, void *const *outs, size_t cap, size_t *n_out)
{
    const size_t BLK = 1u << 16;
    size_t esz = q->cfg.demod ? sizeof(float) : sizeof(orc_cf32);
    size_t produced = 0;
    orc_cf32 *a = (orc_cf32 *)malloc(BLK * sizeof(orc_cf32));
    /* resampler output capacity, as the reference sizes it: 2*ceil(r*nx) (Liquid.chs:81-82) */
    size_t cap_b = q->rs ? (size_t)(2.0 * ceil((double)q->rs->rate * (double)BLK)) + 64 : BLK;
    orc_cf32 *b = (orc_cf32 *)malloc(cap_b * sizeof(orc_cf32));
    unsigned C = q->C;
    int rc = 0;
    for (size_t pos = 0; pos < nx && rc == 0; pos += BLK) {
        unsigned n = (unsigned)((nx - pos < BLK) ? nx - pos : BLK);
        const orc_cf32 *s = x + pos;
        if (q->offset) {
            if (q->mix_up) orc_nco_crcf_mix_block_up(q->offset, s, a, n);
            else           orc_nco_crcf_mix_block_down(q->offset, s, a, n);
            s = a;
        }
        unsigned nr = n;
        orc_cf32 *r = b;
        if (q->rs) orc_msresamp_crcf_execute(q->rs, s, n, b, &nr);
        else       memcpy(b, s, n * sizeof(orc_cf32));
        orc_iirfilt_crcf_execute_block(q->dc, r, nr, r);
        if (C == 1) {
            if (produced + nr > cap) { rc = -1; break; }
            orc_cf32 *tmp = (orc_cf32 *)malloc((nr ? nr : 1) * sizeof(orc_cf32));
            produced += chain_demod(q, 0, r, nr, tmp, (char *)outs[0] + produced * esz);
            free(tmp);
        } else {
            /* assemble whole frames: leftover + new (a frame of the oversampled analyzer is C/2 input samples) */
            const size_t hop = q->fb2 ? C / 2 : C;
            size_t tot = q->frame_fill + nr, nf = tot / hop;
            orc_cf32 *buf = (orc_cf32 *)malloc((tot ? tot : 1) * sizeof(orc_cf32));
            memcpy(buf, q->frame_buf, q->frame_fill * sizeof(orc_cf32));
            memcpy(buf + q->frame_fill, r, nr * sizeof(orc_cf32));
            if (produced + nf > cap) { free(buf); rc = -1; break; }
            if (nf) {
                orc_cf32 *ch = (orc_cf32 *)malloc(nf * C * sizeof(orc_cf32));
                orc_cf32 *tmp = (orc_cf32 *)malloc(nf * sizeof(orc_cf32));
                if (q->fb2) {
                    orc_cf32 *fr = (orc_cf32 *)malloc(C * sizeof(orc_cf32));
                    for (size_t t = 0; t < nf; t++) {
                        orc_firpfbch2_crcf_execute(q->fb2, buf + t * hop, fr);
                        for (unsigned c = 0; c < C; c++) ch[nf * c + t] = fr[c];
                    }
                    free(fr);
                } else
                orc_hs_firpfbch_chan(q->fb, q->fb_nco, C, buf, (unsigned)(nf * C), ch);
                void *dem = malloc(nf * sizeof(orc_cf32));
                size_t nd = 0;                       /* samples per channel after the demodulator (< nf for DeWBFM) */
                for (unsigned c = 0; c < C; c++) {
                    if (!q->cfg.mix) {
                        nd = chain_demod(q, c, ch + nf * c, (unsigned)nf, tmp, (char *)outs[c] + produced * esz);
                    } else {
                        /* mix = foldl1 (zipWith (+)) over channels 1..C (Trans.hs:119-122) */
                        nd = chain_demod(q, c, ch + nf * c, (unsigned)nf, tmp, dem);
                        char *o = (char *)outs[0] + produced * esz;
                        if (c == 0) memcpy(o, dem, nd * esz);
                        else if (q->cfg.demod) { float *of = (float *)o, *df = (float *)dem; for (size_t i = 0; i < nd; i++) of[i] = of[i] + df[i]; }
                        else { orc_cf32 *oc = (orc_cf32 *)o, *dc = (orc_cf32 *)dem; for (size_t i = 0; i < nd; i++) { oc[i].re = oc[i].re + dc[i].re; oc[i].im = oc[i].im + dc[i].im; } }
                    }
                }
                free(dem); free(ch); free(tmp);
                produced += nd;
            }
            q->frame_fill = tot - nf * hop;
            memcpy(q->frame_buf, buf + nf * hop, q->frame_fill * sizeof(orc_cf32));
            free(buf);
        }
    }
    free(a); free(b);
    *n_out = produced;
    return rc;
}
