/*
 * liquid_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH)
 *
 * Sequential float32 restatement of the liquid-dsp v1.3.2 objects that
 * mryndzionek/composable-sdr calls on its receive chain
 * (reference: src/ComposableSDR/Liquid.chs, apps/SoapySDR.hs:181-283).
 *
 * PARITY UNPINNED: liquid-dsp is an un-vendored third-party dependency of the
 * reference (pinned only by .github/workflows/build.yml:15 LIQUIDDSP_VER 1.3.2)
 * and neither its sources nor a binary exist in this environment; the reference
 * has no tests or golden vectors.  This file restates liquid's *published*
 * v1.3.2 algorithms from their documented behaviour.  What IS pinned: the four
 * known answers recovered from the reference's own screen capture
 * (images/ex1_5.gif, see tests/test_oracle_known_answers.py): the firpfbch
 * Kaiser prototype taps, the NCO frequency-word quantisation, the DC-blocker
 * coefficient form, and the per-channel output length invariant.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may link or call anything declared here.
 *
 * Naming: orc_<liquid object>_<method>, same argument order as liquid.
 */
#ifndef LIQUID_ORACLE_H
#define LIQUID_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } orc_cf32; /* == C99 float complex == Haskell Complex CFloat */

/* ---- oracle variant switches (documented in DESIGN.md "uncertain liquid details") ---- */
enum {
    ORC_OPT_VCO_DIRECT = 0,   /* 0 (default): LIQUID_VCO uses the 1024-entry sine table like LIQUID_NCO
                                 (v1.3.x behaviour as recollected); 1: sinf/cosf of the phase directly */
    ORC_OPT_AMPMODEM_PLL = 1, /* 1 (default): DSB non-suppressed demod = carrier PLL; 0: peak detector */
    ORC_OPT_RESAMP_FC_OLD = 2,/* 0 (default): msresamp arbitrary stage = resamp(rate,7,min(0.515*rate,0.49),As,256);
                                 1: older (rate,7,0.4,As,64) */
    ORC_OPT_COUNT = 3
};
void orc_set_option(int opt, int value);
int  orc_get_option(int opt);

/* ---- filter design helpers (liquid src/filter/src/firdes.c, src/math/src/windows.c) ---- */
float    orc_kaiser_beta_As(float As);
unsigned orc_estimate_req_filter_len(float df, float As);
void     orc_firdes_kaiser(unsigned n, float fc, float As, float mu, float *h);

/* ---- nco_crcf (liquid src/nco/src/nco.c)  [Liquid.chs:744-809] ---- */
typedef struct orc_nco_s *orc_nco;
orc_nco  orc_nco_crcf_create(int type);
void     orc_nco_crcf_destroy(orc_nco q);
void     orc_nco_crcf_set_frequency(orc_nco q, float dtheta);
void     orc_nco_crcf_set_phase(orc_nco q, float theta);
uint32_t orc_nco_crcf_get_phase_word(orc_nco q);
uint32_t orc_nco_crcf_get_freq_word(orc_nco q);
void     orc_nco_crcf_step(orc_nco q);
void     orc_nco_crcf_pll_set_bandwidth(orc_nco q, float bw);
void     orc_nco_crcf_pll_step(orc_nco q, float dphi);
float    orc_nco_crcf_get_phase(orc_nco q);
void     orc_nco_crcf_cexpf(orc_nco q, orc_cf32 *y);
void     orc_nco_crcf_mix_block_down(orc_nco q, const orc_cf32 *x, orc_cf32 *y, unsigned n);
void     orc_nco_crcf_mix_block_up(orc_nco q, const orc_cf32 *x, orc_cf32 *y, unsigned n);
const float *orc_nco_sintab(void);   /* the 1024-entry table, sinf(2*pi*i/1024) */

/* ---- msresamp_crcf and its parts (liquid src/filter/src/{msresamp,msresamp2,resamp2,resamp.fixed,firpfb}.c)
 *      [Liquid.chs:56-117] ---- */
typedef struct orc_msresamp_s *orc_msresamp;
orc_msresamp orc_msresamp_crcf_create(float r, float As);
void     orc_msresamp_crcf_destroy(orc_msresamp q);
float    orc_msresamp_crcf_get_rate(orc_msresamp q);
void     orc_msresamp_crcf_execute(orc_msresamp q, const orc_cf32 *x, unsigned nx, orc_cf32 *y, unsigned *ny);
/* introspection used by the parity tests and by csdr's create() cross-check */
unsigned orc_msresamp_num_stages(orc_msresamp q);
unsigned orc_msresamp_stage_m(orc_msresamp q, unsigned stage);            /* semi-length m of half-band stage */
const float *orc_msresamp_stage_h1(orc_msresamp q, unsigned stage);       /* 2m branch taps */
float    orc_msresamp_rate_arbitrary(orc_msresamp q);
uint32_t orc_msresamp_resamp_step(orc_msresamp q);
unsigned orc_msresamp_resamp_npfb(orc_msresamp q);
const float *orc_msresamp_resamp_bank(orc_msresamp q);                    /* [npfb][2m] newest-sample-first */

/* ---- iirfilt_crcf dc blocker (liquid src/filter/src/iirfilt.c)  [Liquid.chs:548-592] ---- */
typedef struct orc_iirfilt_s *orc_iirfilt;
orc_iirfilt orc_iirfilt_crcf_create_dc_blocker(float alpha);
void     orc_iirfilt_crcf_destroy(orc_iirfilt q);
void     orc_iirfilt_crcf_execute_block(orc_iirfilt q, const orc_cf32 *x, unsigned n, orc_cf32 *y);
void     orc_iirfilt_crcf_coeffs(orc_iirfilt q, float b[2], float a[2]);

/* ---- firpfbch_crcf analyzer (liquid src/multichannel/src/firpfbch.c)  [Liquid.chs:730-742, 811-866] ---- */
typedef struct orc_firpfbch_s *orc_firpfbch;
orc_firpfbch orc_firpfbch_crcf_create_kaiser(int type, unsigned M, unsigned m, float As);
void     orc_firpfbch_crcf_destroy(orc_firpfbch q);
void     orc_firpfbch_crcf_analyzer_execute(orc_firpfbch q, const orc_cf32 *x, orc_cf32 *y);
const float *orc_firpfbch_taps(orc_firpfbch q, unsigned *h_len);          /* prototype, 2*M*m used taps */

/* ---- agc_crcf (liquid src/agc/src/agc.c)  [Liquid.chs:658-728] ---- */
typedef struct orc_agc_s *orc_agc;
orc_agc  orc_agc_crcf_create(void);
void     orc_agc_crcf_destroy(orc_agc q);
void     orc_agc_crcf_set_bandwidth(orc_agc q, float bt);
void     orc_agc_crcf_set_signal_level(orc_agc q, float x2);
void     orc_agc_crcf_squelch_enable(orc_agc q);
void     orc_agc_crcf_squelch_set_threshold(orc_agc q, float thr_db);
void     orc_agc_crcf_squelch_set_timeout(orc_agc q, unsigned timeout);
void     orc_agc_crcf_execute_block(orc_agc q, const orc_cf32 *x, unsigned n, orc_cf32 *y);
float    orc_agc_crcf_get_rssi(orc_agc q);
float    orc_agc_crcf_get_gain(orc_agc q);
int      orc_agc_crcf_squelch_get_status(orc_agc q);

/* ---- freqdem (liquid src/modem/src/freqdem.c)  [Liquid.chs:303-334] ---- */
typedef struct orc_freqdem_s *orc_freqdem;
orc_freqdem orc_freqdem_create(float kf);
void     orc_freqdem_destroy(orc_freqdem q);
void     orc_freqdem_demodulate_block(orc_freqdem q, const orc_cf32 *r, unsigned n, float *m);

/* ---- ampmodem (liquid src/modem/src/ampmodem.c)  [Liquid.chs:439-469] ---- */
typedef struct orc_ampmodem_s *orc_ampmodem;
orc_ampmodem orc_ampmodem_create(float mod_index, int type, int suppressed_carrier);
void     orc_ampmodem_destroy(orc_ampmodem q);
void     orc_ampmodem_demodulate_block(orc_ampmodem q, const orc_cf32 *r, unsigned n, float *m);

/* ---- firpfbch2_crcf analyzer (2x oversampled; not called by the reference) ---- */
typedef struct orc_firpfbch2_s *orc_firpfbch2;
orc_firpfbch2 orc_firpfbch2_crcf_create_kaiser(int type, unsigned M, unsigned m, float As);   /* type 0 = analyzer, M even */
void     orc_firpfbch2_crcf_destroy(orc_firpfbch2 q);
const float *orc_firpfbch2_taps(orc_firpfbch2 q, unsigned *h_len);                            /* 2 M m used taps */
void     orc_firpfbch2_crcf_execute(orc_firpfbch2 q, const orc_cf32 *x /* M/2 */, orc_cf32 *y /* M */);

/* ---- iirfilt_rrrf (Butterworth low-pass prototype in second-order sections), Liquid.chs:610-633 ---- */
typedef struct orc_iirfilt_rrrf_s *orc_iirfilt_rrrf;
orc_iirfilt_rrrf orc_iirfilt_rrrf_create_prototype(int ftype, int btype, int format, unsigned n, float fc, float f0,
                                                   float Ap, float As);   /* NULL unless ftype = btype = format = 0 */
void     orc_iirfilt_rrrf_destroy(orc_iirfilt_rrrf q);
unsigned orc_iirfilt_rrrf_coeffs(orc_iirfilt_rrrf q, float *b, float *a);   /* [nsos][3] each; returns nsos */
void     orc_iirfilt_rrrf_execute_block(orc_iirfilt_rrrf q, const float *x, unsigned n, float *y);

/* ---- firdecim_rrrf, Liquid.chs:471-503 ---- */
typedef struct orc_firdecim_s *orc_firdecim;
orc_firdecim orc_firdecim_rrrf_create_kaiser(unsigned M, unsigned m, float As);
void     orc_firdecim_rrrf_destroy(orc_firdecim q);
const float *orc_firdecim_rrrf_taps(orc_firdecim q, unsigned *h_len);
void     orc_firdecim_rrrf_execute_block(orc_firdecim q, const float *x, unsigned n, float *y);   /* n blocks of M -> n */

/* ---- Haskell-side glue restated (reference src/ComposableSDR/Liquid.chs, Trans.hs) ---- */
/* agcExecuteBlock, Liquid.chs:693-705: per-sample execute + squelch gate (status != 3 -> 0) */
void     orc_hs_agc_execute_block(orc_agc q, const orc_cf32 *x, unsigned n, orc_cf32 *y);
/* firpfbchChan, Liquid.chs:827-862: pre-rotate whole chunk, nf = n / C frames, channel-major output [C][nf] */
void     orc_hs_firpfbch_chan(orc_firpfbch fb, orc_nco nco, unsigned C, const orc_cf32 *x, unsigned n, orc_cf32 *y);

/* ---- the whole receive chain, apps/SoapySDR.hs:181-283 (sdrProcess), as one sequential object ---- */
typedef struct {
    double   samplerate;      /* -s   */
    double   offset_hz;       /* --offset (Float in the reference) */
    double   bandwidth_hz;    /* -b, 0 = no resampler */
    int      demod;           /* 0 DeNo, 1 DeNBFM kf, 2 DeAM, 3 DeWBFM decim (kf 0.6, de-emphasis, decimator) */
    float    kf;
    float    agc_thresh_db;   /* -a, 0 = no AGC */
    unsigned channels;        /* -c */
    int      mix;             /* -m */
    unsigned decim;           /* DeWBFM: output decimation */
    int      channelizer;     /* 0: firpfbch_crcf + pre-rotation (the reference, Liquid.chs:811-866); 1: firpfbch2_crcf (2x
                                 oversampled analyzer, M/2 samples in, M channels out per frame, no pre-rotation) */
} orc_chain_cfg;
typedef struct orc_chain_s *orc_chain;
orc_chain orc_chain_create(const orc_chain_cfg *cfg);
void     orc_chain_destroy(orc_chain q);
/* Feed nx input samples.  outs[c] (c < channels, or 1 when mix/channels==1) receives the samples produced by
 * this call (float when demod != 0, cf32 otherwise); n_out = samples produced per output; cap = capacity of
 * every outs[c] in output samples.  Returns 0, or -1 if cap is too small. */
int      orc_chain_process(orc_chain q, const orc_cf32 *x, size_t nx, void *const *outs, size_t cap, size_t *n_out);
unsigned orc_chain_num_outputs(orc_chain q);

#ifdef __cplusplus
}
#endif
#endif
