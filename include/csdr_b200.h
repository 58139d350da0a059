/*
 * csdr_b200.h -- C ABI of libcsdr_b200.so: the B200 (sm_100a) receive chain of composable-sdr.
 *
 * Drop-in boundary.  The reference (mryndzionek/composable-sdr) reaches its DSP through GHC
 * `foreign import ccall unsafe` stubs in src/ComposableSDR/Liquid.chs that bind liquid-dsp symbols.  Each
 * entry point below names the reference import it replaces (file:line) and keeps liquid's argument order, so
 * a maintainer only changes the symbol string of the import (INTEGRATION.md shows the patched stubs).
 * The liquid-named aliases (nco_crcf_create, ...) live in libcsdr_liquid_compat.so.
 *
 * Conventions
 *   - plain C, no CUDA/torch types; handles are opaque pointers, NULL on failure (csdr_last_error() says why);
 *   - sample pointers may be HOST or DEVICE pointers: every call classifies them with
 *     cudaPointerGetAttributes and stages host memory itself.  The callee never retains a pointer;
 *   - csdr_cf32 is two float32 (I then Q) = Haskell `Complex CFloat` = C99 float complex;
 *   - one caller thread per handle, calls in stream order (the reference's RTS is non-threaded);
 *     different handles may be driven from different threads;
 *   - all stream state (NCO phase, FIR delay lines, resampler timing, AGC gain + squelch FSM, FM discriminator
 *     history) lives on the device inside the handle and is carried across calls: any chunking of the same
 *     input produces the same output stream;
 *   - there is no CPU fallback: without a CUDA device every create() fails.
 */
#ifndef CSDR_B200_H
#define CSDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } csdr_cf32;

/* ---------------------------------------------------------------- library ---- */
const char *csdr_version(void);
const char *csdr_last_error(void);                 /* thread-local, "" when the last call succeeded */
int         csdr_device_count(void);               /* number of visible CUDA devices (0 => nothing works) */
int         csdr_set_device(int device);           /* device used by subsequently created handles */
void       *csdr_host_alloc(size_t bytes);         /* pinned host memory (fast async copies); NULL on failure */
void        csdr_host_free(void *p);
uint64_t    csdr_kernel_launches(void);            /* kernels launched by this library so far (bench: gpu_launches) */
int         csdr_synchronize(void);                /* wait for all work queued by this library on the device */
/* options (before creating handles): */
enum {
    CSDR_OPT_VCO_DIRECT   = 0, /* 0: LIQUID_VCO phasor = 1024-level quantised phase (liquid 1.3.x table); 1: full phase */
    CSDR_OPT_AMPMODEM_PLL = 1, /* 1: DSB demod = carrier PLL; 0: peak detector */
    CSDR_OPT_RESAMP_FC_OLD= 2, /* 0: arbitrary stage (m=7, fc=min(.515 r,.49), npfb=256); 1: (7, 0.4, 64) */
    CSDR_OPT_AGC_SEGMENT  = 3, /* AGC time-segment length (samples), default 512 */
    CSDR_OPT_AGC_WARMUP   = 4, /* AGC warm-up length (samples), default 384 */
    CSDR_OPT_GENERIC_FRONTEND = 5, /* 1: always use the run-time-geometry front-end kernel (tests) */
    CSDR_OPT_AGC_EXACT_MATH = 6,  /* 1: library expf/logf/atan2f in the AGC/discriminator loop instead of the SFU forms */
    CSDR_OPT_OVERLAP = 7,         /* 1: overlap the back end of part i with the front end of part i+1 (2 streams) */
    CSDR_OPT_DEBUG = 8,           /* 1: print AGC speculation diagnostics to stderr (synchronises) */
    CSDR_OPT_PFB_VARIANT = 10,    /* channelizer kernel for M = 128..1024: 0 (default) one thread per polyphase branch, window in
                                     registers; 1 the ring-buffer kernel (cross-check); 2: M = 8, 16 on the one-frame-per-thread
                                     tile kernel instead of the two-frame one (cross-check); 3: firpfbch2 for M = 128..1024 as two launches
                                     (even / odd frames) instead of clusters of two CTAs (cross-check) */
    CSDR_OPT_AM_PLL_SEQUENTIAL = 11, /* 1: ampmodem's carrier loop runs as one sequential loop per lane (cross-check); 0 (default):
                                     time segments with a pull-in window, accepted within the loop's own quantisation noise */
    CSDR_OPT_FRONTEND_VARIANT = 9 /* front-end kernel for the standard half-band plan: 1 (default) raw tile by TMA tensor copy,
                                     read in place by the first half-band stage (3 CTAs/SM); 0 register prefetch + mixing
                                     pass (2 CTAs/SM); 2 as 1, warp-specialised (producer / consumer warp groups) */
};
/* 0 if h is not a live handle of this library, else its kind (1 nco, 2 msresamp, 3 iirfilt_crcf, 4 firpfbch, 5 firpfbch2,
 * 6 agc, 7 freqdem, 8 ampmodem, 9 iirfilt_rrrf, 10 firdecim, 11 chain).  Every entry point checks its handle this way: a
 * NULL or foreign pointer sets csdr_last_error() and returns instead of being dereferenced. */
int         csdr_handle_kind(const void *h);
int         csdr_set_option(int opt, int value);
int         csdr_get_option(int opt);

/* ---------------------------------------------------------------- nco_crcf ---- *
 * replaces Liquid.chs:746-780 (nco_crcf_create/_set_frequency/_mix_block_down/_mix_block_up/_print/_destroy) */
typedef struct csdr_nco_s *csdr_nco;
csdr_nco csdr_nco_crcf_create(int type /* 0 LIQUID_NCO, 1 LIQUID_VCO */);
void     csdr_nco_crcf_destroy(csdr_nco q);
void     csdr_nco_crcf_print(csdr_nco q);
void     csdr_nco_crcf_set_frequency(csdr_nco q, float dtheta);
void     csdr_nco_crcf_set_phase(csdr_nco q, float theta);
uint32_t csdr_nco_crcf_get_phase_word(csdr_nco q);
uint32_t csdr_nco_crcf_get_freq_word(csdr_nco q);
/* the rest of the family the reference binds (Liquid.chs:755-770: pll_set_bandwidth, pll_step, step, cexpf, get_phase --
 * the stereo-FM pilot PLL, Liquid.chs:959-988, drives them per sample on handles from nco_crcf_create) plus liquid's
 * other scalar members: host arithmetic on the handle's uint32 phase / frequency words (liquid nco.c) */
void     csdr_nco_crcf_adjust_frequency(csdr_nco q, float df);
void     csdr_nco_crcf_adjust_phase(csdr_nco q, float dphi);
void     csdr_nco_crcf_step(csdr_nco q);
void     csdr_nco_crcf_reset(csdr_nco q);
float    csdr_nco_crcf_get_phase(csdr_nco q);
float    csdr_nco_crcf_get_frequency(csdr_nco q);
void     csdr_nco_crcf_cexpf(csdr_nco q, csdr_cf32 *y);
void     csdr_nco_crcf_pll_set_bandwidth(csdr_nco q, float bw);
void     csdr_nco_crcf_pll_step(csdr_nco q, float dphi);
void     csdr_nco_crcf_mix_block_down(csdr_nco q, const csdr_cf32 *x, csdr_cf32 *y, unsigned n);
void     csdr_nco_crcf_mix_block_up(csdr_nco q, const csdr_cf32 *x, csdr_cf32 *y, unsigned n);

/* ---------------------------------------------------------------- msresamp_crcf ---- *
 * replaces Liquid.chs:58-73 (msresamp_crcf_create/_print/_get_rate/_execute/_destroy) */
typedef struct csdr_msresamp_s *csdr_msresamp;
csdr_msresamp csdr_msresamp_crcf_create(float r, float As);
void     csdr_msresamp_crcf_destroy(csdr_msresamp q);
void     csdr_msresamp_crcf_print(csdr_msresamp q);
float    csdr_msresamp_crcf_get_rate(csdr_msresamp q);
/* y must hold 2*ceil(r*nx) samples, exactly what the reference allocates (Liquid.chs:81-82) */
void     csdr_msresamp_crcf_execute(csdr_msresamp q, const csdr_cf32 *x, unsigned nx, csdr_cf32 *y, unsigned *ny);
/* derived design, for *_print and the parity tests */
unsigned csdr_msresamp_num_stages(csdr_msresamp q);
unsigned csdr_msresamp_stage_m(csdr_msresamp q, unsigned stage);
int      csdr_msresamp_stage_taps(csdr_msresamp q, unsigned stage, float *h1 /* 2m */);
uint32_t csdr_msresamp_resamp_step(csdr_msresamp q);
int      csdr_msresamp_resamp_bank(csdr_msresamp q, float *bank /* npfb*14 */, unsigned *npfb);

/* ---------------------------------------------------------------- iirfilt_crcf (dc blocker) ---- *
 * replaces Liquid.chs:550-567 (iirfilt_crcf_create_dc_blocker/_execute_block/_print/_destroy) */
typedef struct csdr_iirfilt_s *csdr_iirfilt;
csdr_iirfilt csdr_iirfilt_crcf_create_dc_blocker(float alpha);
void     csdr_iirfilt_crcf_destroy(csdr_iirfilt q);
void     csdr_iirfilt_crcf_print(csdr_iirfilt q);
void     csdr_iirfilt_crcf_execute_block(csdr_iirfilt q, const csdr_cf32 *x, unsigned n, csdr_cf32 *y);

/* ---------------------------------------------------------------- firpfbch_crcf ---- *
 * replaces Liquid.chs:732-742 (firpfbch_crcf_create_kaiser/_print/_analyzer_execute/_destroy) */
typedef struct csdr_firpfbch_s *csdr_firpfbch;
csdr_firpfbch csdr_firpfbch_crcf_create_kaiser(int type /* 0 analyzer */, unsigned M, unsigned m, float As);
void     csdr_firpfbch_crcf_destroy(csdr_firpfbch q);
void     csdr_firpfbch_crcf_print(csdr_firpfbch q);
void     csdr_firpfbch_crcf_analyzer_execute(csdr_firpfbch q, const csdr_cf32 *x /* M */, csdr_cf32 *y /* M */);
int      csdr_firpfbch_taps(csdr_firpfbch q, float *h /* 2*M*m */);
/* coarse entry point: replaces the per-frame FFI loop + per-element pokes of firpfbchChan (Liquid.chs:827-862).
 * Pre-rotates the chunk with `nco` (may be NULL), runs nframes = n / M frames and writes channel-major
 * y[M][nframes].  The n % M tail samples are consumed by the NCO and dropped, as the reference does. */
int      csdr_firpfbch_execute_block(csdr_firpfbch q, csdr_nco nco, const csdr_cf32 *x, unsigned n, csdr_cf32 *y);

/* ---------------------------------------------------------------- firpfbch2_crcf ---- *
 * liquid's 2x oversampled analyzer (M/2 samples in, M channels out per frame).  The reference does NOT import it
 * (its channelizer is firpfbch_crcf, Liquid.chs:730-742); offered as the alternative channelizer block the
 * task names, with liquid's signatures so that a future Liquid.chs import would bind it 1:1. */
typedef struct csdr_firpfbch2_s *csdr_firpfbch2;
csdr_firpfbch2 csdr_firpfbch2_crcf_create_kaiser(int type /* 0 analyzer */, unsigned M /* even */, unsigned m, float As);
void     csdr_firpfbch2_crcf_destroy(csdr_firpfbch2 q);
void     csdr_firpfbch2_crcf_print(csdr_firpfbch2 q);
void     csdr_firpfbch2_crcf_execute(csdr_firpfbch2 q, const csdr_cf32 *x /* M/2 */, csdr_cf32 *y /* M */);
int      csdr_firpfbch2_taps(csdr_firpfbch2 q, float *h /* 2*M*m */);
/* coarse entry point: nframes = n / (M/2) frames in one call, channel-major y[M][nframes] */
int      csdr_firpfbch2_execute_block(csdr_firpfbch2 q, const csdr_cf32 *x, unsigned n, csdr_cf32 *y);

/* ---------------------------------------------------------------- agc_crcf ---- *
 * replaces Liquid.chs:660-691 */
typedef struct csdr_agc_s *csdr_agc;
csdr_agc csdr_agc_crcf_create(void);
void     csdr_agc_crcf_destroy(csdr_agc q);
void     csdr_agc_crcf_print(csdr_agc q);
void     csdr_agc_crcf_set_bandwidth(csdr_agc q, float bt);
void     csdr_agc_crcf_set_signal_level(csdr_agc q, float x2);
void     csdr_agc_crcf_squelch_enable(csdr_agc q);
void     csdr_agc_crcf_squelch_set_threshold(csdr_agc q, float thr_db);
void     csdr_agc_crcf_squelch_set_timeout(csdr_agc q, unsigned timeout);
void     csdr_agc_crcf_execute_block(csdr_agc q, const csdr_cf32 *x, unsigned n, csdr_cf32 *y);
float    csdr_agc_crcf_get_rssi(csdr_agc q);
int      csdr_agc_crcf_squelch_get_status(csdr_agc q);
/* coarse entry point: the whole Haskell agcExecuteBlock loop (Liquid.chs:693-705) in one call:
 * execute, then y[i] = 0 unless squelch status == LIQUID_AGC_SQUELCH_SIGNALHI (3). */
int      csdr_agc_squelch_execute_block(csdr_agc q, const csdr_cf32 *x, unsigned n, csdr_cf32 *y);

/* ---------------------------------------------------------------- freqdem ---- *
 * replaces Liquid.chs:305-315 */
typedef struct csdr_freqdem_s *csdr_freqdem;
csdr_freqdem csdr_freqdem_create(float kf);
void     csdr_freqdem_destroy(csdr_freqdem q);
void     csdr_freqdem_print(csdr_freqdem q);
void     csdr_freqdem_demodulate_block(csdr_freqdem q, const csdr_cf32 *r, unsigned n, float *m);

/* ---------------------------------------------------------------- ampmodem ---- *
 * replaces Liquid.chs:441-450 */
typedef struct csdr_ampmodem_s *csdr_ampmodem;
csdr_ampmodem csdr_ampmodem_create(float mod_index, int type /* 0 DSB */, int suppressed_carrier /* 0 */);
void     csdr_ampmodem_destroy(csdr_ampmodem q);
void     csdr_ampmodem_print(csdr_ampmodem q);
void     csdr_ampmodem_demodulate_block(csdr_ampmodem q, const csdr_cf32 *r, unsigned n, float *m);

/* ---------------------------------------------------------------- iirfilt_rrrf (de-emphasis) ---- *
 * replaces Liquid.chs:610-627 (iirfilt_rrrf_create_prototype/_print/_execute_block/_destroy); the reference calls
 * create_prototype 0 0 0 n fc f0 ap as = Butterworth, low-pass, second-order sections (Liquid.chs:629-633), the only
 * family implemented (anything else: NULL + csdr_last_error). */
typedef struct csdr_iirfilt_rrrf_s *csdr_iirfilt_rrrf;
csdr_iirfilt_rrrf csdr_iirfilt_rrrf_create_prototype(int ftype, int btype, int format, unsigned order, float fc, float f0,
                                                     float ap, float as);
void     csdr_iirfilt_rrrf_destroy(csdr_iirfilt_rrrf q);
void     csdr_iirfilt_rrrf_print(csdr_iirfilt_rrrf q);
void     csdr_iirfilt_rrrf_execute_block(csdr_iirfilt_rrrf q, const float *x, unsigned n, float *y);
/* extension: the sections' coefficients, b and a as [sections][3]; returns the number of sections */
unsigned csdr_iirfilt_rrrf_coefficients(csdr_iirfilt_rrrf q, float *b, float *a);

/* ---------------------------------------------------------------- firdecim_rrrf (output decimator) ---- *
 * replaces Liquid.chs:471-485 (firdecim_rrrf_create_kaiser/_print/_execute_block/_destroy).
 * execute_block: n blocks of M input samples -> n output samples (Liquid.chs:497-500). */
typedef struct csdr_firdecim_s *csdr_firdecim;
csdr_firdecim csdr_firdecim_rrrf_create_kaiser(unsigned M, unsigned m, float as);
void     csdr_firdecim_rrrf_destroy(csdr_firdecim q);
void     csdr_firdecim_rrrf_print(csdr_firdecim q);
void     csdr_firdecim_rrrf_execute_block(csdr_firdecim q, const float *x, unsigned n, float *y);

/* ---------------------------------------------------------------- fused chain ---- *
 * The whole of sdrProcess (apps/SoapySDR.hs:181-283) behind one handle:
 *   offset mix -> msresamp(bw/sr, 60 dB) -> dcBlocker(5e-4) -> [firpfbch(C,7,80) ->] C x (agc -> demod) [-> mix]
 * One handle = `nstreams` independent streams with identical parameters (the reference would run one process
 * per stream). */
enum { CSDR_DEMOD_NONE = 0, CSDR_DEMOD_NBFM = 1, CSDR_DEMOD_AM = 2, CSDR_DEMOD_WBFM = 3 };
typedef struct {
    double   samplerate;     /* -s */
    double   offset_hz;      /* --offset */
    double   bandwidth_hz;   /* -b ; 0 = no resampler */
    int      demod;          /* CSDR_DEMOD_* */
    float    kf;             /* DeNBFM kf */
    float    agc_thresh_db;  /* -a ; 0 = no AGC */
    unsigned channels;       /* -c ; 0/1 = no channelizer */
    int      mix;            /* -m */
    unsigned nstreams;       /* 0/1 = single stream */
    int      device;         /* CUDA device ordinal, -1 = current */
    unsigned decim;          /* DeWBFM decim: wbFMDemodulator (kf 0.6, de-emphasis at 5 kHz of the quadrature rate = -b,
                                firdecim by decim; Liquid.chs:652-656, SoapySDR.hs:253-260); 0/1 = no decimation */
    int      channelizer;    /* CSDR_CHANNELIZER_*: 0 = firpfbch_crcf + pre-rotation, what the reference runs (Liquid.chs:811-866);
                                1 = firpfbch2_crcf, liquid's 2x oversampled analyzer (channels even; a frame is channels/2
                                samples, every channel comes out at 2/channels of the input rate, channel c centred on
                                c/channels of the sample rate, no pre-rotation) */
} csdr_chain_cfg;
enum { CSDR_CHANNELIZER_FIRPFBCH = 0, CSDR_CHANNELIZER_FIRPFBCH2 = 1 };
typedef struct csdr_chain_s *csdr_chain;
csdr_chain csdr_chain_create(const csdr_chain_cfg *cfg);
int      csdr_chain_destroy(csdr_chain q);
void     csdr_chain_print(csdr_chain q);
unsigned csdr_chain_num_outputs(csdr_chain q);      /* per stream: C if channels>1 && !mix, else 1 */
size_t   csdr_chain_out_elem_size(csdr_chain q);    /* 4 (float) if demod != NONE else 8 (cf32) */
size_t   csdr_chain_max_output(csdr_chain q, size_t nx); /* upper bound of samples per output for nx inputs */
/* Feed nx samples per stream.  x: stream s at x + s*x_stride (samples).  outs[s*num_outputs + c]: output c of
 * stream s, capacity out_cap samples each.  *n_out: samples written to every output by this call.
 * Host and device pointers are both accepted (host buffers from csdr_host_alloc are copied asynchronously,
 * overlapped with compute).  Returns 0 on success. */
int      csdr_chain_process(csdr_chain q, const csdr_cf32 *x, size_t nx, size_t x_stride,
                            void *const *outs, size_t out_cap, size_t *n_out);
/* File to file(s): the source and sinks either side of the path as apps/SoapySDR.hs wires them --
 * readFromFile (Source.chs:259-271: raw interleaved little-endian float32 I/Q, arrays of `chunk` samples; 0 = 2^24),
 * takeNArr numsamples behind the resampler (SoapySDR.hs:207; 0 = the whole file), fileSink (Sink.hs:29-34) named
 * <out_name>.cf32, or <out_name>_ch<K>.cf32 (K = 1..C) behind the channelizer without --mix (SoapySDR.hs:222-240).
 * Demodulated outputs are raw float32 files (.f32) instead of the reference's libsndfile AU/WAV containers.
 * Reading, the chain and writing overlap (pinned staging buffers).  *n_in: input samples consumed, *n_out: samples
 * written per output file. */
int      csdr_chain_run_file(csdr_chain q, const char *in_path, const char *out_name, uint64_t numsamples, size_t chunk,
                             uint64_t *n_in, uint64_t *n_out);
/* Seed the stream position for time-segment sharding: declare that `n_prior` input samples precede the next
 * call (NCO phase, half-band block alignment, resampler timing and -- behind a channelizer -- the pre-rotation phase
 * and the frame grid are closed-form in the sample index).  The caller feeds csdr_chain_warmup_len() samples of real
 * history first and discards the outputs they produce.  With a channelizer the shard should start on a frame boundary
 * of the stream (resampler output index = 0 mod C; 0 mod C/2 for the firpfbch2 channelizer, whose frame parity follows the
 * absolute frame index), so that every frame is produced by exactly one shard.  DeWBFM: the output decimator's block grid
 * follows the absolute index of the demodulated samples as well. */
int      csdr_chain_seek(csdr_chain q, uint64_t n_prior);
size_t   csdr_chain_warmup_len(csdr_chain q);
/* the CUDA stream (cudaStream_t) the chain launches on, for event timing by the caller */
void    *csdr_chain_cuda_stream(csdr_chain q);
/* event timing of the dominant kernel (k_frontend: mix + msresamp) on the chain's own stream: enable, run, then
 * read the accumulated device time [ms] and launch count (bench.py roofline) */
int      csdr_chain_profile(csdr_chain q, int enable);
double   csdr_chain_frontend_ms(csdr_chain q, uint64_t *launches);
/* number of AGC time segments the last call had to recompute sequentially (speculation misses) */
uint64_t csdr_chain_agc_fixups(csdr_chain q);
/* cumulative counters: [0] gain-loop segments repaired in order, [1] squelch-FSM segments repaired in order,
 * [2] gain-loop segments refined in parallel */
int      csdr_chain_agc_counters(csdr_chain q, uint64_t out[3]);
/* segment length and warm-up length (samples) the gain-loop speculation used in the last call (self-tuning) */
int      csdr_chain_agc_plan(csdr_chain q, int out[2]);

#ifdef __cplusplus
}
#endif
#endif
