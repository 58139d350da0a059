/*
 * liquid_compat.c -- libcsdr_liquid_compat.so: the liquid-dsp symbol names that src/ComposableSDR/Liquid.chs
 * imports for the receive chain, forwarded to libcsdr_b200.so.  Linking the reference with
 * `extra-libraries: SoapySDR, csdr_liquid_compat, csdr_b200, liquid` (composable-sdr.cabal:38) resolves these
 * imports to the GPU blocks with no source change; every other liquid symbol still comes from libliquid.
 * (Kept in a separate library so that libcsdr_b200.so can be loaded next to a real libliquid for A/B checks.)
 */
#define _GNU_SOURCE
#include "../../include/csdr_b200.h"
#include <dlfcn.h>
#include <stdio.h>

typedef csdr_cf32 cf;

/* A liquid family is either shadowed WHOLE -- every function the reference can reach on a handle of that type
 * (nco_crcf: create, print, set_frequency, pll_set_bandwidth, pll_step, step, cexpf, get_phase, set_phase, mix_block_*,
 * destroy; Liquid.chs:746-780) -- or, where liquid keeps a constructor of its own (iirfilt_crcf_create_prototype,
 * Liquid.chs:553-571, used by iirCFilter), the shared members look at the handle's tag word and pass foreign objects on
 * to the next definition in link order, i.e. the real libliquid. */
#define FORWARD_FOREIGN(q, name, proto, args)                                                        \
    if (csdr_handle_kind(q) == 0) {                                                                   \
        static void (*next) proto = 0;                                                                \
        if (!next) next = (void (*) proto)dlsym(RTLD_NEXT, name);                                     \
        if (next) next args;                                                                          \
        else fprintf(stderr, "csdr_liquid_compat: %s called with a foreign handle and no libliquid behind it\n", name); \
        return;                                                                                       \
    }

/* Liquid.chs:746-780 */
void *nco_crcf_create(int type) { return csdr_nco_crcf_create(type); }
void nco_crcf_destroy(void *q) { csdr_nco_crcf_destroy((csdr_nco)q); }
void nco_crcf_print(void *q) { csdr_nco_crcf_print((csdr_nco)q); }
void nco_crcf_set_frequency(void *q, float f) { csdr_nco_crcf_set_frequency((csdr_nco)q, f); }
void nco_crcf_set_phase(void *q, float p) { csdr_nco_crcf_set_phase((csdr_nco)q, p); }
void nco_crcf_adjust_frequency(void *q, float df) { csdr_nco_crcf_adjust_frequency((csdr_nco)q, df); }
void nco_crcf_adjust_phase(void *q, float dphi) { csdr_nco_crcf_adjust_phase((csdr_nco)q, dphi); }
void nco_crcf_step(void *q) { csdr_nco_crcf_step((csdr_nco)q); }
void nco_crcf_reset(void *q) { csdr_nco_crcf_reset((csdr_nco)q); }
float nco_crcf_get_phase(void *q) { return csdr_nco_crcf_get_phase((csdr_nco)q); }
float nco_crcf_get_frequency(void *q) { return csdr_nco_crcf_get_frequency((csdr_nco)q); }
void nco_crcf_cexpf(void *q, cf *y) { csdr_nco_crcf_cexpf((csdr_nco)q, y); }
void nco_crcf_pll_set_bandwidth(void *q, float bw) { csdr_nco_crcf_pll_set_bandwidth((csdr_nco)q, bw); }
void nco_crcf_pll_step(void *q, float dphi) { csdr_nco_crcf_pll_step((csdr_nco)q, dphi); }
void nco_crcf_mix_block_down(void *q, cf *x, cf *y, unsigned n) { csdr_nco_crcf_mix_block_down((csdr_nco)q, x, y, n); }
void nco_crcf_mix_block_up(void *q, cf *x, cf *y, unsigned n) { csdr_nco_crcf_mix_block_up((csdr_nco)q, x, y, n); }

/* Liquid.chs:58-73 */
void *msresamp_crcf_create(float r, float As) { return csdr_msresamp_crcf_create(r, As); }
void msresamp_crcf_destroy(void *q) { csdr_msresamp_crcf_destroy((csdr_msresamp)q); }
void msresamp_crcf_print(void *q) { csdr_msresamp_crcf_print((csdr_msresamp)q); }
float msresamp_crcf_get_rate(void *q) { return csdr_msresamp_crcf_get_rate((csdr_msresamp)q); }
void msresamp_crcf_execute(void *q, cf *x, unsigned nx, cf *y, unsigned *ny) { csdr_msresamp_crcf_execute((csdr_msresamp)q, x, nx, y, ny); }

/* Liquid.chs:550-567 */
void *iirfilt_crcf_create_dc_blocker(float alpha) { return csdr_iirfilt_crcf_create_dc_blocker(alpha); }
void iirfilt_crcf_destroy(void *q)
{
    FORWARD_FOREIGN(q, "iirfilt_crcf_destroy", (void *), (q))
    csdr_iirfilt_crcf_destroy((csdr_iirfilt)q);
}
void iirfilt_crcf_print(void *q)
{
    FORWARD_FOREIGN(q, "iirfilt_crcf_print", (void *), (q))
    csdr_iirfilt_crcf_print((csdr_iirfilt)q);
}
void iirfilt_crcf_execute_block(void *q, cf *x, unsigned n, cf *y)
{
    FORWARD_FOREIGN(q, "iirfilt_crcf_execute_block", (void *, cf *, unsigned, cf *), (q, x, n, y))
    csdr_iirfilt_crcf_execute_block((csdr_iirfilt)q, x, n, y);
}

/* Liquid.chs:732-742 */
void *firpfbch_crcf_create_kaiser(int type, unsigned M, unsigned m, float As) { return csdr_firpfbch_crcf_create_kaiser(type, M, m, As); }
void firpfbch_crcf_destroy(void *q) { csdr_firpfbch_crcf_destroy((csdr_firpfbch)q); }
void firpfbch_crcf_print(void *q) { csdr_firpfbch_crcf_print((csdr_firpfbch)q); }
void firpfbch_crcf_analyzer_execute(void *q, cf *x, cf *y) { csdr_firpfbch_crcf_analyzer_execute((csdr_firpfbch)q, x, y); }

/* Liquid.chs:660-691 */
void *agc_crcf_create(void) { return csdr_agc_crcf_create(); }
void agc_crcf_destroy(void *q) { csdr_agc_crcf_destroy((csdr_agc)q); }
void agc_crcf_print(void *q) { csdr_agc_crcf_print((csdr_agc)q); }
void agc_crcf_set_bandwidth(void *q, float bt) { csdr_agc_crcf_set_bandwidth((csdr_agc)q, bt); }
void agc_crcf_set_signal_level(void *q, float x2) { csdr_agc_crcf_set_signal_level((csdr_agc)q, x2); }
void agc_crcf_squelch_enable(void *q) { csdr_agc_crcf_squelch_enable((csdr_agc)q); }
void agc_crcf_squelch_set_threshold(void *q, float t) { csdr_agc_crcf_squelch_set_threshold((csdr_agc)q, t); }
void agc_crcf_squelch_set_timeout(void *q, unsigned t) { csdr_agc_crcf_squelch_set_timeout((csdr_agc)q, t); }
void agc_crcf_execute_block(void *q, cf *x, unsigned n, cf *y) { csdr_agc_crcf_execute_block((csdr_agc)q, x, n, y); }
float agc_crcf_get_rssi(void *q) { return csdr_agc_crcf_get_rssi((csdr_agc)q); }
int agc_crcf_squelch_get_status(void *q) { return csdr_agc_crcf_squelch_get_status((csdr_agc)q); }

/* Liquid.chs:305-315 */
void *freqdem_create(float kf) { return csdr_freqdem_create(kf); }
void freqdem_destroy(void *q) { csdr_freqdem_destroy((csdr_freqdem)q); }
void freqdem_print(void *q) { csdr_freqdem_print((csdr_freqdem)q); }
void freqdem_demodulate_block(void *q, cf *r, unsigned n, float *m) { csdr_freqdem_demodulate_block((csdr_freqdem)q, r, n, m); }

/* Liquid.chs:441-450 */
void *ampmodem_create(float mod_index, int type, int suppressed) { return csdr_ampmodem_create(mod_index, type, suppressed); }
void ampmodem_destroy(void *q) { csdr_ampmodem_destroy((csdr_ampmodem)q); }
void ampmodem_print(void *q) { csdr_ampmodem_print((csdr_ampmodem)q); }
void ampmodem_demodulate_block(void *q, cf *r, unsigned n, float *m) { csdr_ampmodem_demodulate_block((csdr_ampmodem)q, r, n, m); }

/* Liquid.chs:612-627 */
void *iirfilt_rrrf_create_prototype(int ftype, int btype, int format, unsigned order, float fc, float f0, float ap, float as)
{
    return csdr_iirfilt_rrrf_create_prototype(ftype, btype, format, order, fc, f0, ap, as);
}
void iirfilt_rrrf_destroy(void *q) { csdr_iirfilt_rrrf_destroy((csdr_iirfilt_rrrf)q); }
void iirfilt_rrrf_print(void *q) { csdr_iirfilt_rrrf_print((csdr_iirfilt_rrrf)q); }
void iirfilt_rrrf_execute_block(void *q, float *x, unsigned n, float *y) { csdr_iirfilt_rrrf_execute_block((csdr_iirfilt_rrrf)q, x, n, y); }

/* Liquid.chs:473-485 */
void *firdecim_rrrf_create_kaiser(unsigned M, unsigned m, float as) { return csdr_firdecim_rrrf_create_kaiser(M, m, as); }
void firdecim_rrrf_destroy(void *q) { csdr_firdecim_rrrf_destroy((csdr_firdecim)q); }
void firdecim_rrrf_print(void *q) { csdr_firdecim_rrrf_print((csdr_firdecim)q); }
void firdecim_rrrf_execute_block(void *q, float *x, unsigned n, float *y) { csdr_firdecim_rrrf_execute_block((csdr_firdecim)q, x, n, y); }

/* liquid's firpfbch2_crcf analyzer (not imported by the reference) */
void *firpfbch2_crcf_create_kaiser(int type, unsigned M, unsigned m, float As) { return csdr_firpfbch2_crcf_create_kaiser(type, M, m, As); }
void firpfbch2_crcf_destroy(void *q) { csdr_firpfbch2_crcf_destroy((csdr_firpfbch2)q); }
void firpfbch2_crcf_print(void *q) { csdr_firpfbch2_crcf_print((csdr_firpfbch2)q); }
void firpfbch2_crcf_execute(void *q, cf *x, cf *y) { csdr_firpfbch2_crcf_execute((csdr_firpfbch2)q, x, y); }
