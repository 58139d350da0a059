"""ctypes binding of the CPU oracle (oracle/liquid_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never from the composable-sdr_b200 package.  PARITY UNPINNED (see liquid_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_LIB_FAST = _LIB      # timing uses the same -O2 build (the fastest of the flag sets tried, see Makefile)

OPT_VCO_DIRECT, OPT_AMPMODEM_PLL, OPT_RESAMP_FC_OLD = 0, 1, 2


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("liquid_oracle.c", "liquid_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB


class ChainCfg(C.Structure):
    _fields_ = [("samplerate", C.c_double), ("offset_hz", C.c_double), ("bandwidth_hz", C.c_double),
                ("demod", C.c_int), ("kf", C.c_float), ("agc_thresh_db", C.c_float),
                ("channels", C.c_uint), ("mix", C.c_int), ("decim", C.c_uint), ("channelizer", C.c_int)]


_lib = None
_lib_fast = None


def lib(fast=False):
    global _lib, _lib_fast
    if fast:
        if _lib_fast is None:
            build()
            _lib_fast = _declare(C.CDLL(_LIB_FAST))
        return _lib_fast
    if _lib is None:
        _lib = _declare(C.CDLL(build()))
    return _lib


def _declare(L):
    vp, u, f, i = C.c_void_p, C.c_uint, C.c_float, C.c_int
    sig = {
        "orc_set_option": (None, [i, i]), "orc_get_option": (i, [i]),
        "orc_kaiser_beta_As": (f, [f]), "orc_estimate_req_filter_len": (u, [f, f]),
        "orc_firdes_kaiser": (None, [u, f, f, f, vp]),
        "orc_nco_crcf_create": (vp, [i]), "orc_nco_crcf_destroy": (None, [vp]),
        "orc_nco_crcf_set_frequency": (None, [vp, f]), "orc_nco_crcf_set_phase": (None, [vp, f]),
        "orc_nco_crcf_get_phase_word": (C.c_uint32, [vp]), "orc_nco_crcf_get_freq_word": (C.c_uint32, [vp]),
        "orc_nco_crcf_step": (None, [vp]), "orc_nco_crcf_pll_set_bandwidth": (None, [vp, f]),
        "orc_nco_crcf_pll_step": (None, [vp, f]), "orc_nco_crcf_get_phase": (f, [vp]), "orc_nco_crcf_cexpf": (None, [vp, vp]),
        "orc_nco_crcf_mix_block_down": (None, [vp, vp, vp, u]), "orc_nco_crcf_mix_block_up": (None, [vp, vp, vp, u]),
        "orc_nco_sintab": (C.POINTER(C.c_float), []),
        "orc_firpfbch2_crcf_create_kaiser": (vp, [i, u, u, f]), "orc_firpfbch2_crcf_destroy": (None, [vp]),
        "orc_firpfbch2_taps": (C.POINTER(C.c_float), [vp, C.POINTER(u)]),
        "orc_firpfbch2_crcf_execute": (None, [vp, vp, vp]),
        "orc_iirfilt_rrrf_create_prototype": (vp, [i, i, i, u, f, f, f, f]), "orc_iirfilt_rrrf_destroy": (None, [vp]),
        "orc_iirfilt_rrrf_coeffs": (u, [vp, vp, vp]), "orc_iirfilt_rrrf_execute_block": (None, [vp, vp, u, vp]),
        "orc_firdecim_rrrf_create_kaiser": (vp, [u, u, f]), "orc_firdecim_rrrf_destroy": (None, [vp]),
        "orc_firdecim_rrrf_taps": (C.POINTER(C.c_float), [vp, C.POINTER(u)]),
        "orc_firdecim_rrrf_execute_block": (None, [vp, vp, u, vp]),
        "orc_msresamp_crcf_create": (vp, [f, f]), "orc_msresamp_crcf_destroy": (None, [vp]),
        "orc_msresamp_crcf_get_rate": (f, [vp]),
        "orc_msresamp_crcf_execute": (None, [vp, vp, u, vp, C.POINTER(u)]),
        "orc_msresamp_num_stages": (u, [vp]), "orc_msresamp_stage_m": (u, [vp, u]),
        "orc_msresamp_stage_h1": (C.POINTER(C.c_float), [vp, u]),
        "orc_msresamp_rate_arbitrary": (f, [vp]), "orc_msresamp_resamp_step": (C.c_uint32, [vp]),
        "orc_msresamp_resamp_npfb": (u, [vp]), "orc_msresamp_resamp_bank": (C.POINTER(C.c_float), [vp]),
        "orc_iirfilt_crcf_create_dc_blocker": (vp, [f]), "orc_iirfilt_crcf_destroy": (None, [vp]),
        "orc_iirfilt_crcf_execute_block": (None, [vp, vp, u, vp]),
        "orc_iirfilt_crcf_coeffs": (None, [vp, C.POINTER(f), C.POINTER(f)]),
        "orc_firpfbch_crcf_create_kaiser": (vp, [i, u, u, f]), "orc_firpfbch_crcf_destroy": (None, [vp]),
        "orc_firpfbch_crcf_analyzer_execute": (None, [vp, vp, vp]),
        "orc_firpfbch_taps": (C.POINTER(C.c_float), [vp, C.POINTER(u)]),
        "orc_agc_crcf_create": (vp, []), "orc_agc_crcf_destroy": (None, [vp]),
        "orc_agc_crcf_set_bandwidth": (None, [vp, f]), "orc_agc_crcf_set_signal_level": (None, [vp, f]),
        "orc_agc_crcf_squelch_enable": (None, [vp]), "orc_agc_crcf_squelch_set_threshold": (None, [vp, f]),
        "orc_agc_crcf_squelch_set_timeout": (None, [vp, u]),
        "orc_agc_crcf_execute_block": (None, [vp, vp, u, vp]), "orc_agc_crcf_get_rssi": (f, [vp]),
        "orc_agc_crcf_get_gain": (f, [vp]), "orc_agc_crcf_squelch_get_status": (i, [vp]),
        "orc_freqdem_create": (vp, [f]), "orc_freqdem_destroy": (None, [vp]),
        "orc_freqdem_demodulate_block": (None, [vp, vp, u, vp]),
        "orc_ampmodem_create": (vp, [f, i, i]), "orc_ampmodem_destroy": (None, [vp]),
        "orc_ampmodem_demodulate_block": (None, [vp, vp, u, vp]),
        "orc_hs_agc_execute_block": (None, [vp, vp, u, vp]),
        "orc_hs_firpfbch_chan": (None, [vp, vp, u, vp, u, vp]),
        "orc_chain_create": (vp, [C.POINTER(ChainCfg)]), "orc_chain_destroy": (None, [vp]),
        "orc_chain_process": (i, [vp, vp, C.c_size_t, C.POINTER(vp), C.c_size_t, C.POINTER(C.c_size_t)]),
        "orc_chain_num_outputs": (u, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cf(x):
    return np.ascontiguousarray(x, dtype=np.complex64)


def set_option(opt, value):
    lib().orc_set_option(opt, int(value))


def firdes_kaiser(n, fc, As, mu=0.0):
    h = np.empty(n, np.float32)
    lib().orc_firdes_kaiser(n, fc, As, mu, _p(h))
    return h


def sintab():
    return np.ctypeslib.as_array(lib().orc_nco_sintab(), shape=(1024,)).copy()


class Nco:
    """liquid nco_crcf (reference: ncoCreate, Liquid.chs:782-789)."""

    def __init__(self, freq, nco_type=1):
        self.L = lib()
        self.h = self.L.orc_nco_crcf_create(nco_type)
        self.L.orc_nco_crcf_set_frequency(self.h, freq)

    def close(self):
        if self.h:
            self.L.orc_nco_crcf_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def freq_word(self):
        return self.L.orc_nco_crcf_get_freq_word(self.h)

    @property
    def phase_word(self):
        return self.L.orc_nco_crcf_get_phase_word(self.h)

    def mix_down(self, x):
        x = _cf(x)
        y = np.empty_like(x)
        self.L.orc_nco_crcf_mix_block_down(self.h, _p(x), _p(y), x.size)
        return y

    def mix_up(self, x):
        x = _cf(x)
        y = np.empty_like(x)
        self.L.orc_nco_crcf_mix_block_up(self.h, _p(x), _p(y), x.size)
        return y


class MsResamp:
    """liquid msresamp_crcf (reference: resampler r as, Liquid.chs:76-117)."""

    def __init__(self, rate, As=60.0):
        self.L = lib()
        self.h = self.L.orc_msresamp_crcf_create(rate, As)
        if not self.h:
            raise ValueError("msresamp_crcf_create failed")
        self.rate = self.L.orc_msresamp_crcf_get_rate(self.h)

    def close(self):
        if self.h:
            self.L.orc_msresamp_crcf_destroy(self.h)
            self.h = None

    __del__ = close

    def execute(self, x):
        x = _cf(x)
        cap = 2 * int(np.ceil(self.rate * x.size)) + 64          # reference: 2*ceil(r*nx), Liquid.chs:81-82
        y = np.empty(cap, np.complex64)
        ny = C.c_uint(0)
        self.L.orc_msresamp_crcf_execute(self.h, _p(x), x.size, _p(y), C.byref(ny))
        assert ny.value <= cap
        return y[:ny.value].copy()

    def design(self):
        L, h = self.L, self.h
        S = L.orc_msresamp_num_stages(h)
        ms = [L.orc_msresamp_stage_m(h, s) for s in range(S)]
        h1 = [np.ctypeslib.as_array(L.orc_msresamp_stage_h1(h, s), shape=(2 * ms[s],)).copy() for s in range(S)]
        npfb = L.orc_msresamp_resamp_npfb(h)
        bank = np.ctypeslib.as_array(L.orc_msresamp_resamp_bank(h), shape=(npfb, 14)).copy()
        return dict(S=S, m=ms, h1=h1, rate_arb=L.orc_msresamp_rate_arbitrary(h),
                    step=L.orc_msresamp_resamp_step(h), npfb=npfb, bank=bank)


class DcBlocker:
    """liquid iirfilt_crcf_create_dc_blocker (reference: dcBlocker, Liquid.chs:575-589)."""

    def __init__(self, alpha=0.0005):
        self.L = lib()
        self.h = self.L.orc_iirfilt_crcf_create_dc_blocker(alpha)

    def close(self):
        if self.h:
            self.L.orc_iirfilt_crcf_destroy(self.h)
            self.h = None

    __del__ = close

    def coeffs(self):
        b = (C.c_float * 2)()
        a = (C.c_float * 2)()
        self.L.orc_iirfilt_crcf_coeffs(self.h, b, a)
        return list(b), list(a)

    def execute(self, x):
        x = _cf(x)
        y = np.empty_like(x)
        self.L.orc_iirfilt_crcf_execute_block(self.h, _p(x), x.size, _p(y))
        return y


class Firpfbch:
    """firpfbchChannelizer n (Liquid.chs:811-866): kaiser(m=7, As=80) analyzer + pre-rotation NCO."""

    def __init__(self, nch, m=7, As=80.0):
        self.L = lib()
        self.C = nch
        self.h = self.L.orc_firpfbch_crcf_create_kaiser(0, nch, m, As)
        off = np.float32(-(np.float32(0.5) * (np.float32(nch) - np.float32(1)) / np.float32(nch)
                           * np.float32(2) * np.float32(np.pi)))
        self.nco = Nco(float(off), 1)

    def close(self):
        if self.h:
            self.L.orc_firpfbch_crcf_destroy(self.h)
            self.h = None

    __del__ = close

    def taps(self):
        n = C.c_uint(0)
        p = self.L.orc_firpfbch_taps(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def execute(self, x):
        """x: n samples -> [C][n // C] channel-major (remainder samples are dropped, as in the reference)."""
        x = _cf(x)
        nf = x.size // self.C
        y = np.empty((self.C, nf), np.complex64)
        self.L.orc_hs_firpfbch_chan(self.h, self.nco.h, self.C, _p(x), x.size, _p(y))
        return y

    def analyzer_execute(self, frame):
        frame = _cf(frame)
        y = np.empty(self.C, np.complex64)
        self.L.orc_firpfbch_crcf_analyzer_execute(self.h, _p(frame), _p(y))
        return y


class Agc:
    """automaticGainControl tres (Liquid.chs:693-728)."""

    def __init__(self, thresh_db, bw=0.1, level=1e-3, timeout=1000):
        L = self.L = lib()
        self.h = L.orc_agc_crcf_create()
        L.orc_agc_crcf_set_bandwidth(self.h, bw)
        L.orc_agc_crcf_set_signal_level(self.h, level)
        L.orc_agc_crcf_squelch_enable(self.h)
        L.orc_agc_crcf_squelch_set_threshold(self.h, thresh_db)
        L.orc_agc_crcf_squelch_set_timeout(self.h, timeout)

    def close(self):
        if self.h:
            self.L.orc_agc_crcf_destroy(self.h)
            self.h = None

    __del__ = close

    def execute(self, x):
        """Haskell agcExecuteBlock: gated output."""
        x = _cf(x)
        y = np.empty_like(x)
        self.L.orc_hs_agc_execute_block(self.h, _p(x), x.size, _p(y))
        return y

    def execute_raw(self, x):
        x = _cf(x)
        y = np.empty_like(x)
        self.L.orc_agc_crcf_execute_block(self.h, _p(x), x.size, _p(y))
        return y

    @property
    def gain(self):
        return self.L.orc_agc_crcf_get_gain(self.h)

    @property
    def rssi(self):
        return self.L.orc_agc_crcf_get_rssi(self.h)

    @property
    def status(self):
        return self.L.orc_agc_crcf_squelch_get_status(self.h)


class FreqDem:
    def __init__(self, kf):
        self.L = lib()
        self.h = self.L.orc_freqdem_create(kf)

    def close(self):
        if self.h:
            self.L.orc_freqdem_destroy(self.h)
            self.h = None

    __del__ = close

    def execute(self, x):
        x = _cf(x)
        y = np.empty(x.size, np.float32)
        self.L.orc_freqdem_demodulate_block(self.h, _p(x), x.size, _p(y))
        return y


class AmpModem:
    def __init__(self, mod_index=0.8, am_type=0, suppressed=0):
        self.L = lib()
        self.h = self.L.orc_ampmodem_create(mod_index, am_type, suppressed)

    def close(self):
        if self.h:
            self.L.orc_ampmodem_destroy(self.h)
            self.h = None

    __del__ = close

    def execute(self, x):
        x = _cf(x)
        y = np.empty(x.size, np.float32)
        self.L.orc_ampmodem_demodulate_block(self.h, _p(x), x.size, _p(y))
        return y


DEMOD_NO, DEMOD_NBFM, DEMOD_AM, DEMOD_WBFM = 0, 1, 2, 3


class Firpfbch2:
    """liquid firpfbch2_crcf analyzer (2x oversampled; M/2 in, M out).  Not called by the reference."""

    def __init__(self, nch, m=7, As=80.0):
        self.L = lib()
        self.M = int(nch)
        self.h = self.L.orc_firpfbch2_crcf_create_kaiser(0, self.M, m, As)
        if not self.h:
            raise ValueError("firpfbch2_crcf_create_kaiser: M must be even")

    def close(self):
        if self.h:
            self.L.orc_firpfbch2_crcf_destroy(self.h)
            self.h = None

    __del__ = close

    def taps(self):
        n = C.c_uint(0)
        p = self.L.orc_firpfbch2_taps(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def execute(self, x):
        """whole frames of M/2 samples -> [M][2 len(x) / M] channel-major"""
        x = _cf(x)
        M2 = self.M // 2
        nf = x.size // M2
        y = np.empty((nf, self.M), np.complex64)
        for t in range(nf):
            self.L.orc_firpfbch2_crcf_execute(self.h, x[t * M2:].ctypes.data, y[t].ctypes.data)
        return np.ascontiguousarray(y.T)


class IirFiltRRRF:
    """liquid iirfilt_rrrf_create_prototype, Butterworth low-pass in second-order sections (reference: iirFilter,
    Liquid.chs:629-650)."""

    def __init__(self, order, fc, f0=0.0, ap=10.0, as_=10.0):
        self.L = lib()
        self.h = self.L.orc_iirfilt_rrrf_create_prototype(0, 0, 0, order, fc, f0, ap, as_)
        if not self.h:
            raise ValueError("iirfilt_rrrf_create_prototype: unsupported prototype")

    def close(self):
        if self.h:
            self.L.orc_iirfilt_rrrf_destroy(self.h)
            self.h = None

    __del__ = close

    def coeffs(self):
        b = np.zeros(24, np.float32)
        a = np.zeros(24, np.float32)
        n = self.L.orc_iirfilt_rrrf_coeffs(self.h, b.ctypes.data, a.ctypes.data)
        return b[:3 * n].reshape(n, 3), a[:3 * n].reshape(n, 3)

    def execute(self, x):
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty_like(x)
        self.L.orc_iirfilt_rrrf_execute_block(self.h, x.ctypes.data, x.size, y.ctypes.data)
        return y


class FirDecim:
    """liquid firdecim_rrrf_create_kaiser (reference: firDecimator, Liquid.chs:487-503)."""

    def __init__(self, M, m=10, As=60.0):
        self.L = lib()
        self.M = int(M)
        self.h = self.L.orc_firdecim_rrrf_create_kaiser(self.M, m, As)

    def close(self):
        if self.h:
            self.L.orc_firdecim_rrrf_destroy(self.h)
            self.h = None

    __del__ = close

    def taps(self):
        n = C.c_uint(0)
        p = self.L.orc_firdecim_rrrf_taps(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def execute(self, x):
        """whole blocks of M samples only (len(x) // M outputs), like firDecim (Liquid.chs:497-500)"""
        x = np.ascontiguousarray(x, np.float32)
        n = x.size // self.M
        y = np.empty(n, np.float32)
        self.L.orc_firdecim_rrrf_execute_block(self.h, x.ctypes.data, n, y.ctypes.data)
        return y


class Chain:
    """sdrProcess (apps/SoapySDR.hs:181-283) as one sequential object."""

    def __init__(self, samplerate, offset_hz=0.0, bandwidth_hz=0.0, demod=DEMOD_NO, kf=0.3, agc_thresh_db=0.0,
                 channels=1, mix=False, fast=False, decim=1, channelizer=0):
        self.L = lib(fast)
        self.cfg = ChainCfg(samplerate, offset_hz, bandwidth_hz, demod, kf, agc_thresh_db, channels, int(mix), int(decim), int(channelizer))
        self.h = self.L.orc_chain_create(C.byref(self.cfg))
        self.nout = self.L.orc_chain_num_outputs(self.h)
        self.dtype = np.float32 if demod else np.complex64
        r = (bandwidth_hz / samplerate) if bandwidth_hz else 1.0
        self._ratio = r / max(1, channels) * (2 if channelizer else 1)

    def close(self):
        if self.h:
            self.L.orc_chain_destroy(self.h)
            self.h = None

    __del__ = close

    def process(self, x):
        x = _cf(x)
        cap = int(2 * np.ceil(self._ratio * x.size)) + 4096
        outs = [np.empty(cap, self.dtype) for _ in range(self.nout)]
        arr = (C.c_void_p * self.nout)(*[o.ctypes.data for o in outs])
        n = C.c_size_t(0)
        rc = self.L.orc_chain_process(self.h, _p(x), x.size, arr, cap, C.byref(n))
        if rc != 0:
            raise RuntimeError("orc_chain_process: output capacity too small")
        return [o[:n.value].copy() for o in outs]
