"""Side-by-side run of a REAL liquid-dsp (libliquid.so) against the oracle -- the only way the parity claims of this
repo ever get pinned to the library the reference links (composable-sdr.cabal:38, liquid-dsp 1.3.2).

TEST INFRASTRUCTURE ONLY (tests/test_real_liquid.py and bench.py's cpu_baseline leg).  `find()` looks for the library
(env LIQUID_SO, then ctypes.util.find_library("liquid")); when there is none -- as in the build container and on the
stock GPU image -- everything that needs it skips.  `LiquidChain` drives the library with exactly the call sequences of
src/ComposableSDR/Liquid.chs and the block order of apps/SoapySDR.hs:181-283 (sdrProcess); the symbol prefix is a
parameter so that the same driver can be pointed at oracle/liboracle.so ("orc_" + the liquid name, same signatures),
which is how the driver itself is tested where no libliquid exists.
"""
import ctypes as C
import ctypes.util
import os

import numpy as np

DEMOD_NO, DEMOD_NBFM, DEMOD_AM = 0, 1, 2


def find():
    """path of a real libliquid, or None"""
    p = os.environ.get("LIQUID_SO")
    if p and os.path.exists(p):
        return p
    p = ctypes.util.find_library("liquid")
    if p:
        return p
    for d in ("/usr/lib", "/usr/local/lib", "/usr/lib/x86_64-linux-gnu", "/usr/lib64"):
        for name in ("libliquid.so", "libliquid.so.1", "libliquid.so.1.3"):
            if os.path.exists(os.path.join(d, name)):
                return os.path.join(d, name)
    return None


_SIG = {
    # Liquid.chs:746-780
    "nco_crcf_create": (C.c_void_p, [C.c_int]), "nco_crcf_destroy": (None, [C.c_void_p]),
    "nco_crcf_set_frequency": (None, [C.c_void_p, C.c_float]),
    "nco_crcf_mix_block_down": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]),
    "nco_crcf_mix_block_up": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]),
    # Liquid.chs:58-73
    "msresamp_crcf_create": (C.c_void_p, [C.c_float, C.c_float]), "msresamp_crcf_destroy": (None, [C.c_void_p]),
    "msresamp_crcf_get_rate": (C.c_float, [C.c_void_p]),
    "msresamp_crcf_execute": (None, [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_uint)]),
    # Liquid.chs:550-567
    "iirfilt_crcf_create_dc_blocker": (C.c_void_p, [C.c_float]), "iirfilt_crcf_destroy": (None, [C.c_void_p]),
    "iirfilt_crcf_execute_block": (None, [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    # Liquid.chs:732-742
    "firpfbch_crcf_create_kaiser": (C.c_void_p, [C.c_int, C.c_uint, C.c_uint, C.c_float]),
    "firpfbch_crcf_destroy": (None, [C.c_void_p]),
    "firpfbch_crcf_analyzer_execute": (None, [C.c_void_p, C.c_void_p, C.c_void_p]),
    # Liquid.chs:660-691
    "agc_crcf_create": (C.c_void_p, []), "agc_crcf_destroy": (None, [C.c_void_p]),
    "agc_crcf_set_bandwidth": (None, [C.c_void_p, C.c_float]), "agc_crcf_set_signal_level": (None, [C.c_void_p, C.c_float]),
    "agc_crcf_squelch_enable": (None, [C.c_void_p]), "agc_crcf_squelch_set_threshold": (None, [C.c_void_p, C.c_float]),
    "agc_crcf_squelch_set_timeout": (None, [C.c_void_p, C.c_uint]),
    "agc_crcf_execute_block": (None, [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "agc_crcf_squelch_get_status": (C.c_int, [C.c_void_p]), "agc_crcf_get_rssi": (C.c_float, [C.c_void_p]),
    # Liquid.chs:305-315, 441-450
    "freqdem_create": (C.c_void_p, [C.c_float]), "freqdem_destroy": (None, [C.c_void_p]),
    "freqdem_demodulate_block": (None, [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
    "ampmodem_create": (C.c_void_p, [C.c_float, C.c_int, C.c_int]), "ampmodem_destroy": (None, [C.c_void_p]),
    "ampmodem_demodulate_block": (None, [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]),
}


class Liquid:
    """the hot-path imports of Liquid.chs bound from `path` under `prefix` + liquid's names"""

    def __init__(self, path, prefix=""):
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        for name, (res, args) in _SIG.items():
            fn = getattr(self.dll, prefix + name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)


def _p(a):
    return a.ctypes.data


class LiquidChain:
    """sdrProcess (apps/SoapySDR.hs:181-283) on liquid objects, block by block as the Haskell wrappers call them."""

    def __init__(self, liq, samplerate, offset_hz=0.0, bandwidth_hz=0.0, demod=DEMOD_NO, kf=0.3, agc_thresh_db=0.0,
                 channels=1, mix=False, chunk=1 << 16):
        self.q, self.chunk = liq, int(chunk)
        self.C = max(1, int(channels))
        self.mix, self.demod, self.agc_thr = bool(mix), demod, float(agc_thresh_db)
        q = liq
        # offset: f = 2 pi offset / samplerate :: Float; f > 0 mixDown f, f < 0 mixUp (-f)    (SoapySDR.hs:200-205)
        f = float(np.float32(2) * np.float32(np.pi) * np.float32(offset_hz) / np.float32(samplerate))
        self.nco, self.up = None, f < 0
        if f != 0.0:
            self.nco = q.nco_crcf_create(1)                                   # ncoCreate: LIQUID_VCO (Liquid.chs:784)
            q.nco_crcf_set_frequency(self.nco, abs(f))
        self.rs = q.msresamp_crcf_create(float(np.float32(bandwidth_hz / samplerate)), 60.0) if bandwidth_hz else None
        self.dc = q.iirfilt_crcf_create_dc_blocker(0.0005)                    # Liquid.chs:577
        self.fb = self.fb_nco = None
        if self.C > 1:
            self.fb = q.firpfbch_crcf_create_kaiser(0, self.C, 7, 80.0)       # Liquid.chs:813
            self.fb_nco = q.nco_crcf_create(1)
            n = np.float32(self.C)
            off = -(np.float32(0.5) * (n - np.float32(1)) / n * np.float32(2) * np.float32(np.pi))   # Liquid.chs:817-818
            q.nco_crcf_set_frequency(self.fb_nco, float(off))
        self.agc, self.dem = [], []
        for _ in range(self.C):
            if self.agc_thr != 0.0:                                           # automaticGainControl (Liquid.chs:707-717)
                a = q.agc_crcf_create()
                q.agc_crcf_set_bandwidth(a, 0.1)
                q.agc_crcf_set_signal_level(a, 1e-3)
                q.agc_crcf_squelch_enable(a)
                q.agc_crcf_squelch_set_threshold(a, self.agc_thr)
                q.agc_crcf_squelch_set_timeout(a, 1000)
                self.agc.append(a)
            if demod == DEMOD_NBFM:
                self.dem.append(q.freqdem_create(kf))
            elif demod == DEMOD_AM:
                self.dem.append(q.ampmodem_create(0.8, 0, 0))                 # Liquid.chs:452-453
        self.left = np.empty(0, np.complex64)

    def close(self):
        q = self.q
        for h in self.agc:
            q.agc_crcf_destroy(h)
        for h in self.dem:
            (q.freqdem_destroy if self.demod == DEMOD_NBFM else q.ampmodem_destroy)(h)
        for h, d in ((self.nco, q.nco_crcf_destroy), (self.fb_nco, q.nco_crcf_destroy), (self.rs, q.msresamp_crcf_destroy),
                     (self.dc, q.iirfilt_crcf_destroy), (self.fb, q.firpfbch_crcf_destroy)):
            if h:
                d(h)
        self.agc, self.dem, self.nco, self.fb_nco, self.rs, self.dc, self.fb = [], [], None, None, None, None, None

    def _agc(self, h, x):
        """agcExecuteBlock (Liquid.chs:693-705): per sample execute_block(.., 1, ..), squelch_get_status, get_rssi"""
        q = self.q
        y = np.zeros_like(x)
        one = np.zeros(1, np.complex64)
        xp, step = _p(x), x.itemsize
        for i in range(x.size):
            q.agc_crcf_execute_block(h, xp + i * step, 1, _p(one))
            if q.agc_crcf_squelch_get_status(h) == 3:
                y[i] = one[0]
        return y

    def _demod(self, k, x):
        q = self.q
        if self.agc:
            x = self._agc(self.agc[k], x)
        if not self.dem:
            return x
        m = np.empty(x.size, np.float32)
        (q.freqdem_demodulate_block if self.demod == DEMOD_NBFM else q.ampmodem_demodulate_block)(self.dem[k], _p(x), x.size, _p(m))
        return m

    def process(self, x):
        """whole input -> list of outputs (one, or C without --mix).  The < C samples left at the end are dropped, as
        the reference's last `compact` flush does (Liquid.chs:835)."""
        q = self.q
        x = np.ascontiguousarray(x, np.complex64)
        pieces = []
        for i in range(0, x.size, self.chunk):
            a = np.ascontiguousarray(x[i:i + self.chunk])
            if self.nco:
                m = np.empty_like(a)
                (q.nco_crcf_mix_block_up if self.up else q.nco_crcf_mix_block_down)(self.nco, _p(a), _p(m), a.size)
                a = m
            if self.rs:
                rate = q.msresamp_crcf_get_rate(self.rs)
                y = np.empty(2 * int(np.ceil(rate * a.size)) + 8, np.complex64)
                ny = C.c_uint(0)
                q.msresamp_crcf_execute(self.rs, _p(a), a.size, _p(y), C.byref(ny))
                a = y[:ny.value].copy()
            z = np.empty_like(a)
            if a.size:
                q.iirfilt_crcf_execute_block(self.dc, _p(a), a.size, _p(z))
            pieces.append(z)
        r = np.concatenate(pieces) if pieces else np.empty(0, np.complex64)
        if self.C == 1:
            return [self._demod(0, r)]
        Cn = self.C
        r = np.concatenate([self.left, r])
        nf = r.size // Cn
        self.left = r[nf * Cn:].copy()
        r = np.ascontiguousarray(r[:nf * Cn])
        rot = np.empty_like(r)
        if r.size:
            q.nco_crcf_mix_block_down(self.fb_nco, _p(r), _p(rot), r.size)     # Liquid.chs:847
        chan = np.empty((Cn, nf), np.complex64)
        frame = np.empty(Cn, np.complex64)
        for t in range(nf):                                                    # Liquid.chs:837-849
            q.firpfbch_crcf_analyzer_execute(self.fb, _p(rot) + t * Cn * 8, _p(frame))
            chan[:, t] = frame
        outs = [self._demod(k, np.ascontiguousarray(chan[k])) for k in range(Cn)]
        if self.mix:
            acc = outs[0]
            for o in outs[1:]:
                acc = acc + o                                                  # foldl1 (zipWith (+)), Trans.hs:119-122
            return [acc]
        return outs
