"""Host-side mirror of the reference's block API, on top of the C ABI (include/csdr_b200.h).

The reference's operator interface is the Haskell record `Pipe {_start, _process, _done}` (src/ComposableSDR/
Types.hs:51-55) with `compose` (Types.hs:93-99), `unPipe` / `addPipe` (Types.hs:109-131) and the stream
combinators of src/ComposableSDR/Trans.hs.  No GHC exists in this image, so the mirror is Python: same block
names, same arguments, same chunk protocol (start once, process per chunk in stream order, done once).  Arrays are
numpy (host) arrays or torch CUDA tensors; the C ABI accepts both kinds of pointers.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CsdrError, ChainCfg  # noqa: F401

__all__ = ["Pipe", "Fold", "compose", "unPipe", "addPipe", "takeNArr", "compact", "mux", "mix", "distribute_",
           "mixDown", "mixUp", "resampler", "dcBlocker", "firpfbchChannelizer", "firpfbch2Channelizer", "automaticGainControl",
           "fmDemodulator", "amDemodulator", "iirFilter", "firDecimator", "wbFMDemodulator", "listSink", "DeNo", "DeNBFM",
           "DeAM", "DeWBFM", "Chain", "sdrProcess",
           "CsdrError", "kernel_launches", "set_option", "device_count", "PinnedBuffer"]


# --------------------------------------------------------------------------------------------- array helpers
def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


def _as_cf32(x):
    if _is_torch(x):
        import torch
        if x.dtype != torch.complex64 or not x.is_contiguous():
            x = x.to(torch.complex64).contiguous()
        return x
    return np.ascontiguousarray(x, dtype=np.complex64)


def _pre_sync(x):
    """torch tensors are produced on torch's stream; the library runs on its own streams."""
    if _is_torch(x) and x.is_cuda:
        import torch
        torch.cuda.current_stream(x.device).synchronize()


def _post_sync(x):
    if _is_torch(x) and x.is_cuda:
        _lib.load().csdr_synchronize()


def _empty_like_kind(x, n, dtype):
    if _is_torch(x):
        import torch
        return torch.empty(n, dtype=torch.float32 if dtype == np.float32 else torch.complex64, device=x.device)
    return np.empty(n, dtype=dtype)


def kernel_launches():
    return int(_lib.load().csdr_kernel_launches())


def set_option(opt, value):
    _lib.load().csdr_set_option(int(opt), int(value))


def device_count():
    return int(_lib.load().csdr_device_count())


class PinnedBuffer:
    """Pinned host memory from csdr_host_alloc, viewed as a numpy array."""

    def __init__(self, n, dtype):
        self.L = _lib.load()
        self.nbytes = int(n) * np.dtype(dtype).itemsize
        self.p = self.L.csdr_host_alloc(self.nbytes)
        if not self.p:
            raise CsdrError("csdr_host_alloc failed: " + _lib.last_error())
        buf = (C.c_char * self.nbytes).from_address(self.p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))

    def close(self):
        if self.p:
            self.array = None
            self.L.csdr_host_free(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------- Pipe / Fold
class Pipe:
    """`data Pipe m a b = Pipe {_start :: m r, _process :: r -> a -> m b, _done :: r -> m ()}` (Types.hs:51-55)."""

    def __init__(self, start, process, done):
        self._start, self._process, self._done = start, process, done

    def __mul__(self, other):          # Category (.): (p1 * p2) runs p2 first, like `p1 . p2`
        return compose(self, other)


def compose(p1, p2):
    """Types.hs:93-99: acquire (r1, r2), process2 then process1, release done2 then done1."""
    def start():
        return (p1._start(), p2._start())

    def process(r, a):
        return p1._process(r[0], p2._process(r[1], a))

    def done(r):
        p2._done(r[1])
        p1._done(r[0])
    return Pipe(start, process, done)


identity = Pipe(lambda: None, lambda r, a: a, lambda r: None)


def unPipe(pipe):
    """Types.hs:109-115: returns (stream transformer, cleanup action)."""
    r = pipe._start()

    def process(stream):
        for a in stream:
            yield pipe._process(r, a)
    return process, (lambda: pipe._done(r))


class Fold:
    """Streamly `Fold step initial extract`, as used by addPipe / compact / distribute_."""

    def __init__(self, step, start, done):
        self.step, self.start, self.done = step, start, done

    def run(self, stream):
        s = self.start()
        for a in stream:
            s = self.step(s, a)
        return self.done(s)


def addPipe(pipe, fold):
    """Types.hs:117-131."""
    def start():
        s = fold.start()
        r = pipe._start()
        return [s, r]

    def step(st, a):
        st[0] = fold.step(st[0], pipe._process(st[1], a))
        return st

    def done(st):
        pipe._done(st[1])
        return fold.done(st[0])
    return Fold(step, start, done)


def takeNArr(n, stream):
    """Trans.hs:33-56: pass chunks until n samples were seen, trimming the chunk that crosses n."""
    seen = 0
    for a in stream:
        togo = n - seen
        if togo == 0:
            return
        if togo >= len(a):
            seen += len(a)
            yield a
        else:
            seen = n
            yield a[:togo]


def _concat(a, b):
    if _is_torch(a) or _is_torch(b):
        import torch
        return torch.cat([a, b])
    return np.concatenate([a, b])


def compact(n, fold):
    """Trans.hs:58-85: re-chunk to exactly n samples; the remainder is flushed at the end."""
    def start():
        return [fold.start(), None]

    def step(st, a):
        b = st[1]
        ba = a if b is None or len(b) == 0 else _concat(b, a)
        if len(ba) >= n:
            st[0] = fold.step(st[0], ba[:n])
            st[1] = ba[n:]
        else:
            st[1] = ba
        return st

    def done(st):
        b = st[1]
        if b is None:
            b = np.empty(0, np.complex64)
        return fold.done(fold.step(st[0], b))
    return Fold(step, start, done)


def mux(pipes):
    """Trans.hs:124-129: one resource per list element, zipWithM process."""
    pipes = list(pipes)
    return Pipe(lambda: [p._start() for p in pipes],
                lambda rs, xs: [p._process(r, x) for p, r, x in zip(pipes, rs, xs)],
                lambda rs: [p._done(r) for p, r in zip(pipes, rs)])


def _mix_sum(arrays):
    acc = arrays[0]
    for a in arrays[1:]:
        acc = acc + a                      # foldl1 (zipWith (+)), channel order (Trans.hs:119-122)
    return acc


mix = Pipe(lambda: None, lambda r, arrays: _mix_sum(arrays), lambda r: None)


def distribute_(folds):
    """Trans.hs:101-117."""
    folds = list(folds)
    return Fold(lambda ss, xs: [f.step(s, x) for f, s, x in zip(folds, ss, xs)],
                lambda: [f.start() for f in folds],
                lambda ss: [f.done(s) for f, s in zip(folds, ss)])


def listSink():
    """Collects chunks (stands in for fileSink / audioFileSink, Sink.hs)."""
    def done(chunks):
        chunks = [np.asarray(c.cpu() if _is_torch(c) else c) for c in chunks if len(c)]
        return np.concatenate(chunks) if chunks else np.empty(0)
    return Fold(lambda s, a: s + [a], lambda: [], done)


# --------------------------------------------------------------------------------------------- liquid blocks
class _Handle:
    def __init__(self, create, destroy, what):
        self.L = _lib.load()
        self.h = _lib.check_handle(create(self.L), what)
        self._destroy = destroy

    def close(self):
        if self.h:
            self._destroy(self.L, self.h)
            self.h = None


def _block(create, destroy, process, what):
    def start():
        return _Handle(create, destroy, what)

    def proc(r, a):
        _pre_sync(a)
        out = process(r.L, r.h, a)
        _lib.check_call(what)
        _post_sync(a)
        return out
    return Pipe(start, proc, lambda r: r.close())


def _nco_create(f):
    def create(L):
        h = L.csdr_nco_crcf_create(1)              # 1 == VCO (Liquid.chs:784)
        if h:
            L.csdr_nco_crcf_set_frequency(h, f)
        return h
    return create


def _same_size(fn, out_dtype=np.complex64):
    def process(L, h, a):
        a = _as_cf32(a)
        y = _empty_like_kind(a, len(a), out_dtype)
        fn(L)(h, _ptr(a), len(a), _ptr(y))
        return y
    return process


def mixDown(f):
    """Liquid.chs:799-800."""
    def process(L, h, a):
        a = _as_cf32(a)
        y = _empty_like_kind(a, len(a), np.complex64)
        L.csdr_nco_crcf_mix_block_down(h, _ptr(a), _ptr(y), len(a))
        return y
    return _block(_nco_create(f), lambda L, h: L.csdr_nco_crcf_destroy(h), process, "mixDown")


def mixUp(f):
    """Liquid.chs:808-809."""
    def process(L, h, a):
        a = _as_cf32(a)
        y = _empty_like_kind(a, len(a), np.complex64)
        L.csdr_nco_crcf_mix_block_up(h, _ptr(a), _ptr(y), len(a))
        return y
    return _block(_nco_create(f), lambda L, h: L.csdr_nco_crcf_destroy(h), process, "mixUp")


def resampler(r, as_):
    """Liquid.chs:115-117; output buffer 2*ceil(r*nx) like Liquid.chs:81-82."""
    def process(L, h, a):
        a = _as_cf32(a)
        rate = L.csdr_msresamp_crcf_get_rate(h)
        cap = 2 * int(np.ceil(rate * len(a)))
        y = _empty_like_kind(a, max(cap, 1), np.complex64)
        ny = C.c_uint(0)
        L.csdr_msresamp_crcf_execute(h, _ptr(a), len(a), _ptr(y), C.byref(ny))
        return y[:ny.value]
    return _block(lambda L: L.csdr_msresamp_crcf_create(r, as_), lambda L, h: L.csdr_msresamp_crcf_destroy(h),
                  process, "resampler")


def dcBlocker():
    """Liquid.chs:591-592 (alpha = 0.0005, Liquid.chs:577)."""
    return _block(lambda L: L.csdr_iirfilt_crcf_create_dc_blocker(0.0005), lambda L, h: L.csdr_iirfilt_crcf_destroy(h),
                  _same_size(lambda L: L.csdr_iirfilt_crcf_execute_block), "dcBlocker")


def firpfbchChannelizer(n):
    """Liquid.chs:864-866: kaiser(m=7, As=80) analyzer + pre-rotation NCO; returns a list of n channel arrays."""
    class _Fb:
        def __init__(self):
            self.L = _lib.load()
            self.fb = _lib.check_handle(self.L.csdr_firpfbch_crcf_create_kaiser(0, n, 7, 80.0), "firpfbch_crcf_create_kaiser")
            self.nco = _lib.check_handle(self.L.csdr_nco_crcf_create(1), "nco_crcf_create")
            off = np.float32(0.5) * (np.float32(n) - np.float32(1)) / np.float32(n) * np.float32(2) * np.float32(np.pi)
            self.L.csdr_nco_crcf_set_frequency(self.nco, float(-off))        # Liquid.chs:817-818

        def close(self):
            if self.fb:
                self.L.csdr_firpfbch_crcf_destroy(self.fb)
                self.L.csdr_nco_crcf_destroy(self.nco)
                self.fb = None

    def process(r, a):
        a = _as_cf32(a)
        nf = len(a) // n
        y = _empty_like_kind(a, max(nf * n, 1), np.complex64)
        _pre_sync(a)
        rc = r.L.csdr_firpfbch_execute_block(r.fb, r.nco, _ptr(a), len(a), _ptr(y))
        if rc != 0:
            raise CsdrError("firpfbchChannelizer: " + _lib.last_error())
        _post_sync(a)
        return [y[nf * j: nf * (j + 1)] for j in range(n)]
    return Pipe(_Fb, process, lambda r: r.close())


def firpfbch2Channelizer(n, m=7, as_=80.0):
    """The 2x oversampled analyzer (liquid firpfbch2_crcf; not in the reference): an array of k n/2
    samples gives a list of n channel arrays of k samples each (channel rate = 2 / n of the input rate)."""
    def process(L, h, a):
        a = _as_cf32(a)
        nf = len(a) // (n // 2)
        y = _empty_like_kind(a, max(nf * n, 1), np.complex64)
        if nf and L.csdr_firpfbch2_execute_block(h, _ptr(a), len(a), _ptr(y)) != 0:
            raise CsdrError("firpfbch2Channelizer: " + _lib.last_error())
        return [y[nf * j: nf * (j + 1)] for j in range(n)]
    return _block(lambda L: L.csdr_firpfbch2_crcf_create_kaiser(0, n, m, as_), lambda L, h: L.csdr_firpfbch2_crcf_destroy(h),
                  process, "firpfbch2Channelizer")


def automaticGainControl(tres):
    """Liquid.chs:707-728 (bw 0.1, level 1e-3, squelch on, timeout 1000) with the gate of Liquid.chs:693-705."""
    def create(L):
        h = L.csdr_agc_crcf_create()
        if h:
            L.csdr_agc_crcf_set_bandwidth(h, 0.1)
            L.csdr_agc_crcf_set_signal_level(h, 1e-3)
            L.csdr_agc_crcf_squelch_enable(h)
            L.csdr_agc_crcf_squelch_set_threshold(h, tres)
            L.csdr_agc_crcf_squelch_set_timeout(h, 1000)
        return h

    def process(L, h, a):
        a = _as_cf32(a)
        y = _empty_like_kind(a, len(a), np.complex64)
        if L.csdr_agc_squelch_execute_block(h, _ptr(a), len(a), _ptr(y)) != 0:
            raise CsdrError("automaticGainControl: " + _lib.last_error())
        return y
    return _block(create, lambda L, h: L.csdr_agc_crcf_destroy(h), process, "automaticGainControl")


def fmDemodulator(kf):
    """Liquid.chs:333-334."""
    return _block(lambda L: L.csdr_freqdem_create(kf), lambda L, h: L.csdr_freqdem_destroy(h),
                  _same_size(lambda L: L.csdr_freqdem_demodulate_block, np.float32), "fmDemodulator")


def amDemodulator():
    """Liquid.chs:468-469: ampmodem_create 0.8 0 0."""
    return _block(lambda L: L.csdr_ampmodem_create(0.8, 0, 0), lambda L, h: L.csdr_ampmodem_destroy(h),
                  _same_size(lambda L: L.csdr_ampmodem_demodulate_block, np.float32), "amDemodulator")


def _as_f32(a):
    if _is_torch(a):
        import torch
        return a.to(torch.float32).contiguous()
    return np.ascontiguousarray(a, dtype=np.float32)


def iirFilter(n, fc, f0, ap, as_):
    """Liquid.chs:644-650: iirfilt_rrrf_create_prototype 0 0 0 n fc f0 ap as (Butterworth low-pass, sections)."""
    def process(L, h, a):
        a = _as_f32(a)
        y = _empty_like_kind(a, len(a), np.float32)
        L.csdr_iirfilt_rrrf_execute_block(h, _ptr(a), len(a), _ptr(y))
        return y
    return _block(lambda L: L.csdr_iirfilt_rrrf_create_prototype(0, 0, 0, n, fc, f0, ap, as_),
                  lambda L, h: L.csdr_iirfilt_rrrf_destroy(h), process, "iirFilter")


def firDecimator(m):
    """Liquid.chs:487-503: firdecim_rrrf_create_kaiser m 10 60; an array of length n gives n div m samples, the
    remainder of the array is dropped like in the reference."""
    def process(L, h, a):
        a = _as_f32(a)
        n = len(a) // m
        y = _empty_like_kind(a, n, np.float32)
        if n:
            L.csdr_firdecim_rrrf_execute_block(h, _ptr(a), n, _ptr(y))
        return y
    return _block(lambda L: L.csdr_firdecim_rrrf_create_kaiser(m, 10, 60.0), lambda L, h: L.csdr_firdecim_rrrf_destroy(h),
                  process, "firDecimator")


def wbFMDemodulator(quad_rate, decim):
    """Liquid.chs:652-656."""
    return firDecimator(decim) * iirFilter(2, float(np.float32(5000.0 / quad_rate)), 0.0, 10.0, 10.0) * fmDemodulator(0.6)


# --------------------------------------------------------------------------------------------- the app graph
class DeNo:
    code = 0
    kf = 0.0


class DeNBFM:
    code = 1

    def __init__(self, kf):
        self.kf = float(kf)


class DeAM:
    code = 2
    kf = 0.0


class DeWBFM:
    """DeWBFM decim (apps/SoapySDR.hs:253-260): wbFMDemodulator outBW decim"""
    code = 3
    kf = 0.6

    def __init__(self, decim):
        self.decim = int(decim)


def sdrProcess(src, samplerate, offset=0.0, bandwidth=0.0, numsamples=None, demod=None, agc=0.0, channels=1,
               mix_channels=False):
    """apps/SoapySDR.hs:181-283 assembled from the individual blocks (un-fused; every block is its own C-ABI
    handle, chunks flow exactly like the reference).  `src` yields CF32 chunks.  Returns the sink contents:
    one array, or a list of `channels` arrays."""
    demod = demod or DeNo()
    rs = identity if bandwidth == 0 else resampler(float(np.float32(bandwidth / samplerate)), 60.0)
    f = float(np.float32(2) * np.float32(np.pi) * np.float32(offset) / np.float32(samplerate))
    off = identity if f == 0 else (mixDown(f) if f > 0 else mixUp(-f))
    agc_p = automaticGainControl(agc) if agc != 0.0 else identity
    process, cleanup = unPipe(rs * off)
    stream = process(src)
    if numsamples is not None:
        stream = takeNArr(numsamples, stream)
    if isinstance(demod, DeNBFM):
        dem = fmDemodulator(demod.kf) * agc_p
    elif isinstance(demod, DeAM):
        dem = amDemodulator() * agc_p
    elif isinstance(demod, DeWBFM):
        dem = wbFMDemodulator(bandwidth if bandwidth else samplerate, demod.decim) * agc_p
    else:
        dem = agc_p
    nch, m = channels, 4
    if nch > 1:
        if mix_channels:
            inner = addPipe(mix * mux([dem] * nch) * firpfbchChannelizer(nch), listSink())
        else:
            inner = addPipe(firpfbchChannelizer(nch), distribute_([addPipe(dem, listSink()) for _ in range(nch)]))
    else:
        inner = addPipe(dem, listSink())
    fold = addPipe(dcBlocker(), compact(m * nch * 1024, inner))
    try:
        return fold.run(stream)
    finally:
        cleanup()


# --------------------------------------------------------------------------------------------- fused chain
class Chain:
    """csdr_chain_*: the whole of sdrProcess behind one handle (device-resident state, fused kernels)."""

    def __init__(self, samplerate, offset=0.0, bandwidth=0.0, demod=None, agc=0.0, channels=1, mix_channels=False,
                 nstreams=1, device=-1, channelizer=0):
        demod = demod or DeNo()
        self.L = _lib.load()
        self.cfg = ChainCfg(float(samplerate), float(offset), float(bandwidth), demod.code, float(demod.kf), float(agc),
                            int(channels), int(bool(mix_channels)), int(nstreams), int(device),
                            int(getattr(demod, "decim", 0)), int(channelizer))
        self.h = _lib.check_handle(self.L.csdr_chain_create(C.byref(self.cfg)), "csdr_chain_create")
        self.nout = int(self.L.csdr_chain_num_outputs(self.h))
        self.nstreams = max(1, int(nstreams))
        self.out_dtype = np.float32 if demod.code else np.complex64

    def close(self):
        if self.h:
            self.L.csdr_chain_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def max_output(self, nx):
        return int(self.L.csdr_chain_max_output(self.h, int(nx)))

    @property
    def cuda_stream(self):
        return int(self.L.csdr_chain_cuda_stream(self.h) or 0)

    def run_file(self, in_path, out_name, numsamples=0, chunk=0):
        """soapy-sdr --filename in_path -n numsamples --output out_name: CF32 file in, <out_name>.cf32 / _chK.cf32 /
        .f32 out (apps/SoapySDR.hs:209-240).  Returns (input samples consumed, samples written per output file)."""
        n_in, n_out = C.c_uint64(0), C.c_uint64(0)
        rc = self.L.csdr_chain_run_file(self.h, str(in_path).encode(), str(out_name).encode(), int(numsamples), int(chunk),
                                        C.byref(n_in), C.byref(n_out))
        if rc != 0:
            raise CsdrError("csdr_chain_run_file: " + _lib.last_error())
        return n_in.value, n_out.value

    def seek(self, n_prior):
        if self.L.csdr_chain_seek(self.h, int(n_prior)) != 0:
            raise CsdrError("csdr_chain_seek: " + _lib.last_error())

    def warmup_len(self):
        return int(self.L.csdr_chain_warmup_len(self.h))

    def agc_fixups(self):
        return int(self.L.csdr_chain_agc_fixups(self.h))

    def agc_counters(self):
        """cumulative (gain segments repaired in order, squelch-FSM segments repaired in order, gain segments refined)"""
        v = (C.c_uint64 * 3)()
        self.L.csdr_chain_agc_counters(self.h, v)
        return int(v[0]), int(v[1]), int(v[2])

    def agc_plan(self):
        """(segment length, warm-up length) of the gain-loop speculation in the last call"""
        v = (C.c_int * 2)()
        self.L.csdr_chain_agc_plan(self.h, v)
        return int(v[0]), int(v[1])

    def print(self):
        self.L.csdr_chain_print(self.h)

    def profile(self, enable=True):
        self.L.csdr_chain_profile(self.h, int(enable))

    def frontend_ms(self):
        """(accumulated k_frontend device time in ms, launches) since profile(True)"""
        n = C.c_uint64(0)
        ms = self.L.csdr_chain_frontend_ms(self.h, C.byref(n))
        return float(ms), int(n.value)

    def process_raw(self, x_ptr, nx, x_stride, out_ptrs, out_cap):
        """Pointer-level call (host or device pointers).  Returns samples written per output."""
        arr = (C.c_void_p * len(out_ptrs))(*out_ptrs)
        n = C.c_size_t(0)
        rc = self.L.csdr_chain_process(self.h, x_ptr, int(nx), int(x_stride), arr, int(out_cap), C.byref(n))
        if rc != 0:
            raise CsdrError("csdr_chain_process: " + _lib.last_error())
        return int(n.value)

    def process(self, x):
        """x: [nx] (or [nstreams, nx]) complex64, numpy or torch-cuda.  Returns a list of nstreams*nout arrays."""
        x = _as_cf32(x)
        nx = x.shape[-1]
        stride = nx
        cap = self.max_output(nx)
        nptr = self.nstreams * self.nout
        outs = [_empty_like_kind(x, max(cap, 1), self.out_dtype) for _ in range(nptr)]
        _pre_sync(x)
        n = self.process_raw(_ptr(x), nx, stride, [_ptr(o) for o in outs], cap)
        _post_sync(x)
        return [o[:n] for o in outs]
