"""ctypes wrapper of the TEST-ONLY kernel emulation library (tests/emu/emu_kernels.cpp)."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def load():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        L.emu_frontend.restype = C.c_longlong
        L.emu_frontend.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_ulonglong, C.c_int, C.c_longlong]
        L.emu_backend.restype = C.c_longlong
        L.emu_backend.argtypes = [C.c_int, C.c_longlong, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float,
                                  C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p,
                                  C.POINTER(C.c_ulonglong), C.c_void_p]
        L.emu_pfb.restype = C.c_longlong
        L.emu_pfb.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p]
        L.emu_pfb2.restype = C.c_longlong
        L.emu_pfb2.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.emu_wbfm_tail.restype = C.c_longlong
        L.emu_wbfm_tail.argtypes = [C.c_int, C.c_uint, C.c_float, C.c_uint, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_design_nco_constrain.restype = C.c_uint
        L.emu_design_nco_constrain.argtypes = [C.c_float]
        L.emu_design_rotation.restype = C.c_float
        L.emu_design_rotation.argtypes = [C.c_uint]
        L.emu_design_agc_threshold.restype = C.c_float
        L.emu_design_agc_threshold.argtypes = [C.c_float]
        L.emu_design_firpfbch.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_void_p]
        L.emu_design_msresamp.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_frontend_geometry.argtypes = [C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return Emu(_lib)


class Emu:
    def __init__(self, L):
        self.L = L

    def frontend(self, x, rate, As=60.0, mix_mode=1, freq=0.0, quantize=1, Tc=64, nthreads=64, chunks=None, seek=0,
                 std=True, misalign=0):
        x = np.ascontiguousarray(x, np.complex64)
        chunks = [x.size] if chunks is None else list(chunks)
        ch = np.array(chunks, np.int64)
        cap = int(2 * np.ceil(rate * x.size)) + 64
        y = np.zeros(cap, np.complex64)
        n = self.L.emu_frontend(rate, As, mix_mode, freq, quantize, Tc, nthreads, x.ctypes.data, x.size, ch.ctypes.data,
                                len(chunks), y.ctypes.data, cap, seek, int(std), misalign)      # std: 0 generic kernel, 1 k_frontend_std (register prefetch), 2 k_frontend_direct, 3 k_frontend_ws
        assert n >= 0, "emu_frontend failed"
        return y[:n]

    def backend(self, x, nlanes=1, has_dc=1, has_agc=1, thr=-40.0, demod=1, kf=0.3, L=512, W=384, G=128, chunks=None,
                debug=None):
        x = np.ascontiguousarray(x, np.complex64).reshape(nlanes, -1)
        n = x.shape[1]
        chunks = [n] if chunks is None else list(chunks)
        ch = np.array(chunks, np.int64)
        out = np.zeros((nlanes, n), np.float32 if demod else np.complex64)
        fx = (C.c_ulonglong * 3)(0, 0, 0)
        self.L.emu_backend(nlanes, n, has_dc, 0.0005, has_agc, thr, demod, kf, L, W, G, x.ctypes.data, n, ch.ctypes.data,
                           len(chunks), out.ctypes.data, fx, debug.ctypes.data if debug is not None else None)
        return out, (fx[0], fx[1], fx[2])

    def pfb(self, x, M, kind, chunk_frames=None):
        """firpfbchChannelizer M through the kernel sources; kind 0 generic, 1 tile (M <= 32), 2 ring (M 128..1024)"""
        x = np.ascontiguousarray(x, np.complex64)
        nf = x.size // M
        ch = np.array([nf] if chunk_frames is None else list(chunk_frames), np.int64)
        y = np.zeros((M, nf), np.complex64)
        n = self.L.emu_pfb(M, kind, x.ctypes.data, nf, ch.ctypes.data, len(ch), y.ctypes.data)
        assert n == nf, "emu_pfb failed"
        return y

    def pfb2(self, x, M, chunk_frames=None):
        """firpfbch2 analyzer: ([M, nf] channel-major, the 2 M m taps)"""
        x = np.ascontiguousarray(x, np.complex64)
        nf = x.size // (M // 2)
        ch = np.array([nf] if chunk_frames is None else list(chunk_frames), np.int64)
        y = np.zeros((M, nf), np.complex64)
        taps = np.zeros(2 * M * 7, np.float32)
        n = self.L.emu_pfb2(M, x.ctypes.data, nf, ch.ctypes.data, len(ch), y.ctypes.data, taps.ctypes.data)
        assert n == nf, "emu_pfb2 failed"
        return y, taps

    def wbfm_tail(self, x, order, fc, M, chunks=None):
        """x: [lanes, n] float32 -> ([lanes, n_out] de-emphasised and decimated, b [nsos, 3], a [nsos, 3])"""
        x = np.ascontiguousarray(np.atleast_2d(x), np.float32)
        lanes, n = x.shape
        ch = np.array([n] if chunks is None else list(chunks), np.int64)
        y = np.zeros((lanes, n // M), np.float32)
        nsos = (order + 1) // 2
        b = np.zeros((nsos, 3), np.float32)
        a = np.zeros((nsos, 3), np.float32)
        k = self.L.emu_wbfm_tail(lanes, order, fc, M, x.ctypes.data, n, ch.ctypes.data, len(ch), y.ctypes.data,
                                 b.ctypes.data, a.ctypes.data)
        assert k >= 0, "emu_wbfm_tail failed"
        return y[:, :k], b, a

    def design_msresamp(self, rate, As=60.0):
        S, step, npfb = C.c_uint(0), C.c_uint(0), C.c_uint(0)
        ra = C.c_float(0)
        m = np.zeros(12, np.uint32)
        h1 = np.zeros((12, 32), np.float32)
        bank = np.zeros(256 * 14, np.float32)
        self.L.emu_design_msresamp(rate, As, C.byref(S), m.ctypes.data, h1.ctypes.data, C.byref(ra), C.byref(step),
                                   C.byref(npfb), bank.ctypes.data)
        return dict(S=S.value, m=[int(v) for v in m[:S.value]], h1=[h1[s, :2 * m[s]].copy() for s in range(S.value)],
                    rate_arb=ra.value, step=step.value, npfb=npfb.value, bank=bank.reshape(256, 14))

    def design_firpfbch(self, M, m=7, As=80.0):
        h = np.zeros(2 * M * m, np.float32)
        self.L.emu_design_firpfbch(M, m, As, h.ctypes.data)
        return h

    def geometry(self, rate, As, Tc):
        n = np.zeros(13, np.int32)
        d = np.zeros(13, np.int32)
        hcap, smem = C.c_int(0), C.c_int(0)
        S = self.L.emu_frontend_geometry(rate, As, Tc, n.ctypes.data, d.ctypes.data, C.byref(hcap), C.byref(smem))
        return S, n[:S + 1].tolist(), d[:S + 1].tolist(), hcap.value, smem.value
