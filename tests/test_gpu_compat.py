"""GPU tests that CALL through libcsdr_liquid_compat.so -- the liquid-dsp symbol names src/ComposableSDR/Liquid.chs imports,
with the reference's own call sequences (INTEGRATION.md option A), against the oracle."""
import ctypes as C

import numpy as np
import pytest

from util import assert_parity, chunked

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def liq(cs):
    from composable_sdr_b200 import build
    L = C.CDLL(build.COMPAT)
    vp, u, f, i = C.c_void_p, C.c_uint, C.c_float, C.c_int
    for name, (res, args) in {
        "nco_crcf_create": (vp, [i]), "nco_crcf_destroy": (None, [vp]), "nco_crcf_set_frequency": (None, [vp, f]),
        "nco_crcf_set_phase": (None, [vp, f]), "nco_crcf_get_phase": (f, [vp]), "nco_crcf_step": (None, [vp]),
        "nco_crcf_cexpf": (None, [vp, vp]), "nco_crcf_pll_set_bandwidth": (None, [vp, f]), "nco_crcf_pll_step": (None, [vp, f]),
        "nco_crcf_mix_block_down": (None, [vp, vp, vp, u]),
        "msresamp_crcf_create": (vp, [f, f]), "msresamp_crcf_destroy": (None, [vp]), "msresamp_crcf_get_rate": (f, [vp]),
        "msresamp_crcf_execute": (None, [vp, vp, u, vp, C.POINTER(u)]),
        "iirfilt_crcf_create_dc_blocker": (vp, [f]), "iirfilt_crcf_destroy": (None, [vp]),
        "iirfilt_crcf_execute_block": (None, [vp, vp, u, vp]),
        "agc_crcf_create": (vp, []), "agc_crcf_destroy": (None, [vp]), "agc_crcf_set_bandwidth": (None, [vp, f]),
        "agc_crcf_set_signal_level": (None, [vp, f]), "agc_crcf_squelch_enable": (None, [vp]),
        "agc_crcf_squelch_set_threshold": (None, [vp, f]), "agc_crcf_squelch_set_timeout": (None, [vp, u]),
        "agc_crcf_execute_block": (None, [vp, vp, u, vp]), "agc_crcf_squelch_get_status": (i, [vp]),
        "agc_crcf_get_rssi": (f, [vp]),
        "freqdem_create": (vp, [f]), "freqdem_destroy": (None, [vp]), "freqdem_demodulate_block": (None, [vp, vp, u, vp]),
    }.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


def test_config1_through_the_liquid_names(cs, orc, liq):
    """soapy-sdr -s 2.56e6 --offset 1e5 -b 200000 --demod DeNo, block by block as Liquid.chs drives liquid:
    ncoMixDown (Liquid.chs:793-800), resample (Liquid.chs:76-98: get_rate, buffer of 2*ceil(r*nx), execute),
    iirCFilt (Liquid.chs:582-589); chunks of 1024 samples like the default --chunksize, then a few large ones"""
    x = cs.synth.config1(1 << 18)
    f = float(np.float32(2) * np.float32(np.pi) * np.float32(1e5) / np.float32(2.56e6))
    nco = liq.nco_crcf_create(1)
    assert nco, cs.CsdrError
    liq.nco_crcf_set_frequency(nco, f)
    rs = liq.msresamp_crcf_create(float(np.float32(200e3 / 2.56e6)), 60.0)
    dc = liq.iirfilt_crcf_create_dc_blocker(0.0005)
    assert rs and dc
    outs = []
    for a in chunked(x, [1024] * 40 + [50000, 7, 1 << 16]):
        a = np.ascontiguousarray(a)
        m = np.empty_like(a)
        liq.nco_crcf_mix_block_down(nco, a.ctypes.data, m.ctypes.data, a.size)
        rate = liq.msresamp_crcf_get_rate(rs)
        y = np.empty(2 * int(np.ceil(rate * a.size)), np.complex64)
        ny = C.c_uint(0)
        liq.msresamp_crcf_execute(rs, m.ctypes.data, a.size, y.ctypes.data, C.byref(ny))
        y = y[:ny.value].copy()
        z = np.empty_like(y)
        liq.iirfilt_crcf_execute_block(dc, y.ctypes.data, y.size, z.ctypes.data)
        outs.append(z)
    liq.iirfilt_crcf_destroy(dc)
    liq.msresamp_crcf_destroy(rs)
    liq.nco_crcf_destroy(nco)
    y = np.concatenate(outs)
    ref = orc.Chain(2.56e6, 1e5, 200e3).process(x)[0]
    assert len(y) == len(ref)
    assert_parity(y, ref, what="config 1 through libcsdr_liquid_compat")


def test_pilot_pll_scalar_family(cs, orc, liq):
    """pllCreate / pllStep (Liquid.chs:959-988) on a handle from the aliased nco_crcf_create: get_phase, set_phase,
    cexpf, pll_step, step and a 1-sample mix_block_down per sample -- bit for bit the oracle's nco"""
    O = orc.lib()
    f, bw = 0.35, 0.01
    pe, ss = liq.nco_crcf_create(1), liq.nco_crcf_create(1)
    liq.nco_crcf_set_frequency(pe, f)
    liq.nco_crcf_pll_set_bandwidth(pe, bw)
    liq.nco_crcf_set_frequency(ss, 2 * f)
    ope, oss = O.orc_nco_crcf_create(1), O.orc_nco_crcf_create(1)
    O.orc_nco_crcf_set_frequency(ope, f)
    O.orc_nco_crcf_pll_set_bandwidth(ope, bw)
    O.orc_nco_crcf_set_frequency(oss, 2 * f)
    g = np.random.default_rng(3)
    n = 300
    pilot = np.exp(1j * (0.3502 * np.arange(n) + 0.7)).astype(np.complex64)
    sub = (g.standard_normal(n) + 1j * g.standard_normal(n)).astype(np.complex64)
    c, oc = np.zeros(1, np.complex64), np.zeros(1, np.complex64)
    b, ob = np.zeros(1, np.complex64), np.zeros(1, np.complex64)
    for k in range(n):
        phi, ophi = liq.nco_crcf_get_phase(pe), O.orc_nco_crcf_get_phase(ope)
        assert phi == ophi
        liq.nco_crcf_set_phase(ss, 2 * phi)
        O.orc_nco_crcf_set_phase(oss, 2 * ophi)
        liq.nco_crcf_cexpf(pe, c.ctypes.data)
        O.orc_nco_crcf_cexpf(ope, oc.ctypes.data)
        assert c[0] == oc[0]
        perr = float(np.float32(np.angle(pilot[k] * np.conj(c[0]))))
        liq.nco_crcf_pll_step(pe, perr)
        O.orc_nco_crcf_pll_step(ope, perr)
        liq.nco_crcf_step(pe)
        O.orc_nco_crcf_step(ope)
        a = sub[k:k + 1].copy()
        liq.nco_crcf_mix_block_down(ss, a.ctypes.data, b.ctypes.data, 1)
        O.orc_nco_crcf_mix_block_down(oss, a.ctypes.data, ob.ctypes.data, 1)
        assert abs(b[0] - ob[0]) <= 1e-5 * abs(a[0])
    for h in (pe, ss):
        liq.nco_crcf_destroy(h)
    for h in (ope, oss):
        O.orc_nco_crcf_destroy(h)


def test_agc_per_sample_protocol_through_the_liquid_names(cs, orc, liq):
    """agcExecuteBlock (Liquid.chs:693-705): three FFI calls per sample -- execute_block(.., 1, ..), squelch_get_status,
    get_rssi -- and the Haskell gate, on the aliased symbols; then freqdem on the block"""
    n = 600
    x = cs.synth.config2(40000)[::60][:n].copy()
    h = liq.agc_crcf_create()
    liq.agc_crcf_set_bandwidth(h, 0.1)
    liq.agc_crcf_set_signal_level(h, 1e-3)
    liq.agc_crcf_squelch_enable(h)
    liq.agc_crcf_squelch_set_threshold(h, -40.0)
    liq.agc_crcf_squelch_set_timeout(h, 1000)
    y = np.zeros(n, np.complex64)
    one = np.zeros(1, np.complex64)
    for k in range(n):
        liq.agc_crcf_execute_block(h, x[k:k + 1].ctypes.data, 1, one.ctypes.data)
        st = liq.agc_crcf_squelch_get_status(h)
        liq.agc_crcf_get_rssi(h)
        y[k] = one[0] if st == 3 else 0
    liq.agc_crcf_destroy(h)
    ref = orc.Agc(-40.0).execute(x)
    assert np.array_equal(y == 0, ref == 0)
    assert_parity(y, ref, what="agc per-sample protocol through libcsdr_liquid_compat")
    fd = liq.freqdem_create(0.3)
    m = np.empty(n, np.float32)
    liq.freqdem_demodulate_block(fd, y.ctypes.data, n, m.ctypes.data)
    liq.freqdem_destroy(fd)
    assert_parity(m, orc.FreqDem(0.3).execute(y), what="freqdem through libcsdr_liquid_compat")
