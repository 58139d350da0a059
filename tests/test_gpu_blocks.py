"""GPU parity tests of the individual liquid-compatible blocks, called through the C ABI exactly as the
reference's `_process` functions would (chunk by chunk, host arrays), against the CPU oracle on the same seeded
inputs.  Tolerance: tests/util.py (peak-relative 1e-4 AND SNR >= 80 dB) unless stated."""
import numpy as np
import pytest

from util import assert_parity, chunked, make_signal

pytestmark = pytest.mark.gpu


def run_pipe(cs, pipe, x, sizes):
    process, cleanup = cs.unPipe(pipe)
    out = list(process(chunked(x, sizes)))
    cleanup()
    return out


@pytest.mark.parametrize("up", [False, True])
def test_nco_mix(cs, orc, up):
    x = make_signal(50000, 11)
    f = float(np.float32(0.24543693))
    ref = orc.Nco(f).mix_up(x) if up else orc.Nco(f).mix_down(x)
    y = np.concatenate(run_pipe(cs, cs.mixUp(f) if up else cs.mixDown(f), x, [1024, 1, 4095]))
    assert_parity(y, ref, what="nco")


def test_msresamp_generic_kernel(cs, orc):
    """the run-time-geometry kernel (used for plans other than As = 60) on the standard plan"""
    cs.set_option(5, 1)
    try:
        x = make_signal(150000, 22)
        ref = orc.MsResamp(0.078125).execute(x)
        y = np.concatenate(run_pipe(cs, cs.resampler(0.078125, 60.0), x, [50000, 1024]))
        assert_parity(y, ref, what="generic front end")
        ref = orc.MsResamp(0.1, 45.0).execute(x)            # a non-standard plan always takes this path
        y = np.concatenate(run_pipe(cs, cs.resampler(0.1, 45.0), x, [70000]))
        assert_parity(y, ref, what="As = 45 plan")
    finally:
        cs.set_option(5, 0)


@pytest.mark.parametrize("rate", [0.3, 0.15, 0.078125, 0.04, 0.02, 0.011])
def test_msresamp_register_prefetch_kernel(cs, orc, rate):
    """CSDR_OPT_FRONTEND_VARIANT = 0: k_frontend_std (raw tiles prefetched into registers, separate mixing pass, two
    CTAs per SM) instead of the default k_frontend_direct (TMA-staged tile read in place, three CTAs per SM)."""
    cs.set_option(9, 0)
    try:
        x = make_signal(300000, 23)
        ref = orc.MsResamp(rate).execute(x)
        a = np.concatenate(run_pipe(cs, cs.resampler(rate, 60.0), x, [len(x)]))
        b = np.concatenate(run_pipe(cs, cs.resampler(rate, 60.0), x, [1024 * 9, 77]))
        assert np.array_equal(a, b)
        assert_parity(a, ref, what=f"msresamp (k_frontend_std) {rate}")
    finally:
        cs.set_option(9, 1)


@pytest.mark.parametrize("rate", [0.3, 0.15, 0.078125, 0.04, 0.02, 0.011])
def test_msresamp_warp_specialised_kernel(cs, orc, rate):
    """CSDR_OPT_FRONTEND_VARIANT = 2: k_frontend_ws (producer / consumer warp groups on different tiles)"""
    cs.set_option(9, 2)
    try:
        x = make_signal(300000, 29)
        ref = orc.MsResamp(rate).execute(x)
        a = np.concatenate(run_pipe(cs, cs.resampler(rate, 60.0), x, [len(x)]))
        b = np.concatenate(run_pipe(cs, cs.resampler(rate, 60.0), x, [1024 * 9, 77]))
        assert np.array_equal(a, b)
        assert_parity(a, ref, what=f"msresamp (k_frontend_ws) {rate}")
    finally:
        cs.set_option(9, 1)


@pytest.mark.parametrize("rate", [1.25, 2.0, 3.7, 10.0])
def test_msresamp_interpolation(cs, orc, rate):
    """rate > 1 (MSRESAMP(_interp_execute)): arbitrary stage, then half-band interpolators"""
    x = make_signal(50000, 31)
    ref = orc.MsResamp(rate).execute(x)
    a = np.concatenate(run_pipe(cs, cs.resampler(rate, 60.0), x, [1024]))
    b = np.concatenate(run_pipe(cs, cs.resampler(rate, 60.0), x, [len(x)]))
    assert len(a) == len(ref)
    assert np.array_equal(a, b)
    assert_parity(a, ref, what=f"msresamp interpolation {rate}")


@pytest.mark.parametrize("rate", [0.078125, 0.02, 0.625, 0.5, 0.3, 0.15, 0.04, 0.011])
def test_msresamp(cs, orc, rate):
    x = make_signal(200000, 12)
    ref = orc.MsResamp(rate).execute(x)
    a = np.concatenate(run_pipe(cs, cs.resampler(rate, 60.0), x, [1024]))       # the reference's chunk size
    b = np.concatenate(run_pipe(cs, cs.resampler(rate, 60.0), x, [len(x)]))
    c = np.concatenate(run_pipe(cs, cs.resampler(rate, 60.0), x, [1, 7, 100000, 33]))
    assert len(a) == len(ref) and len(b) == len(ref)
    assert np.array_equal(a, b) and np.array_equal(a, c)       # any chunking gives the same output stream, bit for bit
    assert_parity(a, ref, what=f"msresamp {rate}")


def test_msresamp_output_count_per_call_matches_liquid(cs, orc):
    """the Haskell side sizes arrays from *ny (Liquid.chs:84-98): every call must return the oracle's count"""
    x = make_signal(30000, 13)
    rs = orc.MsResamp(0.078125)
    want = [len(rs.execute(c)) for c in chunked(x, [1000, 13, 2048])]
    got = [len(c) for c in run_pipe(cs, cs.resampler(0.078125, 60.0), x, [1000, 13, 2048])]
    assert got == want


def test_dc_blocker(cs, orc):
    x = make_signal(300000, 14)
    ref = orc.DcBlocker().execute(x)
    y = np.concatenate(run_pipe(cs, cs.dcBlocker(), x, [65536, 1024, 100]))
    assert_parity(y, ref, what="dcBlocker")


@pytest.mark.parametrize("C", [2, 4, 8, 12, 16, 20, 24, 32, 64, 128, 512, 1024])
def test_firpfbch(cs, orc, C):
    """every channelizer kernel: k_pfb_tile (2, 4, 32), k_pfb_tile2 (8, 16), k_pfb_tile_any (12, 20, 24), generic k_pfb (64),
    k_pfb_stream (128, 512, 1024)"""
    nf = 600 if C < 128 else 96
    x = make_signal(C * nf, 15)
    ref = orc.Firpfbch(C).execute(x)
    sizes = [C * 100, C * 7, C * 300]
    outs = run_pipe(cs, cs.firpfbchChannelizer(C), x, sizes)
    y = np.concatenate([np.stack(o) for o in outs], axis=1)
    assert_parity(y, ref, what=f"firpfbch {C}")


def test_firpfbch_tail_samples_are_dropped_like_the_reference(cs, orc):
    C = 16
    x = make_signal(C * 50 + 5, 16)
    ref = orc.Firpfbch(C).execute(x)
    y = np.stack(run_pipe(cs, cs.firpfbchChannelizer(C), x, [len(x)])[0])
    assert y.shape == ref.shape == (C, 50)
    assert_parity(y, ref, what="firpfbch tail")


def keyed_fm(n, seed=0):
    t = np.arange(n)
    g = np.random.default_rng(seed)
    ph = 2 * np.pi * (0.11 * t + 0.02 * np.cumsum(np.sin(2 * np.pi * t / 400)))
    env = ((t // 7000) % 2 == 0)
    return (0.3 * env * np.exp(1j * ph) + 0.001 * (g.standard_normal(n) + 1j * g.standard_normal(n))).astype(np.complex64)


def test_agc_with_squelch_gate(cs, orc):
    x = keyed_fm(120000, 17)
    ref = orc.Agc(-40.0).execute(x)
    y = np.concatenate(run_pipe(cs, cs.automaticGainControl(-40.0), x, [4096, 50000, 1024]))
    mism = np.count_nonzero((y == 0) != (ref == 0))
    assert mism == 0, f"{mism} samples gated differently"
    assert_parity(y, ref, what="agc")


def test_agc_per_sample_protocol(cs, orc):
    """the reference calls execute_block(.., 1, ..) + squelch_get_status + get_rssi per sample (Liquid.chs:697-704)"""
    from composable_sdr_b200 import _lib
    L = _lib.load()
    x = keyed_fm(300, 18)
    x[:200] *= 1.0
    a = orc.Agc(-40.0)
    h = L.csdr_agc_crcf_create()
    L.csdr_agc_crcf_set_bandwidth(h, 0.1)
    L.csdr_agc_crcf_set_signal_level(h, 1e-3)
    L.csdr_agc_crcf_squelch_enable(h)
    L.csdr_agc_crcf_squelch_set_threshold(h, -40.0)
    L.csdr_agc_crcf_squelch_set_timeout(h, 1000)
    y = np.zeros(1, np.complex64)
    for i in range(len(x)):
        r = a.execute_raw(x[i:i + 1])
        L.csdr_agc_crcf_execute_block(h, x[i:i + 1].ctypes.data, 1, y.ctypes.data)
        assert L.csdr_agc_crcf_squelch_get_status(h) == a.status
        assert abs(L.csdr_agc_crcf_get_rssi(h) - a.rssi) < 1e-3
        assert abs(y[0] - r[0]) <= 1e-4 * max(1.0, abs(r[0]))
    L.csdr_agc_crcf_destroy(h)


def test_freqdem(cs, orc):
    x = keyed_fm(100000, 19)
    ref = orc.FreqDem(0.3).execute(x)
    y = np.concatenate(run_pipe(cs, cs.fmDemodulator(0.3), x, [1024, 9999]))
    # the discriminator of a noise-only sample pair is ill-conditioned in |r|: compare where there is signal
    assert_parity(y, ref, rel=2e-4, what="freqdem")


@pytest.mark.parametrize("pll", [1, 0])
def test_ampmodem(cs, orc, pll):
    from composable_sdr_b200 import _lib
    cs.set_option(_lib.OPT_AMPMODEM_PLL, pll)
    orc.set_option(orc.OPT_AMPMODEM_PLL, pll)
    try:
        n = 60000
        k = np.arange(n)
        g = np.random.default_rng(20)
        x = (0.5 * (1 + 0.8 * np.sin(2 * np.pi * 0.03 * k)) * np.exp(2j * np.pi * 0.0004 * k + 0.5j)
             + 0.002 * (g.standard_normal(n) + 1j * g.standard_normal(n))).astype(np.complex64)
        ref = orc.AmpModem(0.8).execute(x)
        y = np.concatenate(run_pipe(cs, cs.amDemodulator(), x, [1024, 20000, 30]))
        # the PLL is a feedback loop through a 1024-level phase quantiser: allow 3e-4 of peak (SNR still >= 80 dB)
        assert_parity(y, ref, rel=3e-4 if pll else 1e-4, what=f"ampmodem pll={pll}")
    finally:
        cs.set_option(_lib.OPT_AMPMODEM_PLL, 1)
        orc.set_option(orc.OPT_AMPMODEM_PLL, 1)


def test_device_pointers_are_accepted(cs, orc):
    import torch
    x = make_signal(100000, 21)
    ref = orc.MsResamp(0.078125).execute(x)
    xd = torch.from_numpy(x).cuda()
    y = torch.cat(run_pipe(cs, cs.resampler(0.078125, 60.0), xd, [40000])).cpu().numpy()
    assert_parity(y, ref, what="device pointers")


def test_unfused_app_graph_config1(cs, orc):
    """sdrProcess assembled from the individual blocks with the reference's chunk protocol (chunk 1024,
    compact 4096) == the oracle's chain.  Config 1: -s 2.56e6 --offset 1e5 -b 200000 --demod DeNo."""
    x = cs.synth.config1(400000)
    ref = orc.Chain(2.56e6, 1e5, 200e3).process(x)[0][:30000]
    y = cs.sdrProcess(chunked(x, [1024]), 2.56e6, offset=1e5, bandwidth=200e3, numsamples=30000)
    assert_parity(y, ref, what="unfused config 1")


@pytest.mark.parametrize("order,fc", [(2, 0.025), (2, 0.005), (5, 0.1)])
def test_iirfilt_rrrf_butterworth(cs, orc, order, fc):
    """iirFilter n fc 0 10 10 (Liquid.chs:644-650): design and the parallel-scan execution against the sequential
    direct form II, across chunk sizes that are not multiples of the 256-sample segments"""
    x = np.random.default_rng(31).standard_normal(300000).astype(np.float32) + np.float32(0.25)
    o = orc.IirFiltRRRF(order, fc)
    ref = o.execute(x)
    y = np.concatenate(run_pipe(cs, cs.iirFilter(order, fc, 0.0, 10.0, 10.0), x, [1024, 9999, 1, 255, 100001]))
    assert_parity(y, ref, what=f"iirfilt_rrrf order {order} fc {fc}")
    L = cs._lib.load()
    h = L.csdr_iirfilt_rrrf_create_prototype(0, 0, 0, order, fc, 0.0, 10.0, 10.0)
    b = np.zeros(((order + 1) // 2, 3), np.float32)
    a = np.zeros_like(b)
    assert L.csdr_iirfilt_rrrf_coefficients(h, b.ctypes.data, a.ctypes.data) == (order + 1) // 2
    L.csdr_iirfilt_rrrf_destroy(h)
    bo, ao = o.coeffs()
    assert np.abs(b - bo).max() <= 2e-5 * np.abs(bo).max() and np.abs(a - ao).max() <= 2e-6
    # only the family the reference asks for exists; anything else fails loudly
    assert not L.csdr_iirfilt_rrrf_create_prototype(1, 0, 0, 2, fc, 0.0, 10.0, 10.0)
    assert b"Butterworth" in L.csdr_last_error()


@pytest.mark.parametrize("M", [1, 4, 10])
def test_firdecim_rrrf(cs, orc, M):
    """firDecimator m (Liquid.chs:487-503): arrays whose length is a multiple of m (the reference drops the rest)"""
    x = np.random.default_rng(32).standard_normal(40000 * M).astype(np.float32)
    ref = orc.FirDecim(M).execute(x)
    y = np.concatenate(run_pipe(cs, cs.firDecimator(M), x, [1024 * M, 3 * M, 7777 * M]))
    assert_parity(y, ref, what=f"firdecim_rrrf M={M}")


def test_unfused_wbfm_demodulator(cs, orc):
    """wbFMDemodulator quadRate decim = firDecimator . iirFilter . fmDemodulator 0.6 (Liquid.chs:652-656), block by
    block with the reference's 1024-sample arrays"""
    x = keyed_fm(200 * 1024, 23)
    ref = orc.FirDecim(4).execute(orc.IirFiltRRRF(2, np.float32(5000.0 / 200e3)).execute(orc.FreqDem(0.6).execute(x)))
    y = np.concatenate(run_pipe(cs, cs.wbFMDemodulator(200e3, 4), x, [1024]))
    assert_parity(y, ref, rel=2e-4, what="wbFMDemodulator")


@pytest.mark.parametrize("M", [2, 4, 8, 12, 16, 20, 32, 64, 128, 1024])
def test_firpfbch2_oversampled_analyzer(cs, orc, M):
    """firpfbch2_crcf (2x oversampled): coarse block call in chunks of odd and even frame counts, and
    liquid's per-frame execute, against the oracle's sequential object"""
    nf = 45 if M >= 128 else 400
    x = make_signal(M // 2 * nf, 29)
    ref = orc.Firpfbch2(M).execute(x)
    M2 = M // 2
    outs = run_pipe(cs, cs.firpfbch2Channelizer(M), x, [M2 * 7, M2 * 30, M2])
    y = np.stack([np.concatenate([o[c] for o in outs]) for c in range(M)])
    assert_parity(y, ref, what=f"firpfbch2 M={M}")
    L = cs._lib.load()
    h = L.csdr_firpfbch2_crcf_create_kaiser(0, M, 7, 80.0)
    fr = np.zeros((5, M), np.complex64)
    for t in range(5):
        L.csdr_firpfbch2_crcf_execute(h, x[t * M2:].ctypes.data, fr[t].ctypes.data)
    L.csdr_firpfbch2_crcf_destroy(h)
    assert_parity(fr.T, ref[:, :5], what=f"firpfbch2 M={M}, frame by frame")
    assert not L.csdr_firpfbch2_crcf_create_kaiser(0, 7, 7, 80.0) and b"even" in L.csdr_last_error()


def test_firpfbch2_cluster_pair_matches_two_launches(cs, orc):
    """M = 1024: the even / odd passes as clusters of two CTAs (default) and as two launches (CSDR_OPT_PFB_VARIANT = 3) give the
    same bits, for an odd and an even number of frames per call"""
    M, M2 = 1024, 512
    x = make_signal(M2 * 131, 31)
    outs = {}
    for variant in (0, 3):
        cs.set_option(10, variant)
        try:
            o = run_pipe(cs, cs.firpfbch2Channelizer(M), x, [M2 * 33, M2 * 64, M2 * 1])
        finally:
            cs.set_option(10, 0)
        outs[variant] = np.stack([np.concatenate([q[c] for q in o]) for c in range(M)])
    assert outs[0].shape == outs[3].shape
    assert np.array_equal(outs[0], outs[3])
    assert_parity(outs[0], orc.Firpfbch2(M).execute(x), what="firpfbch2 M=1024, cluster pair")
