"""The oracle against the only known answers the reference offers: numbers recovered from the
reference's own screen capture images/ex1_5.gif (README.md:195) of
    soapy_sdr -n 16000000 -f 433.9e6 -s 3.2e6 -b 1.6e6 --demod "DeNo" -g 35 -a -50 -c 20
plus structural properties every correct restatement must have.  CPU only."""
import numpy as np
import pytest

from util import make_signal

# firpfbch_crcf_print of the 20-channel, m=7, As=80 analyzer: h[249..279] (GIF frame 26)
KA_TAPS = np.array([
    -0.00394838, -0.00372026, -0.00341510, -0.00305220, -0.00265029, -0.00222706, -0.00179865, -0.00137929,
    -0.00098111, -0.00061390, -0.00028515, 0.00000000, 0.00023857, 0.00042962, 0.00057400, 0.00067408,
    0.00073350, 0.00075681, 0.00074922, 0.00071631, 0.00066374, 0.00059706, 0.00052148, 0.00044173,
    0.00036191, 0.00028547, 0.00021510, 0.00015277, 0.00009977, 0.00005670, 0.00002359])


def test_ka1_firpfbch_prototype_taps(orc):
    h = orc.Firpfbch(20).taps()
    assert h.size == 280                                   # 2*M*m taps printed
    assert np.max(np.abs(h[249:280] - KA_TAPS)) < 1.5e-8   # print precision is 1e-8


def test_ka2_nco_frequency_word(orc):
    # "Offsetting frequency by -2.984513 using VCO: nco [phase: 0x00000000 rad, freq: 0x86666600 rad/sample]"
    fb = orc.Firpfbch(20)
    assert fb.nco.phase_word == 0
    assert fb.nco.freq_word == 0x86666600


def test_ka3_dc_blocker_coefficients(orc):
    # "iir filter [normal]: b: 1.00000000 -1.00000000  a: 1.00000000 -0.99900001" (alpha was 0.001 in the GIF)
    b, a = orc.DcBlocker(0.001).coeffs()
    assert b == [1.0, -1.0]
    assert "%.8f" % a[1] == "-0.99900001"
    b, a = orc.DcBlocker(0.0005).coeffs()                  # today's value, Liquid.chs:577
    assert abs(a[1] + 0.9995) < 1e-7


def test_ka4_output_length_invariant(orc):
    # README.md:191-193: -n 16000000 -c 20 -> 20 files of 800000 CF32 samples; scaled down 1000x here:
    # 32000 input samples at r = 0.5, -n 16000 -> 20 channels x 800 samples
    x = make_signal(32000 + 4000)
    ch = orc.Chain(3.2e6, 0.0, 1.6e6, orc.DEMOD_NO, 0.0, -50.0, 20, False)
    rs = orc.MsResamp(0.5)
    # -n counts post-resample samples (takeNArr after the resampler, SoapySDR.hs:207): find the input length
    n_in = 0
    produced = 0
    while produced < 16000:
        produced += len(rs.execute(x[n_in:n_in + 1000]))
        n_in += 1000
    assert produced == 16000
    outs = ch.process(x[:n_in])
    assert len(outs) == 20 and all(len(o) == 800 for o in outs)


def test_msresamp_plan(orc):
    d = orc.MsResamp(0.078125).design()                   # config 1: 200 kHz / 2.56 MHz
    assert d["S"] == 3 and d["m"] == [10, 5, 3]
    assert d["rate_arb"] == 0.625 and d["npfb"] == 256
    assert d["step"] == 26843546                          # round(float32(2^24 / 0.625))
    d = orc.MsResamp(0.02).design()                       # config 5
    assert d["S"] == 5 and d["m"] == [10, 5, 3, 3, 3]
    assert d["step"] == 26214400


def test_msresamp_dc_gain_and_count(orc):
    rs = orc.MsResamp(0.078125)
    y = rs.execute(np.ones(80000, np.complex64))
    assert len(y) == 6250                                  # exactly r * nx once the block pipeline is full
    assert abs(y[-1].real - 1.0) < 2e-3 and abs(y[-1].imag) < 1e-6


def test_msresamp_chunk_invariance(orc):
    x = make_signal(30000, 1)
    a = orc.MsResamp(0.078125).execute(x)
    rs = orc.MsResamp(0.078125)
    b = np.concatenate([rs.execute(x[i:i + 1024]) for i in range(0, len(x), 1024)])
    assert np.array_equal(a, b)


def test_msresamp_linearity(orc):
    x1, x2 = make_signal(20000, 2), make_signal(20000, 3)
    y1, y2 = orc.MsResamp(0.3).execute(x1), orc.MsResamp(0.3).execute(x2)
    y12 = orc.MsResamp(0.3).execute(x1 + x2)
    assert np.max(np.abs(y12 - (y1 + y2))) < 1e-5


@pytest.mark.parametrize("C", [16, 20, 64])
def test_pfb_tone_lands_in_bin(orc, C):
    fb = orc.Firpfbch(C)
    k = 5
    n = np.arange(C * 300)
    f = (k - (C - 1) / 2.0) / C                            # channel centre before the pre-rotation
    y = fb.execute(np.exp(2j * np.pi * f * n).astype(np.complex64))
    mag = np.abs(y[:, -1])
    assert np.argmax(mag) == k
    others = np.delete(mag, k)
    assert others.max() < 2e-2 * mag[k]                    # NCO phase is quantised to 1024 levels (~ -55 dBc spurs)


def test_agc_converges_to_unit_level_and_gates(orc):
    n = 6000
    x = (0.02 * np.exp(2j * np.pi * 0.01 * np.arange(n))).astype(np.complex64)
    x[3000:] *= 1e-4                                        # drop 80 dB -> below the -40 dB squelch threshold
    agc = orc.Agc(-40.0)
    y = agc.execute(x)
    assert abs(np.abs(y[2500]) - 1.0) < 1e-3                # unit output level while locked
    assert np.all(y[:2] == 0)                               # ENABLED -> RISE -> SIGNALHI takes two samples
    assert np.all(y[3100:] == 0)                            # squelched after the fall


def test_freqdem_of_a_tone(orc):
    f, kf = 0.05, 0.3
    x = np.exp(2j * np.pi * f * np.arange(1000)).astype(np.complex64)
    m = orc.FreqDem(kf).execute(x)
    assert m[0] == 0.0
    assert np.max(np.abs(m[1:] - f / kf)) < 1e-5


@pytest.mark.parametrize("pll", [0, 1])
def test_ampmodem_recovers_tone(orc, pll):
    orc.set_option(orc.OPT_AMPMODEM_PLL, pll)
    try:
        n = 8000
        k = np.arange(n)
        msg = np.sin(2 * np.pi * 0.05 * k)
        x = (0.5 * (1 + 0.8 * msg)).astype(np.complex64)
        y = orc.AmpModem(0.8).execute(x)
        seg = y[4000:]
        d = 50 if pll else 25          # group delay: 25-tap-centre dc-block FIR (+ 25-sample carrier-path delay)
        c = np.dot(seg, msg[4000 - d:n - d]) / np.dot(msg[4000:], msg[4000:])
        assert 0.3 < c < 0.6
    finally:
        orc.set_option(orc.OPT_AMPMODEM_PLL, 1)


def test_chain_matches_block_composition(orc):
    """orc.Chain (the restated sdrProcess) == the blocks applied one after the other."""
    x = make_signal(40000, 5)
    f = float(np.float32(2) * np.float32(np.pi) * np.float32(1e5) / np.float32(2.56e6))
    ref = orc.FreqDem(0.3).execute(orc.Agc(-40.0).execute(orc.DcBlocker().execute(
        orc.MsResamp(np.float32(200e3 / 2.56e6)).execute(orc.Nco(f).mix_down(x)))))
    ch = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_NBFM, 0.3, -40.0)
    got = np.concatenate([ch.process(x[i:i + 7000])[0] for i in range(0, len(x), 7000)])
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("order,fc", [(2, 0.025), (2, 5000.0 / 1.0e6), (3, 0.1), (6, 0.2)])
def test_iirdes_butterworth_matches_scipy(orc, order, fc):
    """the restated liquid_iirdes (Butterworth, low-pass, second-order sections) against an independent
    implementation of the same textbook design: scipy.signal.butter with the bilinear pre-warp"""
    ss = pytest.importorskip("scipy.signal")
    f = orc.IirFiltRRRF(order, fc)
    b, a = f.coeffs()
    sos = np.concatenate([b, a], axis=1).astype(np.float64)
    w, h = ss.sosfreqz(sos, worN=512)
    _, href = ss.freqz(*ss.butter(order, 2 * fc), worN=512)
    assert np.abs(h - href).max() <= 2e-4
    x = np.random.default_rng(3).standard_normal(4000).astype(np.float32)
    y = f.execute(x)
    assert np.abs(y - ss.lfilter(*ss.butter(order, 2 * fc), x.astype(np.float64))).max() <= 2e-4 * np.abs(y).max()


def test_firdecim_is_a_decimated_convolution(orc):
    """firdecim_rrrf: 2 M m + 1 Kaiser taps, output k = the convolution sampled at input k M"""
    for M in (1, 4, 5):
        d = orc.FirDecim(M)
        h = d.taps()
        assert len(h) == 2 * M * 10 + 1 and np.allclose(h, h[::-1], atol=1e-7) and abs(h.sum() - M) < 0.02 * M
        x = np.random.default_rng(M).standard_normal(1000 * M + 3).astype(np.float32)
        y = np.concatenate([d.execute(x[:400 * M]), d.execute(x[400 * M:])])
        full = np.convolve(x.astype(np.float64), h.astype(np.float64))
        n = 400 + (len(x) - 400 * M) // M
        assert len(y) == n
        assert np.abs(y - full[0:n * M:M]).max() <= 1e-5 * np.abs(full).max()


def test_chain_wbfm_matches_block_composition(orc):
    """orc.Chain with DeWBFM == fm 0.6 -> iirFilter 2 (5000/quadRate) -> firDecimator decim after the agc
    (Liquid.chs:652-656), the decimator fed as a stream"""
    x = make_signal(60000, 6)
    f = float(np.float32(2) * np.float32(np.pi) * np.float32(1e5) / np.float32(2.56e6))
    fm = orc.FreqDem(0.6).execute(orc.Agc(-40.0).execute(orc.DcBlocker().execute(
        orc.MsResamp(np.float32(200e3 / 2.56e6)).execute(orc.Nco(f).mix_down(x)))))
    ref = orc.FirDecim(4).execute(orc.IirFiltRRRF(2, np.float32(5000.0 / 200e3)).execute(fm))
    ch = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_WBFM, 0.6, -40.0, decim=4)
    got = np.concatenate([ch.process(x[i:i + 7001])[0] for i in range(0, len(x), 7001)])
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("M", [8, 20])
def test_firpfbch2_tone_lands_in_bin_at_twice_the_channel_rate(orc, M):
    """the 2x oversampled analyzer: M/2 samples per frame, a tone at channel k's centre comes out of channel k with unit
    gain as a constant phasor (the commutator's alternation removes the (-1)^k per frame), its neighbours 6 dB down
    (cut-off 1/M), the rest suppressed"""
    k, nf = 3, 300
    n = np.arange(M // 2 * nf)
    x = np.exp(2j * np.pi * k / M * n).astype(np.complex64)
    y = orc.Firpfbch2(M).execute(x)
    mag = np.abs(y[:, -1])
    assert abs(mag[k] - 1.0) < 2e-3
    assert abs(mag[k - 1] - 0.5) < 5e-3 and abs(mag[k + 1] - 0.5) < 5e-3
    others = [c for c in range(M) if c not in (k - 1, k, k + 1)]
    assert mag[others].max() < 1e-3
    assert np.abs(y[k, -1] - y[k, -2]) < 1e-4


def test_oracle_chain_with_the_oversampled_channelizer(orc):
    """cfg.channelizer = 1 (firpfbch2_crcf in the chain: SURVEY 8f N1): the chain restatement equals its blocks composed
    by hand -- dc blocker, frames of M/2 samples through the sequential firpfbch2 object, per-channel agc + freqdem"""
    import composable_sdr_b200.synth as synth
    M = 16
    x = synth.config3(1 << 15, channels=M)
    outs = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NBFM, 0.3, -40.0, M, False, channelizer=1).process(x)
    assert len(outs) == M and len(outs[0]) == 2 * len(x) // M
    ch = orc.Firpfbch2(M).execute(orc.DcBlocker().execute(x))
    for c in (0, 3, 8, 15):
        ref = orc.FreqDem(0.3).execute(orc.Agc(-40.0).execute(ch[c]))
        assert np.array_equal(outs[c], ref)
