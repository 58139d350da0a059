"""GPU parity tests of the fused chain (csdr_chain_*, the object bench.py drives) against the oracle's restatement
of sdrProcess (apps/SoapySDR.hs:181-283) for the five BASELINE configs, plus size-independent properties at
BASELINE's full sizes.  Tolerance: tests/util.py."""
import numpy as np
import pytest

from util import REL_TOL_AFTER_DCBLOCK, REL_TOL_FM_NOISE, assert_parity, chunked, make_signal, snr_db

pytestmark = pytest.mark.gpu


def _ranges(n, sizes):
    pos, i = 0, 0
    while pos < n:
        m = sizes[i % len(sizes)]
        yield pos, min(n, pos + m)
        pos += m
        i += 1


def run_chain(chain, x, sizes):
    """feed x ([n] or [nstreams, n]) in chunks of the given sizes; returns the concatenated outputs"""
    acc = None
    for i, j in _ranges(x.shape[-1], sizes):
        o = chain.process(np.ascontiguousarray(x[..., i:j]))
        acc = [[v] for v in o] if acc is None else [a + [v] for a, v in zip(acc, o)]
    return [np.concatenate(a) for a in acc]


def test_config1_mix_resample(cs, orc):
    """soapy-sdr -s 2.56e6 --offset 1e5 -b 200000 --demod DeNo  (mix + msresamp [+ dc blocker])"""
    x = cs.synth.config1(1 << 21)
    ref = orc.Chain(2.56e6, 1e5, 200e3).process(x)[0]
    a = run_chain(cs.Chain(2.56e6, 1e5, 200e3), x, [1 << 20])[0]
    b = run_chain(cs.Chain(2.56e6, 1e5, 200e3), x, [1024 * 37, 5, 1 << 19])[0]
    assert len(a) == len(ref) == len(b)
    assert_parity(a, ref, what="config 1")
    assert_parity(b, ref, what="config 1 (ragged chunks)")


def test_config2_fm_with_agc(cs, orc):
    """2.56 MS/s -> 200 kHz, NBFM, AGC -40 dB, carrier keyed on/off (squelch).  The bench workload."""
    x = cs.synth.config2(1 << 22)
    ref = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_NBFM, 0.3, -40.0).process(x)[0]
    ch = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
    y = run_chain(ch, x, [1 << 21])[0]
    assert len(y) == len(ref)
    # the squelch gate must open and close on the same samples
    assert np.count_nonzero((y == 0) != (ref == 0)) == 0
    assert_parity(y, ref, rel=REL_TOL_AFTER_DCBLOCK, what="config 2")
    y2 = run_chain(cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0), x, [300001, 1024])[0]
    assert_parity(y2, ref, rel=REL_TOL_AFTER_DCBLOCK, what="config 2 (ragged chunks)")


def test_config2_large_chunk(cs, orc):
    """a 2^24-sample chunk, plain and with CSDR_OPT_OVERLAP (back end of part i on a second stream while the front
    end filters part i+1): same result as the oracle"""
    x = cs.synth.config2(1 << 24)
    ref = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_NBFM, 0.3, -40.0).process(x)[0]
    y = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0).process(x)[0]
    assert len(y) == len(ref)
    assert np.count_nonzero((y == 0) != (ref == 0)) == 0
    per = 1.0 / 0.3                                         # 2 pi of the discriminator in output units (1 / kf)
    assert_parity(y, ref, rel=REL_TOL_AFTER_DCBLOCK, period=per, what="config 2, 2^24-sample chunk")
    cs.set_option(7, 1)
    try:
        y2 = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0).process(x)[0]
    finally:
        cs.set_option(7, 0)
    assert_parity(y2, ref, rel=REL_TOL_AFTER_DCBLOCK, period=per, what="config 2, 2^24-sample chunk, overlapped")
    # a host chunk this large is fed in pipelined parts (copy of part i+1 under the kernels of part i); a ragged
    # length leaves a short last part, and the same samples resident on the device take the single-launch path
    import torch
    m = (1 << 24) + 12345
    xr = np.concatenate([x, x[:12345]])
    yh = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0).process(xr)[0]
    yd = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0).process(torch.from_numpy(xr).cuda())[0].cpu().numpy()
    assert len(yh) == len(yd) and abs(len(yh) - m * 5 / 64) <= 1
    assert_parity(yh[:len(ref)], ref, rel=REL_TOL_AFTER_DCBLOCK, period=per, what="pipelined host chunk")
    assert_parity(yh, yd, rel=REL_TOL_AFTER_DCBLOCK, period=per, what="pipelined host chunk vs device chunk")


def test_config3_channelizer_per_channel_fm(cs, orc):
    """2.56 MS/s into -c 16, per-channel AGC + NBFM, 16 outputs"""
    x = cs.synth.config3(1 << 20)
    ref = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NBFM, 0.3, -40.0, 16, False).process(x)
    outs = run_chain(cs.Chain(2.56e6, demod=cs.DeNBFM(0.3), agc=-40.0, channels=16), x, [1 << 19, 65536 + 3])
    assert len(outs) == 16
    for c in range(16):
        assert len(outs[c]) == len(ref[c])
        assert np.count_nonzero((outs[c] == 0) != (ref[c] == 0)) <= 2
        assert_parity(outs[c], ref[c], rel=REL_TOL_FM_NOISE, period=1 / 0.3, what=f"config 3 channel {c}")


def test_config3_raw_channels_and_mix(cs, orc):
    x = cs.synth.config3(1 << 18)
    ref = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NO, 0.0, 0.0, 16, False).process(x)
    outs = run_chain(cs.Chain(2.56e6, channels=16), x, [100000])
    for c in range(16):
        assert_parity(outs[c], ref[c], what=f"config 3 DeNo channel {c}")
    refm = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NBFM, 0.3, -40.0, 16, True).process(x)[0]
    m = run_chain(cs.Chain(2.56e6, demod=cs.DeNBFM(0.3), agc=-40.0, channels=16, mix_channels=True), x, [100000])[0]
    assert_parity(m, refm, rel=REL_TOL_FM_NOISE, what="config 3 --mix")


def test_mix_variants_sum_the_right_buffers(cs, orc):
    """--mix (Trans.hs:119-122) behind every kind of summand: gated cf32 (AGC, DeNo: the gate-aware sum over sample pairs),
    ungated cf32 (no AGC: the plain sum), and -- once -- the plain sum forced for the gated discriminator case must equal
    the gate-aware one bit for bit except for the sign of a zero"""
    x = cs.synth.config3(1 << 18)
    refc = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NO, 0.0, -40.0, 16, True).process(x)[0]
    yc = run_chain(cs.Chain(2.56e6, agc=-40.0, channels=16, mix_channels=True), x, [100000, 33])[0]
    assert len(yc) == len(refc)
    assert_parity(yc[64:], refc[64:], what="--mix of gated cf32 channels")
    refn = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NO, 0.0, 0.0, 16, True).process(x)[0]
    yn = run_chain(cs.Chain(2.56e6, channels=16, mix_channels=True), x, [100000, 33])[0]
    assert_parity(yn, refn, what="--mix of raw channels")
    # the sum of the individually produced channels (no --mix) equals the chain's own --mix output
    chans = run_chain(cs.Chain(2.56e6, demod=cs.DeNBFM(0.3), agc=-40.0, channels=16), x, [100000, 33])
    fold = chans[0].copy()
    for c in range(1, 16):
        fold = (fold + chans[c]).astype(np.float32)
    ym = run_chain(cs.Chain(2.56e6, demod=cs.DeNBFM(0.3), agc=-40.0, channels=16, mix_channels=True), x, [100000, 33])[0]
    assert np.array_equal(fold, ym)


def _gate_report(y, ref):
    """(number of samples whose squelch gate differs, mask of samples whose gate agrees)"""
    gy, gr = (np.asarray(y) != 0), (np.asarray(ref) != 0)
    return int(np.count_nonzero(gy != gr)), gy == gr


def test_config4_wideband_1024_channels_predemod(cs, orc):
    """config 4 at SURVEY 8(d)'s parity size (2^24 samples = 16384 frames of 1024 channels), BEFORE the demodulator:
    every channel's AGC output (CF32, gated) against the oracle -- gate bits exactly, samples to 80 dB / 1e-4."""
    x = cs.synth.config4(1 << 24)
    ref = orc.Chain(1e9, 0.0, 0.0, orc.DEMOD_NO, 0.0, -40.0, 1024, False).process(x)
    outs = run_chain(cs.Chain(1e9, agc=-40.0, channels=1024), x, [1 << 23, (1 << 22) + 1024 * 3 + 5, 1 << 24])
    assert len(outs) == 1024
    skip = 512        # the filterbank's switch-on click opens every squelch; where each one closes is compared too
    mism, worst, active = 0, np.inf, 0
    for c in range(1024):
        assert len(outs[c]) == len(ref[c]) == 16384
        k, same = _gate_report(outs[c], ref[c])
        mism += k
        r = ref[c][skip:]
        if np.count_nonzero(r) > r.size // 2:
            active += 1
            m = same[skip:]
            s_db = snr_db(outs[c][skip:][m], r[m])
            worst = min(worst, s_db)
            assert_parity(outs[c][skip:][m], r[m], what=f"config 4 pre-demod, channel {c}")
    assert active == 64
    assert mism == 0, f"{mism} gate bits differ over 1024 channels x 16384 frames"
    assert worst >= 80.0


def test_config4_wideband_1024_channels_mix(cs, orc):
    """1024-channel firpfbch with --mix at SURVEY 8(d)'s parity size (2^24 samples), per-channel AGC + NBFM summed
    over the channels, against the oracle on one GPU: SNR >= 80 dB (the north star's tolerance)"""
    x = cs.synth.config4(1 << 24)
    ref = orc.Chain(1e9, 0.0, 0.0, orc.DEMOD_NBFM, 0.3, -40.0, 1024, True).process(x)[0]
    y = run_chain(cs.Chain(1e9, demod=cs.DeNBFM(0.3), agc=-40.0, channels=1024, mix_channels=True), x, [1 << 23])[0]
    assert len(y) == len(ref) == 16384
    # the switch-on click of the filterbank opens every squelch for ~170 frames (compared above, before the
    # demodulator); the sum of 1024 discriminators is compared behind it
    s_db = snr_db(y[512:], ref[512:])
    assert s_db >= 80.0, f"config 4 --mix: SNR {s_db:.1f} dB"
    assert_parity(y[512:], ref[512:], rel=REL_TOL_FM_NOISE, what="config 4 --mix")


def _assert_channels(outs, ref, what, fm, rel, max_gate_mismatch=2, skip=64):
    """per channel: squelch gates agree (up to `max_gate_mismatch` threshold-marginal samples); channels that carry a
    signal (gate open most of the time) match to the stated tolerance.  A channel that holds only noise opens for a few
    samples behind the filterbank's switch-on click: what it shows there is the stop-band leakage of the carriers in
    OTHER channels (~80 dB down), i.e. float32 rounding of the polyphase sums (1e-7 of the wide-band level) is ~1e-3 of
    the channel's own level -- those few samples are held to 40 dB only."""
    assert len(outs) == len(ref)
    for c in range(len(ref)):
        assert len(outs[c]) == len(ref[c]) and len(ref[c]) > skip + 1000, (what, c, len(outs[c]), len(ref[c]))
        k, same = _gate_report(outs[c], ref[c])
        assert k <= max_gate_mismatch, f"{what} channel {c}: {k} gate bits differ"
        if not np.any(ref[c][skip:]):
            assert not np.any(outs[c][skip:])
            continue
        m = same[skip:]
        if fm:
            m = m & np.concatenate([[True], m[:-1]])          # the discriminator looks at the previous sample as well
        carries_signal = np.count_nonzero(ref[c][skip:]) > 0.4 * (len(ref[c]) - skip)
        if carries_signal:
            assert_parity(outs[c][skip:][m], ref[c][skip:][m], rel=rel, period=(1 / 0.3) if fm else None, what=f"{what} channel {c}")
        else:
            assert_parity(outs[c][skip:][m], ref[c][skip:][m], rel=3e-2, snr=40.0, period=(1 / 0.3) if fm else None,
                          what=f"{what} channel {c} (noise only)")


def test_readme_example3_resample_then_channelize(cs, orc):
    """the reference's one published scenario (README.md:182-193, graph apps/SoapySDR.hs:206-226):
    soapy-sdr -s 3.2e6 -b 1.6e6 -a -50 -c 20 --demod DeNo -- resampler AND channelizer in one chain, a channel count
    that is not a power of two, 20 CF32 outputs of n * 0.5 / 20 samples each"""
    n = 1 << 22
    x = cs.synth.example3(n)
    ref = orc.Chain(3.2e6, 0.0, 1.6e6, orc.DEMOD_NO, 0.0, -50.0, 20, False).process(x)
    ch = cs.Chain(3.2e6, 0.0, 1.6e6, agc=-50.0, channels=20)
    outs = run_chain(ch, x, [1 << 21, 300001, 4099, 1 << 22])
    assert len(outs) == 20 and len(outs[0]) == n // 2 // 20
    _assert_channels(outs, ref, "README example 3", fm=False, rel=1e-4)
    # whole input in one call: same result
    outs1 = run_chain(cs.Chain(3.2e6, 0.0, 1.6e6, agc=-50.0, channels=20), x, [n])
    _assert_channels(outs1, ref, "README example 3 (one chunk)", fm=False, rel=1e-4)


def test_offset_resample_channelize_fm(cs, orc):
    """offset mix + resampler + 16-channel channelizer + per-channel AGC + NBFM in ONE chain (the path the metric
    names: mix -> resample -> PFB -> demod), and its --mix sum"""
    n = 1 << 22
    x = cs.synth.example3_offset(n)
    ref = orc.Chain(2.56e6, 1e5, 1.28e6, orc.DEMOD_NBFM, 0.3, -40.0, 16, False).process(x)
    outs = run_chain(cs.Chain(2.56e6, 1e5, 1.28e6, cs.DeNBFM(0.3), agc=-40.0, channels=16), x, [1 << 21, 777777, 1 << 22])
    assert len(outs) == 16 and len(outs[0]) == n // 2 // 16
    _assert_channels(outs, ref, "mix+resample+PFB+FM", fm=True, rel=REL_TOL_FM_NOISE)
    refm = orc.Chain(2.56e6, 1e5, 1.28e6, orc.DEMOD_NBFM, 0.3, -40.0, 16, True).process(x)[0]
    ym = run_chain(cs.Chain(2.56e6, 1e5, 1.28e6, cs.DeNBFM(0.3), agc=-40.0, channels=16, mix_channels=True), x, [n])[0]
    assert len(ym) == len(refm)
    # the sum inherits the few threshold-marginal gate samples of the fading channels: compare where all gates agree
    agree = np.ones(len(refm), bool)
    for c in range(16):
        k, same = _gate_report(outs[c], ref[c])
        agree &= same & np.concatenate([[True], same[:-1]])
    assert np.count_nonzero(~agree) <= 64
    # (a sum of discriminators is ambiguous by whole turns of any of them: arg() of a noise-only sample pair next to +-pi)
    assert_parity(ym[64:][agree[64:]], refm[64:][agree[64:]], rel=REL_TOL_FM_NOISE, period=1 / 0.3, what="mix+resample+PFB+FM --mix")


def test_config5_batch_of_streams_am(cs, orc):
    """independent 10 MS/s captures: offset mix + resample + AGC + AM demod, one handle for all streams.

    liquid's DSB demodulator is a carrier PLL through a 1024-level phase quantiser: a perturbation of its INPUT by 1e-7
    (137 dB below the signal) moves the oracle's own output by -60 dB -- measured below on the oracle alone -- so the
    chain is held (a) to 80 dB in front of the demodulator, and (b) through the demodulator block alone, fed the
    oracle's samples, and (c) end to end, to what that sensitivity allows."""
    S, n = 8, 1 << 21
    x = cs.synth.config5(n, S)
    skip = 15000                     # AGC attack + carrier PLL pull-in (437 Hz offset, loop bandwidth 1e-3)
    outs = run_chain(cs.Chain(10e6, 1e6, 200e3, cs.DeAM(), agc=-40.0, nstreams=S), x, [n // 2, 100000, 1 << 20])
    pres = run_chain(cs.Chain(10e6, 1e6, 200e3, agc=-40.0, nstreams=S), x, [n // 2, 100000, 1 << 20])
    g = np.random.default_rng(1)
    for s in range(S):
        ref = orc.Chain(10e6, 1e6, 200e3, orc.DEMOD_AM, 0.0, -40.0).process(x[s])[0]
        pre = orc.Chain(10e6, 1e6, 200e3, orc.DEMOD_NO, 0.0, -40.0).process(x[s])[0]
        assert len(outs[s]) == len(ref) and abs(len(ref) - n / 50) <= 1
        assert len(ref) - skip >= 25000   # (the comparisons below must not be empty: 41 943 outputs per stream)
        assert np.abs(ref[skip:]).max() > 0.1
        # (a) in front of the demodulator
        assert np.array_equal(pres[s] == 0, pre == 0)
        assert_parity(pres[s][64:], pre[64:], rel=REL_TOL_AFTER_DCBLOCK, what=f"config 5 stream {s}, AGC output")
        if s < 2:
            # (c) the oracle's own sensitivity to a 1e-7 perturbation of the demodulator's input
            pert = (pre * (1 + 1e-7 * (g.standard_normal(pre.size) + 1j * g.standard_normal(pre.size)))).astype(np.complex64)
            floor_db = snr_db(orc.AmpModem(0.8).execute(pert)[skip:], ref[skip:])
            assert 50.0 < floor_db < 75.0, floor_db
            # (b) the demodulator block on the oracle's own samples: its 51-tap filters sum in a different order (1e-7),
            # which is such a perturbation
            am = cs.amDemodulator()
            r = am._start()
            yb = np.concatenate([am._process(r, pre[:20000]), am._process(r, pre[20000:])])
            am._done(r)
            assert snr_db(yb[skip:], ref[skip:]) >= floor_db - 6.0
            assert_parity(yb[skip:], ref[skip:], rel=1e-2, snr=50.0, what=f"config 5 stream {s}, ampmodem on the oracle's samples")
        got = snr_db(outs[s][skip:], ref[skip:])
        assert got >= 55.0, f"config 5 stream {s}: SNR {got:.1f} dB"
        assert_parity(outs[s][skip:], ref[skip:], rel=1e-2, snr=55.0, what=f"config 5 stream {s}")


def test_chain_with_the_oversampled_channelizer(cs, orc):
    """channelizer = firpfbch2_crcf (liquid's 2x oversampled analyzer as the chain's channelizer block):
    16 channels at 2/16 of the input rate each, per-channel AGC + NBFM, ragged chunks (left-overs of less than half a
    frame), and the --mix sum, against the oracle's chain with the sequential firpfbch2 object"""
    x = cs.synth.config3(1 << 20)
    ref = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NBFM, 0.3, -40.0, 16, False, channelizer=1).process(x)
    outs = run_chain(cs.Chain(2.56e6, demod=cs.DeNBFM(0.3), agc=-40.0, channels=16, channelizer=1), x, [1 << 19, 65536 + 3, 5, 1 << 20])
    assert len(outs) == 16 and len(outs[0]) == len(ref[0]) == (1 << 20) // 8
    _assert_channels(outs, ref, "firpfbch2 chain", fm=True, rel=REL_TOL_FM_NOISE)
    refc = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NO, 0.0, 0.0, 16, False, channelizer=1).process(x[:1 << 18])
    outc = run_chain(cs.Chain(2.56e6, channels=16, channelizer=1), x[:1 << 18], [100000, 7])
    for c in range(16):
        assert_parity(outc[c], refc[c], what=f"firpfbch2 chain, DeNo channel {c}")
    # 20 channels (not a power of two) behind the resampler, as in README example 3
    x3 = cs.synth.example3(1 << 20)
    r3 = orc.Chain(3.2e6, 0.0, 1.6e6, orc.DEMOD_NO, 0.0, 0.0, 20, False, channelizer=1).process(x3)
    o3 = run_chain(cs.Chain(3.2e6, 0.0, 1.6e6, channels=20, channelizer=1), x3, [300001, 1 << 20])
    level = max(float(np.abs(r3[c][64:]).max()) for c in range(20))
    for c in range(20):
        assert len(o3[c]) == len(r3[c]) == (1 << 20) // 2 // 10
        # channels that carry a signal to the stated tolerance; a channel that holds only noise and the stop-band leakage
        # of the others sits 60+ dB below them: float32 rounding of the polyphase sums (1e-7 of the wide-band level) is
        # 1e-4 of ITS level, so it is held to 1e-6 of the filterbank's output level instead
        if float(np.abs(r3[c][64:]).max()) > 0.05 * level:
            assert_parity(o3[c][64:], r3[c][64:], rel=2e-4, what=f"firpfbch2 chain behind the resampler, channel {c}")
        else:
            assert float(np.abs(o3[c][64:] - r3[c][64:]).max()) <= 1e-6 * level, f"noise-only channel {c}"


def test_time_segment_sharding_matches_single_stream(cs, orc):
    """multi-GPU partitioning of one long stream: a shard seeks to its start minus the warm-up, feeds
    the warm-up history, drops the outputs it produces, and then equals the single-stream result."""
    x = cs.synth.config2(3 << 20, keyed=None)
    full = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
    ref = full.process(x)[0]
    shard = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
    start = 2 << 20
    warm = shard.warmup_len()
    assert warm < start
    pre = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
    n_before = len(pre.process(x[:start - warm])[0])
    n_warm = len(pre.process(x[start - warm:start])[0])
    shard.seek(start - warm)
    y = shard.process(x[start - warm:])[0]
    assert len(y) == len(ref) - n_before
    assert_parity(y[n_warm:], ref[n_before + n_warm:], rel=REL_TOL_AFTER_DCBLOCK, what="time-segment shard")


def test_channelizer_time_segment_sharding(cs, orc):
    """multi-GPU partitioning of the channelizer chains (time segments, every shard produces all channels
    of its frames): a shard that seeks to a frame boundary minus the warm-up equals the single-stream result -- with the
    resampler in front (README example 3 shape, 20 channels: the frame grid follows the resampler's output index) and
    for the bare 64-channel filterbank with --mix"""
    from composable_sdr_b200 import shard
    n = 1 << 22
    x = cs.synth.example3(n)
    ref = orc.Chain(3.2e6, 0.0, 1.6e6, orc.DEMOD_NO, 0.0, -50.0, 20, False).process(x)
    align = shard.frame_alignment(20, 1, 2)
    assert align == 40
    (a0, a1), (b0, b1) = shard.time_segments(n, 2, align)
    assert a0 == 0 and a1 == b0 and b1 == n and b0 % align == 0
    first = cs.Chain(3.2e6, 0.0, 1.6e6, agc=-50.0, channels=20).process(x[a0:a1])
    second = cs.Chain(3.2e6, 0.0, 1.6e6, agc=-50.0, channels=20)
    warm = shard.seek_shard(second, b0, lambda i, j: x[i:j])
    assert 0 < warm < b0
    tail = second.process(x[b0:b1])
    outs = [np.concatenate([u, v]) for u, v in zip(first, tail)]
    _assert_channels(outs, ref, "example 3, two time shards", fm=False, rel=1e-4)
    # a shard that does NOT start on a frame boundary still lands on the stream's frame grid: the frame its start falls
    # into is completed by its first samples (the warm-up supplied the frame's head)
    odd = b0 + 2 * 7                      # 7 resampler outputs into a frame
    third = cs.Chain(3.2e6, 0.0, 1.6e6, agc=-50.0, channels=20)
    shard.seek_shard(third, odd, lambda i, j: x[i:j])
    t3 = third.process(x[odd:b1])
    k0 = b0 // 40
    for c in (2, 9, 14):
        assert len(t3[c]) == len(ref[c]) - k0
        assert_parity(t3[c][64:], ref[c][k0 + 64:], what=f"unaligned shard start, channel {c}")
    # 64 channels, --mix, FM, no resampler
    x = cs.synth.config4(1 << 21, channels=64, active=8, sr=1e8)
    refm = orc.Chain(1e8, 0.0, 0.0, orc.DEMOD_NBFM, 0.3, -40.0, 64, True).process(x)[0]
    (a0, a1), (b0, b1) = shard.time_segments(len(x), 2, shard.frame_alignment(64))
    mk = lambda: cs.Chain(1e8, demod=cs.DeNBFM(0.3), agc=-40.0, channels=64, mix_channels=True)
    ya = mk().process(x[a0:a1])[0]
    sh = mk()
    shard.seek_shard(sh, b0, lambda i, j: x[i:j])
    yb = sh.process(x[b0:b1])[0]
    y = np.concatenate([ya, yb])
    assert len(y) == len(refm)
    assert snr_db(y[512:], refm[512:]) >= 60.0


def test_wbfm_time_segment_sharding(cs, orc):
    """DeWBFM shards by time segments too: the output decimator's block grid follows the absolute index of the demodulated
    samples, so a shard that starts anywhere continues the single-stream output (decim 4 and 5: 5 does not divide the
    resampler's output count at the shard start)"""
    from composable_sdr_b200 import shard
    x = cs.synth.config2(3 << 20, keyed=None)
    for decim in (4, 5):
        mk = lambda: cs.Chain(2.56e6, 1e5, 200e3, cs.DeWBFM(decim), agc=-40.0)
        ref = mk().process(x)[0]
        start = (2 << 20) + 12345
        pre = mk()
        n_before = len(pre.process(x[:start])[0])
        sh = mk()
        warm = shard.seek_shard(sh, start, lambda i, j: x[i:j])
        assert 0 < warm < start
        y = sh.process(x[start:])[0]
        assert len(y) == len(ref) - n_before
        assert_parity(y, ref[n_before:], rel=REL_TOL_AFTER_DCBLOCK, what=f"DeWBFM {decim} time-segment shard")


def test_firpfbch2_time_segment_sharding(cs, orc):
    """the channelizer the task names (firpfbch2_crcf, frames of C/2 samples) shards by time segments as well: the frame
    grid and the sign (-1)^(c t) of its per-channel factor follow the absolute position, for a shard that starts on an
    even frame, on an odd frame and inside a frame"""
    from composable_sdr_b200 import shard
    x = cs.synth.config4(1 << 21, channels=64, active=8, sr=1e8)
    ref = orc.Chain(1e8, 0.0, 0.0, orc.DEMOD_NO, 0.0, -40.0, 64, False, channelizer=1).process(x)
    mk = lambda: cs.Chain(1e8, agc=-40.0, channels=64, channelizer=1)
    hop = 32
    for start in ((1 << 20), (1 << 20) + hop, (1 << 20) + hop + 5):
        sh = mk()
        warm = shard.seek_shard(sh, start, lambda i, j: x[i:j])
        assert 0 < warm < start
        outs = sh.process(x[start:])
        k0 = start // hop                                    # whole frames in front of the shard
        for c in (0, 1, 31, 32, 63):
            assert len(outs[c]) == len(ref[c]) - k0
            assert_parity(outs[c][64:], ref[c][k0 + 64:], what=f"firpfbch2 shard at {start}, channel {c}")
        sh.close()


def test_full_size_properties_on_device(cs):
    """BASELINE-size chunk (2^26 samples, device resident): exact output count, linearity of the front end,
    chunk invariance, and few AGC speculation misses."""
    import torch
    n = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(1)
    k = torch.arange(n, device="cuda", dtype=torch.float64)
    carrier = 0.5 * torch.exp(1j * (2 * np.pi * 1e5 / 2.56e6 * k - 50.0 * torch.cos(2 * np.pi * 1e3 / 2.56e6 * k)))
    noise = 0.05 * torch.complex(torch.randn(n, generator=g, device="cuda"), torch.randn(n, generator=g, device="cuda"))
    x1 = carrier.to(torch.complex64)
    x2 = noise.to(torch.complex64)
    del carrier, noise, k
    def fe():
        return cs.Chain(2.56e6, 1e5, 200e3)
    y1 = fe().process(x1)[0]
    y2 = fe().process(x2)[0]
    y12 = fe().process(x1 + x2)[0]
    assert len(y1) == n * 200e3 / 2.56e6                   # 5 242 880 outputs exactly
    lin = (y12 - (y1 + y2)).abs().max().item()
    assert lin <= 1e-4 * y12.abs().max().item()
    c = fe()
    parts = torch.cat([c.process(x1[i:i + (1 << 24) + 8])[0] for i in range(0, n, (1 << 24) + 8)])
    assert parts.shape == y1.shape                          # (the dc blocker's fp64 carries depend on the chunking)
    assert (parts - y1).abs().max().item() <= 1e-4 * y1.abs().max().item()
    r1 = cs.resampler(0.078125, 60.0)                      # the front end alone is bit-exact under re-chunking
    h = r1._start()
    a = r1._process(h, x2)
    r1._done(h)
    h = r1._start()
    b = torch.cat([r1._process(h, x2[i:i + (1 << 24) + 8]) for i in range(0, n, (1 << 24) + 8)])
    r1._done(h)
    assert torch.equal(a, b)
    ch = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
    a = ch.process(x1 + x2)[0]
    assert len(a) == len(y1)
    assert torch.isfinite(a).all()
    assert ch.agc_fixups() < 64
    # the demodulated tone: 1 kHz at deviation 50 kHz -> amplitude (dev/fs_out)/kf
    seg = a[100000:200000].double()
    assert 0.75 < seg.abs().max().item() < 1.05            # (dev / fs_out) / kf = 0.833 plus noise


def test_wbfm_tail_in_the_chain(cs, orc):
    """DeWBFM decim: agc -> freqdem 0.6 -> de-emphasis (2nd-order Butterworth at 5 kHz of the quadrature
    rate) -> firdecim, single stream whole / ragged chunks, two streams, and per channel behind the channelizer"""
    x = cs.synth.config2(1 << 21)
    ref = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_WBFM, 0.6, -40.0, decim=4).process(x)[0]
    per = 1.0 / 0.6 * 0.01                      # a 2 pi slip of the discriminator after the de-emphasis: not periodic any more
    del per
    y = run_chain(cs.Chain(2.56e6, 1e5, 200e3, cs.DeWBFM(4), agc=-40.0), x, [1 << 21])[0]
    assert len(y) == len(ref) == (1 << 21) * 5 // 64 // 4
    assert_parity(y, ref, rel=REL_TOL_AFTER_DCBLOCK, what="DeWBFM 4")
    y = run_chain(cs.Chain(2.56e6, 1e5, 200e3, cs.DeWBFM(4), agc=-40.0), x, [300001, 1024, 77])[0]
    assert_parity(y, ref, rel=REL_TOL_AFTER_DCBLOCK, what="DeWBFM 4 (ragged chunks)")
    # decimation 5 does not divide the chunk's output: pending samples wait in the handle
    ref5 = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_WBFM, 0.6, -40.0, decim=5).process(x)[0]
    y = run_chain(cs.Chain(2.56e6, 1e5, 200e3, cs.DeWBFM(5), agc=-40.0), x, [123457])[0]
    assert_parity(y, ref5, rel=REL_TOL_AFTER_DCBLOCK, what="DeWBFM 5 (ragged chunks)")
    # two streams in one handle
    x2 = np.stack([x[:1 << 19], x[1 << 19:1 << 20]])
    outs = run_chain(cs.Chain(2.56e6, 1e5, 200e3, cs.DeWBFM(4), agc=-40.0, nstreams=2), x2, [1 << 18, 99999])
    for s in range(2):
        r = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_WBFM, 0.6, -40.0, decim=4).process(x2[s])[0]
        assert_parity(outs[s], r, rel=REL_TOL_AFTER_DCBLOCK, what=f"DeWBFM 4, stream {s}")
    # behind the channelizer: one tail per channel, and the --mix sum of the tails
    x3 = cs.synth.config3(1 << 19)
    r3 = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_WBFM, 0.6, -40.0, 16, False, decim=4).process(x3)
    o3 = run_chain(cs.Chain(2.56e6, demod=cs.DeWBFM(4), agc=-40.0, channels=16), x3, [1 << 18, 5000, 100000])
    assert len(o3) == 16
    for c in range(16):
        assert_parity(o3[c], r3[c], rel=REL_TOL_FM_NOISE, what=f"DeWBFM 4 behind -c 16, channel {c}")
    r3m = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_WBFM, 0.6, -40.0, 16, True, decim=4).process(x3)[0]
    o3m = run_chain(cs.Chain(2.56e6, demod=cs.DeWBFM(4), agc=-40.0, channels=16, mix_channels=True), x3, [1 << 19])[0]
    assert_parity(o3m, r3m, rel=REL_TOL_FM_NOISE, what="DeWBFM 4 behind -c 16 --mix")
