"""GPU edge cases through the C ABI: empty and sub-frame chunks, error reporting, digital silence (the AGC gain freezes:
speculation must be repaired), non-finite input."""
import numpy as np
import pytest

from util import REL_TOL_AFTER_DCBLOCK, assert_parity, make_signal
from test_gpu_chain import run_chain

pytestmark = pytest.mark.gpu


def test_channelizer_chain_with_chunks_smaller_than_a_frame(cs, orc):
    """left-over path: chunks of 1..40 samples into a 16-channel chain (a frame is 16 samples)"""
    x = cs.synth.config3(1 << 14)
    ref = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NO, 0.0, 0.0, 16, False).process(x)
    outs = run_chain(cs.Chain(2.56e6, channels=16), x, [5, 3, 9, 1, 40, 1000, 7, 16, 15, 4096])
    for c in range(16):
        assert len(outs[c]) == len(ref[c])
        assert_parity(outs[c], ref[c], rel=REL_TOL_AFTER_DCBLOCK, what=f"sub-frame chunks, channel {c}")


def test_empty_and_single_sample_calls(cs, orc):
    x = cs.synth.config2(200000)
    ref = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_NBFM, 0.3, -40.0).process(x)[0]
    ch = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
    parts = []
    for i, j in [(0, 0), (0, 1), (1, 1), (1, 2), (2, 100000), (100000, 100000), (100000, 200000)]:
        o = ch.process(np.ascontiguousarray(x[i:j]))
        parts.append(o[0])
        if i == j:
            assert len(o[0]) == 0
    y = np.concatenate(parts)
    assert len(y) == len(ref)
    assert_parity(y, ref, rel=REL_TOL_AFTER_DCBLOCK, period=1 / 0.3, what="empty / 1-sample calls")


def test_errors_are_reported_not_fatal(cs):
    with pytest.raises(cs.CsdrError, match="samplerate"):
        cs.Chain(0.0, 0.0, 200e3)                                  # liquid would print and exit(1)
    with pytest.raises(cs.CsdrError, match="kf"):
        cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.0), agc=-40.0)
    # the library is still usable afterwards
    ch = cs.Chain(2.56e6, 1e5, 200e3)
    assert len(ch.process(make_signal(4096, 1))[0]) > 0


def test_digital_silence_freezes_the_gain_and_is_repaired(cs, orc):
    """exact zeros: y2' decays below 1e-6 and liquid stops updating the gain, so no warm-up can re-derive it; the
    verify / refine / in-order repair passes must restore the sequential result (gate positions exactly)"""
    n = 1 << 18
    t = np.arange(n)
    g = np.random.default_rng(5)
    x = (0.2 * np.exp(2j * np.pi * 0.01 * t) + 0.002 * (g.standard_normal(n) + 1j * g.standard_normal(n))).astype(np.complex64)
    x[60000:150000] = 0
    ref = orc.Agc(-40.0).execute(x)
    agc = cs.automaticGainControl(-40.0)
    r = agc._start()
    y = np.concatenate([agc._process(r, x[:100000]), agc._process(r, x[100000:])])
    agc._done(r)
    assert np.array_equal(y == 0, ref == 0)
    assert_parity(y, ref, what="agc over digital silence")


def test_non_finite_input_does_not_hang(cs):
    x = make_signal(100000, 3)
    x[5000] = np.nan
    x[70000] = np.inf
    ch = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
    y = ch.process(x)[0]
    assert len(y) > 0
    ch2 = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)          # a fresh handle is unaffected
    assert np.all(np.isfinite(ch2.process(make_signal(100000, 3))[0]))
