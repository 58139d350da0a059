"""Committed fixtures (tests/golden, made by tests/golden/make_golden.py): the reference's known answers and the
oracle's regression vectors.  CPU part: the oracle still reproduces them.  GPU part: the CUDA path reproduces them
without consulting the oracle at run time."""
import os

import numpy as np
import pytest

from util import REL_TOL_AFTER_DCBLOCK, assert_parity, chunked

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(GOLD, "oracle_vectors.npz"))


@pytest.fixture(scope="module")
def ka():
    return np.load(os.path.join(GOLD, "known_answers.npz"))


def test_known_answers_fixture(orc, ka):
    assert np.max(np.abs(orc.Firpfbch(20).taps()[249:280] - ka["firpfbch20_taps_249_279"])) < 1.5e-8
    assert orc.Firpfbch(20).nco.freq_word == int(ka["nco_rotation_freq_word_c20"][0])
    assert abs(orc.DcBlocker(0.001).coeffs()[1][1] - float(ka["dc_blocker_a1_alpha_1e3"][0])) < 1e-8


def test_oracle_reproduces_its_vectors(orc, vec):
    x = vec["x"]
    f = float(np.float32(0.24543693))
    assert np.array_equal(orc.Nco(f).mix_down(x), vec["nco_down"])
    assert np.array_equal(orc.MsResamp(0.078125).execute(x), vec["msresamp_0p078125"])
    assert np.array_equal(orc.MsResamp(0.3).execute(x), vec["msresamp_0p3"])
    assert np.array_equal(orc.DcBlocker().execute(x), vec["dcblock"])
    assert np.array_equal(orc.Firpfbch(16).execute(x[:16 * 300]), vec["firpfbch16"])
    assert np.array_equal(orc.Agc(-40.0).execute(x * np.float32(0.05)), vec["agc_m40"])
    assert np.array_equal(orc.FreqDem(0.3).execute(x), vec["freqdem_0p3"])
    y = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_NBFM, 0.3, -40.0).process(vec["c2_x"])[0]
    assert np.array_equal(y, vec["c2_y"])
    xr = np.ascontiguousarray(x.real)
    assert np.array_equal(orc.IirFiltRRRF(2, 0.025).execute(xr), vec["iirfilt_butter2_0p025"])
    assert np.array_equal(orc.FirDecim(4).execute(xr), vec["firdecim4"])
    y = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_WBFM, 0.6, -40.0, decim=4).process(vec["c2_x"])[0]
    assert np.array_equal(y, vec["c2_wbfm4_y"])


@pytest.mark.gpu
def test_gpu_blocks_against_golden(cs, vec):
    x = vec["x"]
    f = float(np.float32(0.24543693))

    def run(pipe, sizes=(1024, 333)):
        process, cleanup = cs.unPipe(pipe)
        out = list(process(chunked(x, list(sizes))))
        cleanup()
        return out
    assert_parity(np.concatenate(run(cs.mixDown(f))), vec["nco_down"], what="nco")
    assert_parity(np.concatenate(run(cs.resampler(0.078125, 60.0))), vec["msresamp_0p078125"], what="msresamp")
    assert_parity(np.concatenate(run(cs.resampler(0.3, 60.0))), vec["msresamp_0p3"], what="msresamp 0.3")
    assert_parity(np.concatenate(run(cs.dcBlocker())), vec["dcblock"], what="dcblock")
    assert_parity(np.concatenate(run(cs.fmDemodulator(0.3))), vec["freqdem_0p3"], rel=2e-4, what="freqdem")
    xr = np.ascontiguousarray(x.real)

    def run_real(pipe, sizes):
        process, cleanup = cs.unPipe(pipe)
        out = list(process(chunked(xr, list(sizes))))
        cleanup()
        return np.concatenate(out)
    assert_parity(run_real(cs.iirFilter(2, 0.025, 0.0, 10.0, 10.0), (1024, 333)), vec["iirfilt_butter2_0p025"], what="iirfilt_rrrf")
    assert_parity(run_real(cs.firDecimator(4), (1024, 332)), vec["firdecim4"], what="firdecim_rrrf")


@pytest.mark.gpu
def test_gpu_chain_against_golden(cs, vec):
    y = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0).process(vec["c2_x"])[0]
    assert np.array_equal(y == 0, vec["c2_y"] == 0)
    assert_parity(y, vec["c2_y"], rel=REL_TOL_AFTER_DCBLOCK, what="config 2 golden")
    w = cs.Chain(2.56e6, 1e5, 200e3, cs.DeWBFM(4), agc=-40.0).process(vec["c2_x"])[0]
    assert_parity(w, vec["c2_wbfm4_y"], rel=REL_TOL_AFTER_DCBLOCK, what="config 2 with DeWBFM 4, golden")
    outs = cs.Chain(2.56e6, demod=cs.DeNBFM(0.3), agc=-40.0, channels=16).process(vec["c3_x"])
    for c in range(16):
        assert_parity(outs[c], vec["c3_y"][c], rel=3e-4, what=f"config 3 golden channel {c}")
