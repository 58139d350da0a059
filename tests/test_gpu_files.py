"""GPU tests of the file source / sinks either side of the path: csdr_chain_run_file against the
oracle's chain on the same CF32 file contents, with the names and the takeNArr count of apps/SoapySDR.hs:207-240."""
import os

import numpy as np
import pytest

from util import REL_TOL_AFTER_DCBLOCK, assert_parity

pytestmark = pytest.mark.gpu


def test_config1_file_to_file(cs, orc, tmp_path):
    """soapy-sdr --filename synth.cf32 -n 100000 -s 2.56e6 --offset 1e5 -b 200000 --demod DeNo --output out"""
    x = cs.synth.config1(1 << 21)
    src = tmp_path / "synth.cf32"
    x.tofile(src)                                              # interleaved little-endian float32 I/Q
    ch = cs.Chain(2.56e6, 1e5, 200e3)
    n_in, n_out = ch.run_file(src, tmp_path / "out", numsamples=100000, chunk=300000)
    assert n_out == 100000 and n_in % 300000 == 0 and n_in < x.size
    y = np.fromfile(tmp_path / "out.cf32", np.complex64)
    assert y.size == 100000
    ref = orc.Chain(2.56e6, 1e5, 200e3).process(x)[0]
    assert_parity(y, ref[:100000], what="config 1 file to file")
    # the whole file, default chunk; a trailing partial sample (4 stray bytes) is ignored
    with open(src, "ab") as f:
        f.write(b"\x00\x00\x80\x3f")
    n_in, n_out = cs.Chain(2.56e6, 1e5, 200e3).run_file(src, tmp_path / "all")
    assert n_in == x.size and n_out == len(ref)
    assert_parity(np.fromfile(tmp_path / "all.cf32", np.complex64), ref, what="config 1 whole file")


def test_channelizer_and_demod_file_names(cs, orc, tmp_path):
    """-c 4 writes <name>_ch1.cf32 .. _ch4.cf32 with numsamples / 4 samples each; a demodulator writes float32"""
    x = cs.synth.config3(1 << 18)
    src = tmp_path / "wide.cf32"
    x.tofile(src)
    n_in, n_out = cs.Chain(2.56e6, channels=4).run_file(src, tmp_path / "pfb", numsamples=40002, chunk=50001)
    assert n_out == 10000
    ref = orc.Chain(2.56e6, 0.0, 0.0, orc.DEMOD_NO, 0.3, 0.0, 4, False).process(x)
    for k in range(4):
        y = np.fromfile(tmp_path / f"pfb_ch{k + 1}.cf32", np.complex64)
        assert y.size == 10000
        assert_parity(y, ref[k][:10000], rel=REL_TOL_AFTER_DCBLOCK, what=f"-c 4 file, channel {k + 1}")
    assert not os.path.exists(tmp_path / "pfb.cf32")
    x2 = cs.synth.config2(1 << 20)
    x2.tofile(tmp_path / "fm.cf32")
    n_in, n_out = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0).run_file(tmp_path / "fm.cf32", tmp_path / "audio", chunk=1 << 18)
    r2 = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_NBFM, 0.3, -40.0).process(x2)[0]
    y2 = np.fromfile(tmp_path / "audio.f32", np.float32)
    assert n_out == y2.size == len(r2)
    assert_parity(y2, r2, rel=REL_TOL_AFTER_DCBLOCK, what="NBFM file to file")
    with pytest.raises(cs.CsdrError, match="cannot open"):
        cs.Chain(2.56e6).run_file(tmp_path / "missing.cf32", tmp_path / "x")
