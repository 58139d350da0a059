"""Product-side parameter derivation (composable-sdr_b200/csrc/design.hpp, compiled into the test-only emulation
library) against the oracle's independent restatement.  CPU only."""
import numpy as np
import pytest


@pytest.mark.parametrize("rate,As", [(0.078125, 60.0), (0.02, 60.0), (0.5, 60.0), (0.3, 60.0), (0.078125, 40.0),
                                     (0.011, 80.0)])
def test_msresamp_design_matches_oracle(orc, emu, rate, As):
    a = emu.design_msresamp(rate, As)
    b = orc.MsResamp(rate, As).design()
    assert a["S"] == b["S"] and a["m"] == b["m"]
    assert a["rate_arb"] == b["rate_arb"] and a["step"] == b["step"] and a["npfb"] == b["npfb"]
    for s in range(a["S"]):
        assert np.max(np.abs(a["h1"][s] - b["h1"][s])) < 1e-7
    assert np.max(np.abs(a["bank"] - b["bank"])) < 2e-7


@pytest.mark.parametrize("M", [16, 20, 1024])
def test_firpfbch_design_matches_oracle(orc, emu, M):
    assert np.max(np.abs(emu.design_firpfbch(M) - orc.Firpfbch(M).taps())) < 1e-8


def test_nco_words(orc, emu):
    for f in [0.24543693, -2.984513, 1e-3, 3.0, 6.2, -0.5]:
        assert emu.L.emu_design_nco_constrain(f) == orc.Nco(f).freq_word
    for C in [16, 20, 1024]:
        assert emu.L.emu_design_nco_constrain(emu.L.emu_design_rotation(C)) == orc.Firpfbch(C).nco.freq_word


def test_agc_threshold_is_exact_boundary(orc, emu):
    """rssi(g) > T  <=>  g < g_thr for the float32 values around the boundary (the kernels compare gains)."""
    for T in [-40.0, -50.0, -12.5]:
        g_thr = np.float32(emu.L.emu_design_agc_threshold(T))
        below = np.nextafter(g_thr, np.float32(0))
        rssi = lambda g: np.float32(-20 * np.log10(np.float64(g)))
        assert not (rssi(g_thr) > np.float32(T))
        assert rssi(below) > np.float32(T)


def test_frontend_geometry(emu):
    S, n, d, hcap, smem = emu.geometry(0.078125, 60.0, 464)       # standard plan -> compile-time geometry
    assert S == 3
    assert d == [0, -39, -97, -206]                       # lo_L = (c_lo << L) + d[L]; top level shifted by one
    assert n[0] == 416 and all(v % 8 == 0 for v in n)
    assert hcap >= 7 + (16 << 3) + 206
    assert smem <= 113 * 1024                             # 2 CTAs / SM (levels + raw staging buffer + bank)
