"""Generates tests/golden/*.npz.

The reference (Haskell on an un-vendored liquid-dsp) cannot be run in this environment, so these are NOT reference
outputs: they are (a) the known answers recovered from the reference's own screen capture (known_answers.npz, values
typed from images/ex1_5.gif) and (b) regression vectors of the CPU oracle on small seeded
inputs (oracle_vectors.npz), which pin the oracle's behaviour so that a change to it is noticed, and give the GPU
tests a fixture that does not depend on the oracle being rebuilt identically.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from util import make_signal  # noqa: E402
import composable_sdr_b200.synth as synth  # noqa: E402


def main():
    ka = dict(
        firpfbch20_taps_249_279=np.array([
            -0.00394838, -0.00372026, -0.00341510, -0.00305220, -0.00265029, -0.00222706, -0.00179865, -0.00137929,
            -0.00098111, -0.00061390, -0.00028515, 0.00000000, 0.00023857, 0.00042962, 0.00057400, 0.00067408,
            0.00073350, 0.00075681, 0.00074922, 0.00071631, 0.00066374, 0.00059706, 0.00052148, 0.00044173,
            0.00036191, 0.00028547, 0.00021510, 0.00015277, 0.00009977, 0.00005670, 0.00002359]),
        nco_rotation_freq_word_c20=np.array([0x86666600], np.uint64),
        dc_blocker_a1_alpha_1e3=np.array([-0.99900001]),
        per_channel_samples_n16000000_c20=np.array([800000], np.int64),
    )
    np.savez(os.path.join(HERE, "known_answers.npz"), **ka)

    v = {}
    x = make_signal(6000, 101)
    v["x"] = x
    f = float(np.float32(0.24543693))
    v["nco_down"] = O.Nco(f).mix_down(x)
    v["msresamp_0p078125"] = O.MsResamp(0.078125).execute(x)
    v["msresamp_0p3"] = O.MsResamp(0.3).execute(x)
    v["dcblock"] = O.DcBlocker().execute(x)
    v["firpfbch16"] = O.Firpfbch(16).execute(x[:16 * 300])
    v["agc_m40"] = O.Agc(-40.0).execute(x * np.float32(0.05))
    v["freqdem_0p3"] = O.FreqDem(0.3).execute(x)
    xs = synth.config2(1 << 16)
    v["c2_x"] = xs
    v["c2_y"] = O.Chain(2.56e6, 1e5, 200e3, O.DEMOD_NBFM, 0.3, -40.0).process(xs)[0]
    x3 = synth.config3(1 << 14)
    v["c3_x"] = x3
    v["c3_y"] = np.stack(O.Chain(2.56e6, 0.0, 0.0, O.DEMOD_NBFM, 0.3, -40.0, 16, False).process(x3))
    # wide-band FM tail
    xr = np.ascontiguousarray(x.real)
    v["iirfilt_butter2_0p025"] = O.IirFiltRRRF(2, 0.025).execute(xr)
    v["firdecim4"] = O.FirDecim(4).execute(xr)
    v["c2_wbfm4_y"] = O.Chain(2.56e6, 1e5, 200e3, O.DEMOD_WBFM, 0.6, -40.0, decim=4).process(xs)[0]
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **v)
    print({k: a.shape for k, a in v.items()})


if __name__ == "__main__":
    main()
