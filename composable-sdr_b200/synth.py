"""Synthetic CF32 inputs for the BASELINE configs.  numpy Philox, seed 0x5D2B200; zero-mean by
construction.  Used by tests, smoke() and bench.py (the bench's device-resident input is generated the same way
on the GPU with torch, see bench.py)."""
import numpy as np

SEED = 0x5D2B200


def _rng(stream_id):
    return np.random.Generator(np.random.Philox(key=SEED + int(stream_id)))


def noise(n, sigma, stream_id=0):
    g = _rng(stream_id)
    return (sigma * (g.standard_normal(n, dtype=np.float32) + 1j * g.standard_normal(n, dtype=np.float32))).astype(np.complex64)


def fm_carrier(n, sr, f0, amp=0.5, dev=50e3, f_audio=1e3, keyed=None, k0=0):
    """FM-modulated tone at f0 Hz; keyed=(on_s, period_s) switches the carrier on/off (exercises the squelch)."""
    k = np.arange(k0, k0 + n, dtype=np.float64)
    phase = 2 * np.pi * f0 * k / sr - (dev / f_audio) * np.cos(2 * np.pi * f_audio * k / sr)
    x = amp * np.exp(1j * phase)
    if keyed is not None:
        on, period = keyed
        x = x * ((k / sr) % period < on)
    return x.astype(np.complex64)


def am_carrier(n, sr, f0, amp=0.5, index=0.8, f_audio=1e3, k0=0):
    k = np.arange(k0, k0 + n, dtype=np.float64)
    env = 1.0 + index * np.sin(2 * np.pi * f_audio * k / sr)
    return (amp / (1 + index) * env * np.exp(2j * np.pi * f0 * k / sr)).astype(np.complex64)


def tone(n, sr, f0, amp, k0=0):
    k = np.arange(k0, k0 + n, dtype=np.float64)
    return (amp * np.exp(2j * np.pi * f0 * k / sr)).astype(np.complex64)


def config1(n, k0=0):
    """C1/C2 input: SR 2.56 MS/s, FM carrier at +100 kHz, interferer at -400 kHz, noise."""
    return _cfg12(n, 2.56e6, None, k0)


def _cfg12(n, sr, keyed, k0):
    g = np.random.Generator(np.random.Philox(key=[SEED + 12, int(k0)]))
    nz = 0.05 * (g.standard_normal(n, dtype=np.float32) + 1j * g.standard_normal(n, dtype=np.float32))
    x = nz + fm_carrier(n, sr, 1e5, 0.5, 50e3, 1e3, keyed, k0) + tone(n, sr, -4e5, 0.3, k0)
    return x.astype(np.complex64)


def config2(n, k0=0, keyed=(0.05, 0.2)):
    """C2: as C1 with the carrier keyed 50 ms on / 200 ms period."""
    return _cfg12(n, 2.56e6, keyed, k0)


def config3(n, channels=16, sr=2.56e6):
    """C3: NBFM carriers at the channel centres f_k = (k - (C-1)/2)/C * SR, every other one keyed."""
    x = noise(n, 0.01, 3).astype(np.complex64)
    for k in range(channels):
        f0 = (k - (channels - 1) / 2.0) / channels * sr
        keyed = (0.002, 0.008) if (k % 2) else None
        x = x + fm_carrier(n, sr, f0, 0.05 + 0.01 * k, 0.02 * sr / channels, sr / channels / 40.0, keyed)
    return x.astype(np.complex64)


def config4(n, channels=1024, active=64, sr=1e9):
    """C4: `active` of `channels` channels carry NBFM with random amplitudes (64 of 1024 active)."""
    g = _rng(4)
    # channel outputs of the un-normalised M-point analyzer carry a tone of amplitude A as A*M and white noise of
    # std s as ~s*sqrt(M): noise-only channels must stay below the -40 dB squelch threshold (they are gated), active
    # ones well above it
    x = noise(n, 3e-5, 40)
    idx = g.choice(channels, size=active, replace=False)
    # (the 1024-level NCO of the pre-rotation leaves spurs ~45 dB below a carrier in other channels: carriers
    # are kept below 0 dB at the analyzer output so that their spurs stay clear of the -40 dB squelch threshold)
    amps = g.uniform(1e-4, 7e-4, size=active)
    for k, a in zip(idx, amps):
        f0 = (k - (channels - 1) / 2.0) / channels * sr
        x = x + fm_carrier(n, sr, f0, a, 0.1 * sr / channels, sr / channels / 50.0)
    return x.astype(np.complex64)


def config5(n, nstreams, sr=10e6):
    """C5: per stream one AM carrier (index 0.8, 1 kHz tone) near +1 MHz plus noise.  The carrier is 437 Hz off the
    mixer frequency: a carrier exactly at +1 MHz lands on DC and is removed by the chain's dc blocker, which leaves
    the carrier-tracking AM demodulator nothing to lock to."""
    out = np.empty((nstreams, n), np.complex64)
    for s in range(nstreams):
        out[s] = noise(n, 0.02, 500 + s) + am_carrier(n, sr, 1e6 + 437.0 + 3.0 * s, 0.4 + 0.002 * s, 0.8, 1e3 + 10 * s)
    return out


def _example3(n, sr, offset, bw, channels, active, noise_sigma, stream):
    x = noise(n, noise_sigma, stream)
    k = np.arange(n, dtype=np.float64)
    for i, c in enumerate(active):
        f0 = offset + (c - (channels - 1) / 2.0) / channels * bw
        car = fm_carrier(n, sr, f0, 0.01 + 0.004 * i, 0.1 * bw / channels, bw / channels / 40.0)
        if i % 3 == 1:
            # slow raised-cosine fading (period 12 ms) through the squelch threshold: the gate opens and closes without
            # key clicks (an abrupt on/off splatters into every channel and drives liquid's gain loop into its frozen
            # state, where one-ulp differences persist for ever)
            car = car * (0.5 - 0.5 * np.cos(2 * np.pi * k / (0.012 * sr))).astype(np.float32) ** 4
        x = x + car
    return x.astype(np.complex64)


def example3(n, sr=3.2e6, bw=1.6e6, channels=20, active=(2, 5, 9, 10, 14, 18), noise_sigma=1e-5):
    """README example 3 (README.md:182-193: -s 3.2e6 -b 1.6e6 -a -50 -c 20 --demod DeNo): carriers at some channel
    centres of the RESAMPLED band, f_k = (k - (C-1)/2)/C * bw, in front of the resampler; the rest of the band holds
    noise far below the squelch threshold (the analyzer shows a tone of amplitude A as ~A*C)."""
    return _example3(n, sr, 0.0, bw, channels, active, noise_sigma, 30)


def example3_offset(n, sr=2.56e6, offset=1e5, bw=1.28e6, channels=16, active=(1, 4, 7, 8, 12, 15), noise_sigma=1e-5):
    """as example3 with an offset mix in front: the carriers sit at offset + the channel centres of the resampled band"""
    return _example3(n, sr, offset, bw, channels, active, noise_sigma, 31)
