"""Shared helpers for the parity tests.

Tolerance (stated once, used everywhere): the north star asks for "max relative error <= 1e-4 or output SNR vs
reference >= 80 dB".  All arithmetic is float32; `assert_parity` checks BOTH
    max |y - ref| <= rel * max |ref|          (peak-relative error, rel = 1e-4 unless a test says otherwise)
    SNR = 10 log10( sum |ref|^2 / sum |y - ref|^2 ) >= 80 dB
"""
import numpy as np

REL_TOL = 1e-4
SNR_DB = 80.0
# Discriminator / AM outputs taken AFTER the chain's dc blocker: the wanted carrier sits at (or next to) DC after the
# offset mix, so the blocker's state |v| ~ |x| / alpha is ~10^3 times the signal and float32 rounding of the
# recurrence (~6e-8 |v|) is ~1e-4 of the signal; where the squelch is open on noise only (~30 dB lower) the
# discriminator turns that into ~5e-3 rad.  Any two float32 evaluation orders of liquid's recurrence differ by this
# much.  Such outputs are held to SNR >= 80 dB and a peak-relative error of 5e-3.
REL_TOL_AFTER_DCBLOCK = 5e-3
# FM discriminator on a channel that carries noise or a weak signal: arg(conj(r[n-1]) r[n]) is ill-conditioned wherever
# |r[n-1] r[n]| is small (the float32 rounding of the channelizer, ~1e-7 of the channel's level, is divided by that
# product), so single samples may differ by ~1e-3 of the +-pi range while the SNR stays above 80 dB
REL_TOL_FM_NOISE = 5e-3


def snr_db(y, ref):
    y = np.asarray(y)
    ref = np.asarray(ref)
    num = float(np.sum(np.abs(ref.astype(np.complex128)) ** 2))
    den = float(np.sum(np.abs(y.astype(np.complex128) - ref.astype(np.complex128)) ** 2))
    if den == 0.0:
        return np.inf
    if num == 0.0:
        return -np.inf
    return 10.0 * np.log10(num / den)


def assert_parity(y, ref, rel=REL_TOL, snr=SNR_DB, what="", period=None, allow_empty=False):
    """period: for discriminator outputs, 1/kf (= 2 pi in units of the output).  arg() has a branch cut at +-pi: on a
    noise-only sample pair that lands next to it, rounding decides between -pi+e and +pi-e, so errors are compared
    modulo the period."""
    y = np.asarray(y)
    ref = np.asarray(ref)
    assert y.shape == ref.shape, f"{what}: shape {y.shape} != {ref.shape}"
    if ref.size == 0:
        # an empty comparison proves nothing: a test that slices everything away must say so explicitly
        assert allow_empty, f"{what}: nothing to compare (empty arrays)"
        return
    if period is not None:
        d = y.astype(np.float64) - ref.astype(np.float64)
        y = (ref.astype(np.float64) + (d - period * np.round(d / period)))
    peak = float(np.max(np.abs(ref)))
    err = float(np.max(np.abs(y - ref)))
    s = snr_db(y, ref)
    assert err <= rel * max(peak, 1e-30), f"{what}: max err {err:.3e} > {rel:g} * peak {peak:.3e} (snr {s:.1f} dB)"
    assert s >= snr, f"{what}: SNR {s:.1f} dB < {snr} dB"


def chunked(x, sizes):
    pos = 0
    i = 0
    while pos < len(x):
        n = sizes[i % len(sizes)]
        yield x[pos:pos + n]
        pos += n
        i += 1


def make_signal(n, seed=0, amp=0.5):
    """zero-mean complex test input: noise + two tones"""
    g = np.random.default_rng(seed)
    k = np.arange(n)
    x = 0.05 * (g.standard_normal(n) + 1j * g.standard_normal(n))
    x = x + amp * np.exp(2j * np.pi * 0.0391 * k + 1j * 3.0 * np.sin(2 * np.pi * k / 2560.0))
    x = x + 0.3 * np.exp(-2j * np.pi * 0.156 * k)
    return x.astype(np.complex64)


def away_from_gate_edges(ref, guard=1):
    """mask of samples whose squelch gate and whose neighbours' gates are all open or all closed.  At an edge the
    discriminator sees one zeroed sample and returns arg(+-0 +- j0) = 0 or +-pi: the sign comes from the sign of a
    possibly tiny ungated component, which float32 rounding can flip; those samples are compared separately."""
    closed = (np.asarray(ref) == 0)
    edge = np.zeros(closed.shape, bool)
    d = closed[1:] != closed[:-1]
    for k in range(-guard, guard + 1):
        idx = np.nonzero(d)[0] + 1 + k
        idx = idx[(idx >= 0) & (idx < closed.size)]
        edge[idx] = True
    # isolated +-pi artefacts also appear one sample after the gate opens
    return ~edge
