"""The real kernel sources (composable-sdr_b200/csrc/*.cuh) run under the TEST-ONLY CPU thread emulator and are
compared with the oracle: tile geometry, halo/history carry, resampler timing, dc scan and AGC speculation.
(The GPU parity tests proper are tests/test_gpu_*.py.)  CPU only."""
import numpy as np
import pytest

from util import assert_parity, make_signal


@pytest.mark.parametrize("rate,As,Tc,nthreads,mix,std", [
    (0.078125, 60.0, 464, 256, 1, False), (0.078125, 60.0, 64, 64, 2, False), (0.02, 60.0, 128, 128, 1, False),
    (0.625, 60.0, 256, 128, 0, False), (0.078125, 40.0, 64, 64, 1, True), (0.078125, 80.0, 64, 64, 0, True),
    # compile-time-geometry kernel k_frontend_std<S>, S = 1..6
    (0.3, 60.0, 0, 256, 1, True), (0.2, 60.0, 0, 256, 2, True), (0.078125, 60.0, 0, 256, 1, True),
    (0.04, 60.0, 0, 256, 0, True), (0.02, 60.0, 0, 256, 1, True), (0.011, 60.0, 0, 256, 1, True),
    # k_frontend_direct<S>, S = 1..6
    (0.3, 60.0, 0, 256, 1, 2), (0.2, 60.0, 0, 256, 2, 2), (0.078125, 60.0, 0, 256, 1, 2), (0.078125, 60.0, 0, 256, 0, 2),
    (0.04, 60.0, 0, 256, 2, 2), (0.02, 60.0, 0, 256, 1, 2), (0.011, 60.0, 0, 256, 1, 2),
    # k_frontend_ws<S> (warp-specialised), S = 2..6
    (0.2, 60.0, 0, 256, 2, 3), (0.078125, 60.0, 0, 256, 1, 3), (0.078125, 60.0, 0, 256, 0, 3), (0.04, 60.0, 0, 256, 2, 3),
    (0.02, 60.0, 0, 256, 1, 3), (0.011, 60.0, 0, 256, 1, 3)])
def test_frontend_matches_oracle(orc, emu, rate, As, Tc, nthreads, mix, std):
    x = make_signal(40000, 7)
    f = float(np.float32(0.24543693))
    xm = {0: lambda v: v, 1: orc.Nco(f).mix_down, 2: orc.Nco(f).mix_up}[mix](x)
    ref = orc.MsResamp(rate, As).execute(xm)
    y = emu.frontend(x, rate, As=As, mix_mode=mix, freq=f, Tc=Tc, nthreads=nthreads, std=std)
    assert_parity(y, ref, what="frontend")
    if std:
        # 8-byte aligned chunk: the scalar loader path must give the same bits as the float4 path
        assert np.array_equal(emu.frontend(x, rate, As=As, mix_mode=mix, freq=f, Tc=Tc, nthreads=nthreads, std=std, misalign=1), y)


@pytest.mark.parametrize("rate,mix", [(1.3, 0), (2.0, 1), (3.7, 2), (9.1, 0), (1.0000001, 0)])
def test_frontend_interpolation_matches_oracle(orc, emu, rate, mix):
    """msresamp_crcf with rate > 1: arbitrary stage first, then half-band interpolators (interp.cuh)"""
    x = make_signal(6000, 17)
    f = float(np.float32(0.24543693))
    xm = {0: lambda v: v, 1: orc.Nco(f).mix_down, 2: orc.Nco(f).mix_up}[mix](x)
    ref = orc.MsResamp(rate).execute(xm)
    y = emu.frontend(x, rate, mix_mode=mix, freq=f)
    assert_parity(y, ref, what=f"interpolating msresamp {rate}")
    sizes = [1, 7, 1000, 3, 2048]
    sizes.append(len(x) - sum(sizes))
    assert np.array_equal(emu.frontend(x, rate, mix_mode=mix, freq=f, chunks=sizes), y)


def test_frontend_chunk_invariance_is_bit_exact(emu):
    x = make_signal(20000, 8)
    sizes = [1, 7, 1000, 3, 4096, 5000]
    sizes.append(len(x) - sum(sizes))
    for std in (0, 1, 2, 3):
        a = emu.frontend(x, 0.078125, freq=0.3, std=std)
        b = emu.frontend(x, 0.078125, freq=0.3, chunks=sizes, std=std)
        assert np.array_equal(a, b)


def test_frontend_seek_with_warmup(orc, emu):
    """time-segment sharding: a shard that starts at sample n0 is seeded with seek(n0 - warm) and fed `warm`
    samples of real history; after the FIR history is filled the outputs equal the single-stream result."""
    x = make_signal(40000, 9)
    f = float(np.float32(0.7))
    rate = 0.078125
    ref = orc.MsResamp(rate).execute(orc.Nco(f).mix_down(x))
    start, warm = 20001, 1024
    y = emu.frontend(x[start - warm:], rate, freq=f, seek=start - warm)
    n_before = len(orc.MsResamp(rate).execute(x[:start - warm]))
    tail = ref[n_before:]
    assert len(y) == len(tail)
    assert_parity(y[100:], tail[100:], what="seek")


@pytest.mark.parametrize("kw", [dict(), dict(demod=0), dict(has_dc=0), dict(has_agc=0, demod=0), dict(L=128, W=384),
                                dict(L=64, W=384, G=64, demod=0), dict(L=256, W=512, has_dc=0), dict(has_agc=0)])
def test_backend_matches_oracle(orc, emu, kw):
    n = 30000
    t = np.arange(n)
    g = np.random.default_rng(3)
    ph = 2 * np.pi * (0.11 * t + 0.02 * np.cumsum(np.sin(2 * np.pi * t / 400)))
    env = ((t // 5000) % 2 == 0)
    x = (0.3 * env * np.exp(1j * ph) + 0.001 * (g.standard_normal(n) + 1j * g.standard_normal(n))).astype(np.complex64)
    y = x
    if kw.get("has_dc", 1):
        y = orc.DcBlocker().execute(y)
    if kw.get("has_agc", 1):
        y = orc.Agc(-40.0).execute(y)
    if kw.get("demod", 1):
        y = orc.FreqDem(0.3).execute(y)
    out, fixups = emu.backend(x, **kw)
    if kw.get("has_agc", 1):
        assert np.array_equal(out[0] == 0, y == 0)          # squelch gate aligned sample for sample
    assert_parity(out[0], y, what=f"backend {kw}")
    out2, _ = emu.backend(x, chunks=[100, 5000, 1, 12000, n - 17101], **kw)
    assert_parity(out2[0], y, what=f"backend chunked {kw}")


def test_backend_speculation_misses_are_repaired(orc, emu):
    """a stream that sits in exact digital silence freezes the AGC gain, which no warm-up can guess: the
    verify/fix-up pass must re-run those segments and still match the sequential loop.  When the weak signal resumes the
    gain sits at its 1e6 clamp and overshoots: for a few dozen samples the products conj(y[n-1]) y[n] are SUBNORMAL, and
    the discriminator's quotient has to stay exact there (be_atan2 lifts both operands before the fast reciprocal).  Only
    this harness reaches that case: every chain of the product has a dc blocker in front, which leaves no exact zeros."""
    n = 12000
    x = np.zeros(n, np.complex64)
    x[:3000] = 0.2 * np.exp(2j * np.pi * 0.07 * np.arange(3000))
    x[9000:] = 0.01 * np.exp(2j * np.pi * 0.03 * np.arange(3000))
    ref = orc.FreqDem(0.3).execute(orc.Agc(-40.0).execute(x))
    out, fixups = emu.backend(x, has_dc=0)
    assert_parity(out[0], ref, what="frozen gain")


@pytest.mark.parametrize("M,kind", [(16, 0), (20, 0), (4, 1), (16, 1), (32, 1), (8, 4), (16, 4), (6, 5), (12, 5), (20, 5), (128, 2), (256, 2), (512, 2), (128, 3), (256, 3), (512, 3)])
def test_channelizer_kernels_match_oracle(orc, emu, M, kind):
    """k_pfb (any M), k_pfb_tile (register DFT, M <= 32), k_pfb_tile2 (two frames per thread, M = 8, 16: kind 4), k_pfb_tile_any (even M, not a power of two: kind 5) and k_pfb_ring (M = 128..1024) against firpfbchChannelizer"""
    nf = 71 if M >= 128 else 300
    x = make_signal(M * nf, 21)
    ref = orc.Firpfbch(M).execute(x)
    y = emu.pfb(x, M, kind)
    assert_parity(y, ref, what=f"channelizer M={M} kind={kind}")
    chunks = [1, 13, nf - 40, 26]
    assert_parity(emu.pfb(x, M, kind, chunks), ref, what=f"channelizer M={M} kind={kind}, chunked")


@pytest.mark.parametrize("order,fc,M", [(2, 0.025, 4), (2, 0.005, 5), (5, 0.1, 1), (2, 0.2, 16)])
def test_wbfm_tail_kernels_match_oracle(orc, emu, order, fc, M):
    """k_iir2_seg / k_iir2_carry / k_iir2_apply (affine scans of the de-emphasis sections) + k_firdecim against the
    oracle's sequential iirfilt_rrrf -> firdecim_rrrf, whole and in ragged chunks, two lanes"""
    n = 6000 + M * 7 + 3
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((2, n)) + np.array([[0.3], [-0.2]])).astype(np.float32)
    ref = []
    for lane in range(2):
        f, d = orc.IirFiltRRRF(order, fc), orc.FirDecim(M)
        ref.append(d.execute(f.execute(x[lane])))
        bo, ao = f.coeffs()
    y, b, a = emu.wbfm_tail(x, order, fc, M)
    # (float32 design: 1 - pole cancels for narrow cut-offs, the two complex divisions round differently)
    assert np.abs(b - bo).max() <= 2e-5 * np.abs(bo).max() and np.abs(a - ao).max() <= 2e-6
    assert y.shape[1] == n // M
    for lane in range(2):
        assert_parity(y[lane], ref[lane], what=f"wbfm tail order={order} fc={fc} M={M}")
    yc, _, _ = emu.wbfm_tail(x, order, fc, M, chunks=[1, 255, 257, 1000, 3, n - 1516])
    for lane in range(2):
        assert_parity(yc[lane], ref[lane], what=f"wbfm tail order={order} fc={fc} M={M}, chunked")


@pytest.mark.parametrize("M", [2, 16, 20, 64])
def test_firpfbch2_kernel_matches_oracle(orc, emu, M):
    """k_pfb with a hop of M/2 and the oversampled analyzer's per-channel factor against the oracle's sequential
    firpfbch2 object (alternating window halves, backward DFT), whole and in chunks of odd frame counts"""
    nf = 201
    x = make_signal(M // 2 * nf, 23)
    o = orc.Firpfbch2(M)
    ref = o.execute(x)
    y, taps = emu.pfb2(x, M)
    assert np.abs(taps - o.taps()).max() <= 2e-6 * np.abs(taps).max()
    assert_parity(y, ref, what=f"firpfbch2 M={M}")
    y, _ = emu.pfb2(x, M, [1, 13, nf - 41, 27])
    assert_parity(y, ref, what=f"firpfbch2 M={M}, chunked")
