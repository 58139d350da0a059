"""Two (or more) GPUs of one box: the shards of a stream, each on its own device, concatenate to the single-stream result
(time segments with overlap-save history, no data-path collective).  Skipped on a one-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import numpy as np
import pytest

from util import REL_TOL_AFTER_DCBLOCK, REL_TOL_FM_NOISE, assert_parity, snr_db

pytestmark = pytest.mark.gpu


def _need_two(cs):
    if cs.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_config2_time_segments_on_two_devices(cs, orc):
    from composable_sdr_b200 import shard
    _need_two(cs)
    world = min(cs.device_count(), 4)
    n = 3 << 21
    x = cs.synth.config2(n, keyed=None)
    ref = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_NBFM, 0.3, -40.0).process(x)[0]
    parts = []
    for rank, (a, b) in enumerate(shard.time_segments(n, world)):
        ch = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0, device=rank)
        shard.seek_shard(ch, a, lambda i, j: x[i:j])
        parts.append(ch.process(x[a:b])[0])
        ch.close()
    y = np.concatenate(parts)
    assert len(y) == len(ref)
    assert np.count_nonzero((y == 0) != (ref == 0)) == 0
    assert_parity(y, ref, rel=REL_TOL_AFTER_DCBLOCK, period=1 / 0.3, what=f"config 2 on {world} devices")


def test_channelizer_mix_time_segments_on_two_devices(cs, orc):
    """config 4's shape (channelizer + per-channel FM + --mix), 64 channels: frame-aligned shards, one per device; the
    --mix sum needs no collective because every shard holds all channels of its frames"""
    from composable_sdr_b200 import shard
    _need_two(cs)
    world = min(cs.device_count(), 4)
    x = cs.synth.config4(1 << 22, channels=64, active=8, sr=1e8)
    ref = orc.Chain(1e8, 0.0, 0.0, orc.DEMOD_NBFM, 0.3, -40.0, 64, True).process(x)[0]
    parts = []
    for rank, (a, b) in enumerate(shard.time_segments(len(x), world, shard.frame_alignment(64))):
        ch = cs.Chain(1e8, demod=cs.DeNBFM(0.3), agc=-40.0, channels=64, mix_channels=True, device=rank)
        shard.seek_shard(ch, a, lambda i, j: x[i:j])
        parts.append(ch.process(x[a:b])[0])
        ch.close()
    y = np.concatenate(parts)
    assert len(y) == len(ref)
    assert snr_db(y[512:], ref[512:]) >= 60.0


def test_config5_streams_dealt_to_devices(cs, orc):
    from composable_sdr_b200 import shard
    _need_two(cs)
    world = min(cs.device_count(), 4)
    S, n = 8, 1 << 20
    x = cs.synth.config5(n, S)
    for rank in range(world):
        mine = shard.stream_shard(S, world, rank)
        ch = cs.Chain(10e6, 1e6, 200e3, agc=-40.0, nstreams=len(mine), device=rank)
        outs = ch.process(np.ascontiguousarray(x[mine]))
        ch.close()
        for k, s in enumerate(mine):
            pre = orc.Chain(10e6, 1e6, 200e3, orc.DEMOD_NO, 0.0, -40.0).process(x[s])[0]
            assert len(outs[k]) == len(pre)
            assert np.array_equal(outs[k] == 0, pre == 0)
            assert_parity(outs[k][64:], pre[64:], rel=REL_TOL_AFTER_DCBLOCK, what=f"stream {s} on device {rank}")
