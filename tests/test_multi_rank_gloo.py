"""world_size-2 test of the multi-GPU host logic on CPU (gloo): time-segment plan, seek + overlap-save warm-up,
output-count exchange, concatenation and the max-over-ranks timing reduce.  The per-rank front end is the real kernel
source under the test-only CPU emulator (no GPU here); the gathered result must equal the single-stream oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from composable_sdr_b200 import shard
    from emu import emulib
    from oracle import oracle as O
    from util import make_signal
    emu = emulib.load()
    total, rate, warm = 48000, 0.078125, 2048
    f = float(np.float32(0.3))
    x = make_signal(total, 77)
    seek, first, stop, start = shard.shard_input_range(total, world, rank, warm)
    y = emu.frontend(x[first:stop], rate, freq=f, seek=seek)
    # outputs produced by the warm-up part are dropped: their count is what a chain fed x[first:start] emits
    n_warm = len(emu.frontend(x[first:start], rate, freq=f, seek=seek)) if start > first else 0
    y = y[n_warm:]
    counts = shard.gather_counts(dist, len(y))
    t_max = shard.max_over_ranks(dist, 1.0 + rank)
    buf = torch.zeros(max(counts), dtype=torch.complex64)
    buf[:len(y)] = torch.from_numpy(y)
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    if rank == 0:
        full = np.concatenate([g[:c].numpy() for g, c in zip(gathered, counts)])
        ref = O.MsResamp(rate).execute(O.Nco(f).mix_down(x))
        ok = len(full) == len(ref) and float(np.max(np.abs(full - ref))) <= 1e-4 * float(np.max(np.abs(ref)))
        q.put((ok, len(full), len(ref), t_max, counts))
    dist.barrier()
    dist.destroy_process_group()


def test_time_segments_cover_the_stream():
    from composable_sdr_b200 import shard
    for total, world in [(10, 3), (1 << 20, 8), (7, 8)]:
        segs = shard.time_segments(total, world)
        assert segs[0][0] == 0 and segs[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:]))
    assert shard.stream_shard(10, 4, 1) == [1, 5, 9]
    # channelizer chains: interior boundaries on whole frames
    for total, world, align in [(1000, 3, 64), (1 << 24, 8, 1024), (100, 4, 40)]:
        segs = shard.time_segments(total, world, align)
        assert segs[0][0] == 0 and segs[-1][1] == total
        assert all(a[1] == b[0] and a[1] % align == 0 for a, b in zip(segs, segs[1:]))
    assert shard.frame_alignment(1024) == 1024 and shard.frame_alignment(20, 1, 2) == 40 and shard.frame_alignment(16, 5, 64) == 1024


@pytest.mark.timeout(300)
def test_two_ranks_time_sharding_equals_single_stream():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n, nref, t_max, counts = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert ok, (n, nref, counts)
    assert t_max == 2.0 and sum(counts) == nref
