"""Partitioning of the receive chain over the GPUs of one box.  No data-path collective is needed:

* independent streams (config 5): streams are dealt round-robin to ranks;
* one long stream (configs 1/2): contiguous time segments, one per rank.  A rank seeks its chain to
  `start - warmup` (closed-form NCO phase, half-band block alignment and resampler timing), feeds `warmup` samples of
  real history and drops the outputs they produce (overlap-save);
* channelizer (configs 3/4): time segments as well (every rank produces all channels of its segment), so `--mix`
  is a per-rank sum and results are only concatenated.

torch.distributed (NCCL on the GPUs, gloo in the CPU tests) is used for the barrier, the max-over-ranks of the
timing and, if the caller wants everything on one rank, the gather of the outputs.
"""


def time_segments(total, world, align=1):
    """[(start, stop)] of `world` contiguous segments covering [0, total); every interior boundary is a multiple of
    `align` input samples (channelizer chains: a whole number of frames, so that each frame has one owner)."""
    total, world, align = int(total), int(world), max(1, int(align))
    units = total // align
    base, rem = divmod(units, world)
    out, pos = [], 0
    for r in range(world):
        n = (base + (1 if r < rem else 0)) * align
        stop = total if r == world - 1 else pos + n
        out.append((pos, stop))
        pos = stop
    return out


def stream_shard(nstreams, world, rank):
    """stream indices owned by `rank` (round-robin)."""
    return list(range(rank, nstreams, world))


def shard_input_range(total, world, rank, warmup, align=1):
    """(seek_position, first_input_sample, stop_sample, samples_to_discard_from) for a time-segment shard:
    the rank feeds input[first:stop] after seeking to `first`; inputs before `start` are warm-up."""
    start, stop = time_segments(total, world, align)[rank]
    first = max(0, start - int(warmup))
    return first, first, stop, start


def frame_alignment(channels, rate_num=1, rate_den=1):
    """input samples per whole number of channelizer frames when the resampler's rate is rate_num / rate_den exactly
    (1/1 = no resampler): a boundary at a multiple of it falls on a frame boundary of the stream"""
    c = max(1, int(channels)) * int(rate_den)
    return c // _gcd(c, int(rate_num))


def _gcd(a, b):
    while b:
        a, b = b, a % b
    return a


def seek_shard(chain, start, feed, drop_outputs=True):
    """Position `chain` (a fresh composable_sdr_b200.Chain) at input sample `start` of its stream: seek to start - warm-up,
    run the warm-up history through it -- `feed(first, start)` must return those input samples, numpy or torch -- and
    discard what it produces.  Returns the number of warm-up samples fed."""
    warm = min(int(start), chain.warmup_len())
    if start <= 0:
        return 0
    chain.seek(start - warm)
    if warm:
        chain.process(feed(start - warm, start))
    return warm


def gather_counts(dist, n_local, device=None):
    """all ranks learn how many outputs every rank produced (for concatenation offsets)."""
    import torch
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    outs = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [int(o.item()) for o in outs]


def max_over_ranks(dist, value, device=None):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
