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


def time_segments(total, world):
    """[(start, stop)] of `world` contiguous segments covering [0, total)."""
    base, rem = divmod(int(total), int(world))
    out, pos = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((pos, pos + n))
        pos += n
    return out


def stream_shard(nstreams, world, rank):
    """stream indices owned by `rank` (round-robin)."""
    return list(range(rank, nstreams, world))


def shard_input_range(total, world, rank, warmup):
    """(seek_position, first_input_sample, stop_sample, samples_to_discard_from) for a time-segment shard:
    the rank feeds input[first:stop] after seeking to `first`; inputs before `start` are warm-up."""
    start, stop = time_segments(total, world)[rank]
    first = max(0, start - int(warmup))
    return first, first, stop, start


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
