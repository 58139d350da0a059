#!/usr/bin/env python
"""bench.py -- throughput of the fused receive chain on N B200s (one process per GPU).

Workload (BASELINE.json configs[1], "C2"): 2.56 MS/s synthetic CF32 -> offset mix (+100 kHz) -> msresamp to
200 kHz -> dc blocker -> AGC/squelch (-40 dB) -> NBFM demod (kf 0.3).  A "step" is one pass of the chain over one
chunk of 2^LOG2N input samples that is already resident in HBM (the chunk is 2 GiB at the default 2^28, far larger
than the 126 MB L2, so no L2 flush is needed between steps); stream state carries from step to step.

    python bench.py [--gpus N --steps K --warmup W]                 # our CUDA path
    python bench.py --impl reference [--gpus N --steps K --warmup W]  # the CPU restatement (oracle port) on host cores

N > 1: launched by torchrun, rank r processes its own time segment of the one stream (seek + overlap-save warm-up,
no data-path collective) -> weak scaling; NCCL is used only for the barrier and the max-over-ranks of the timing.
Prints ONE JSON line on rank 0.

The same run also measures the other BASELINE configs, device-resident with CUDA events, and reports them under
"per_config" (C1 mix+msresamp, C3 16-channel PFB + per-channel FM, C4 1024-channel PFB + FM + --mix, C5 256 streams AM):
the channelizer chains shard by frame-aligned time segments (weak), config 5 by streams (its 256 streams are dealt to
the ranks: strong).  `--configs ""` skips them.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, OFFSET, BW, KF, AGC_DB = 2.56e6, 1e5, 200e3, 0.3, -40.0
RATE = BW / SR
B_ALG = 8.0 + 4.0 * RATE          # read CF32 once, write F32 audio once  [bytes / input sample]
METRIC = "Msamples/s CF32 through mix->resample->AGC->FM demod chain (config 2)"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic():
    """DRAM bytes per k_frontend launch from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  NVML from a
    thread (a sample every ~0.5 ms: the timed region of the default run is only ~15 ms); nvidia-smi -lms if the
    NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index, uuid=None):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.sm, self.reasons, self.smmax = [], set(), None
        self.h = None
        self.run = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)) if not str(uuid).startswith("GPU-") else str(uuid))
                except Exception:
                    h = None
            self.h = h or pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.smmax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        for bit, name in self.REASONS:
            if mask & bit:
                self.reasons.add(name)

    def _loop(self):
        while self.run:
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        if self.h is not None:
            self.run = True
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.h is not None:
            try:
                self._sample()          # the GPU is still draining the last launches: counts as under load
            except Exception:
                pass
            self.run = False
            self.t.join(timeout=1)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.smmax,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def pin_to_gpu_numa(torch, local):
    """Run this rank on the CPUs next to its GPU (NVML's affinity mask) BEFORE any pinned buffer is allocated, so that
    the staging memory of the end-to-end leg lands on the GPU's own NUMA node.  Returns the CPU list (or None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def h2d_bandwidth(torch, dist, nbytes=1 << 29, reps=4):
    """plain pinned host -> device copies on every rank at the same time: the PCIe / host-memory ceiling of the
    end-to-end leg (GB/s of this rank)"""
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.zero_()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    e1.synchronize()
    return reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9


def _fm_block(torch, n, k0, f_cyc, amp, dev_rad, fa_cyc, dtype64=True):
    """amp * exp(j (2 pi f k - dev cos(2 pi fa k))) for k = k0 .. k0 + n - 1; phases in float64, result complex64"""
    import math
    k = torch.arange(k0, k0 + n, device="cuda", dtype=torch.float64)
    ph = torch.remainder(f_cyc * k, 1.0) * (2 * math.pi) - dev_rad * torch.cos(2 * math.pi * fa_cyc * k)
    return torch.polar(torch.full((n,), float(amp), device="cuda", dtype=torch.float32), ph.to(torch.float32))


def channelizer_input(torch, n, channels, active, amp_lo, amp_hi, noise, seed, block_log2=24):
    """`active` FM carriers on channel centres f_c = (c - (C-1)/2) / C (cycles/sample) + noise.  One 2^block_log2 block
    is synthesised and repeated: centre frequencies have a period of 2C samples and the modulation a whole number of
    cycles per block, so the repetition is seamless."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    nb = min(n, 1 << block_log2)
    x = noise * torch.complex(torch.randn(nb, generator=g, device="cuda"), torch.randn(nb, generator=g, device="cuda"))
    idx = torch.randperm(channels, generator=g, device="cuda")[:active].tolist()
    for i, c in enumerate(sorted(idx)):
        amp = amp_lo + (amp_hi - amp_lo) * ((i * 7919) % active) / max(1, active - 1)
        x += _fm_block(torch, nb, 0, (c - (channels - 1) / 2.0) / channels, amp, 2.0, (1 + i % 5) * 64.0 / nb)
    x = x.to(torch.complex64)
    return x.repeat(n // nb) if n > nb else x


def am_streams_input(torch, n, streams, first_stream=0, stride=1):
    """config 5: per stream one AM carrier (index 0.8, 1 kHz tone at 10 MS/s) 437 Hz + 3 Hz * s off the +1 MHz mixer"""
    import math
    out = torch.empty((streams, n), dtype=torch.complex64, device="cuda")
    k = torch.arange(n, device="cuda", dtype=torch.float64)
    for j in range(streams):
        s = first_stream + j * stride
        g = torch.Generator(device="cuda").manual_seed(5000 + s)
        env = (0.4 * (1.0 + 0.8 * torch.cos(2 * math.pi * 1e-4 * k))).to(torch.float32)
        ph = (torch.remainder((1e6 + 437.0 + 3.0 * (s % 64)) / 10e6 * k, 1.0) * (2 * math.pi)).to(torch.float32)
        out[j] = torch.polar(env, ph) + 0.02 * torch.complex(torch.randn(n, generator=g, device="cuda"),
                                                              torch.randn(n, generator=g, device="cuda"))
    return out


def device_input(torch, n, rank):
    """config-2 signal, generated on the GPU: keyed FM carrier at +100 kHz, interferer at -400 kHz, noise."""
    import math
    g = torch.Generator(device="cuda").manual_seed(0x5D2B200 + rank)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    blk = 1 << 24
    k_first = rank * n                 # every rank synthesises its OWN stretch of the one stream
    for i in range(0, n, blk):
        m = min(blk, n - i)
        k = torch.arange(k_first + i, k_first + i + m, device="cuda", dtype=torch.float64)
        ph = 2 * math.pi * OFFSET / SR * k - 50.0 * torch.cos(2 * math.pi * 1e3 / SR * k)
        on = ((k / SR) % 0.2 < 0.05).to(torch.float64)
        sig = 0.5 * on * torch.exp(1j * ph) + 0.3 * torch.exp(-2j * math.pi * 4e5 / SR * k)
        nz = 0.05 * torch.complex(torch.randn(m, generator=g, device="cuda"), torch.randn(m, generator=g, device="cuda"))
        x[i:i + m] = sig.to(torch.complex64) + nz
    return x


def cpu_port_throughput(seconds=12.0, threads=1, log2n=23):
    """the CPU restatement (oracle port, gcc -O2) on a bounded sample: `threads` independent
    chains, each repeatedly processing a 2^log2n-sample block of the config-2 signal for ~`seconds`."""
    import numpy as np
    from oracle import oracle as O
    import composable_sdr_b200.synth as synth
    O.build()
    x = synth.config2(1 << log2n)
    chains = [O.Chain(SR, OFFSET, BW, O.DEMOD_NBFM, KF, AGC_DB, fast=True) for _ in range(threads)]
    for c in chains:                                   # warm-up + page-in
        c.process(x[:1 << 18])
    counts = [0] * threads
    stop_at = time.perf_counter() + seconds

    def work(i):
        while time.perf_counter() < stop_at:
            chains[i].process(x)
            counts[i] += x.size
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return sum(counts) / dt / 1e6, sum(counts)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    import composable_sdr_b200.synth as synth
    O.build()
    cores = os.cpu_count() or 1
    n = 1 << 22                                        # per-thread block per step
    x = synth.config2(n)
    chains = [O.Chain(SR, OFFSET, BW, O.DEMOD_NBFM, KF, AGC_DB, fast=True) for _ in range(cores)]

    def step():
        th = [threading.Thread(target=c.process, args=(x,)) for c in chains]
        for t in th:
            t.start()
        for t in th:
            t.join()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps * cores * n / dt / 1e6
    sample = f"{cores} threads x 2^22-sample block of the config-2 signal per step (independent chains, one per core)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: 2.56 MS/s CF32 -> mix 100 kHz -> msresamp 200 kHz -> dcblock -> AGC -40 dB -> NBFM",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = liquid-dsp via Haskell, cannot be built here (no GHC, no libliquid); this arm times "
                "the repo's C restatement (oracle/, parity unpinned) on all host cores",
    }))


def measure_config(torch, cs, dist, rank, world, local, chain, x, steps, warmup, seek_to=0, align=1):
    """device-resident throughput of one chain on this rank's input x ([n] or [streams, n]): CUDA events on the chain's
    stream around `steps` calls, barrier + synchronize on both sides, max over ranks.  seek_to > 0: the rank's stretch of
    the stream starts there (time-segment shard): seek + overlap-save warm-up first, outputs discarded."""
    from composable_sdr_b200 import shard
    nx = x.shape[-1]
    cap = chain.max_output(nx)
    nptr = chain.nstreams * chain.nout
    dt = torch.float32 if chain.out_dtype.__name__ == "float32" else torch.complex64
    outs = [torch.empty(max(cap, 1), dtype=dt, device="cuda") for _ in range(nptr)]
    ptrs = [o.data_ptr() for o in outs]
    torch.cuda.synchronize()
    if seek_to > 0:
        assert seek_to % align == 0
        warm = min(chain.warmup_len(), nx, seek_to)
        chain.seek(seek_to - warm)
        chain.process_raw(x.data_ptr(), warm, nx, ptrs, cap)           # history (this rank's own samples stand in)
    stream = torch.cuda.ExternalStream(chain.cuda_stream, device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        chain.process_raw(x.data_ptr(), nx, nx, ptrs, cap)
    barrier()
    l0 = cs.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ny = 0
    for _ in range(steps):
        ny = chain.process_raw(x.data_ptr(), nx, nx, ptrs, cap)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = cs.kernel_launches() - l0
    ms_max = shard.max_over_ranks(dist, ms, device="cuda") if dist is not None else ms
    return {"ms": ms_max, "outputs_per_step": int(ny), "launches_per_step": launches / steps,
            "agc_counters": list(chain.agc_counters()), "agc_plan": list(chain.agc_plan())}


def per_config(torch, cs, dist, rank, world, local, args, peak):
    """C1, C3, C4, C5 (SURVEY 8d sizes unless overridden): {value [Msamples/s, all ranks], ms_per_step, roofline, ...}"""
    from composable_sdr_b200 import shard
    want = [c for c in args.configs.split(",") if c]
    out = {}

    def entry(name, workload, b_alg, samples_per_step_all_ranks, m, scaling, sharding, steps):
        ms_step = m["ms"] / steps
        value = samples_per_step_all_ranks / (ms_step * 1e-3) / 1e6
        ach = b_alg * samples_per_step_all_ranks / world / (ms_step * 1e-3) / 1e9      # per GPU
        out[name] = {"value": value, "unit": "Msamples/s", "ms_per_step": ms_step, "steps": steps, "scaling": scaling,
                     "config": {"workload": workload, "input_samples_per_step": int(samples_per_step_all_ranks), "sharding": sharding},
                     "roofline": {"bound": "hbm", "scope": "whole chain (all kernels of a step)", "achieved": ach, "peak": peak,
                                  "unit": "GB/s", "frac": ach / peak, "bytes_per_sample": b_alg, "traffic": None},
                     "launches": m["launches_per_step"], "outputs_per_step": m["outputs_per_step"],
                     "agc_counters": m["agc_counters"], "agc_plan(L,W)": m["agc_plan"]}

    def cleanup():
        import gc
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    steps_total = lambda st: st + max(args.warmup, 3)
    if "C1" in want:
        n = 1 << args.log2n
        x = device_input(torch, n, rank)
        ch = cs.Chain(SR, OFFSET, BW, device=local)
        m = measure_config(torch, cs, dist, rank, world, local, ch, x, 10, args.warmup, seek_to=rank * steps_total(10) * n)
        entry("C1", "C1: 2.56 MS/s CF32 -> mix 100 kHz -> msresamp 200 kHz -> dcblock (DeNo)", 8.0 + 8.0 * RATE, world * n, m,
              "weak", "time segments of one stream per rank", 10)
        ch.close(); del x, ch; cleanup()
    if "C3" in want:
        n = 1 << args.log2n_c3
        x = channelizer_input(torch, n, 16, 16, 0.05, 0.2, 0.01, 300 + rank)
        ch = cs.Chain(2.56e6, demod=cs.DeNBFM(KF), agc=AGC_DB, channels=16, device=local)
        m = measure_config(torch, cs, dist, rank, world, local, ch, x, 5, args.warmup, seek_to=rank * steps_total(5) * n,
                           align=shard.frame_alignment(16))
        entry("C3", "C3: 2.56 MS/s CF32 -> dcblock -> firpfbch 16 channels -> per-channel AGC -40 dB + NBFM, 16 F32 outputs", 12.0,
              world * n, m, "weak", "frame-aligned time segments per rank, every rank produces all channels", 5)
        ch.close(); del x, ch; cleanup()
    if "C4" in want:
        n = 1 << args.log2n_c4
        x = channelizer_input(torch, n, 1024, 64, 1e-4, 7e-4, 3e-5, 400 + rank)
        ch = cs.Chain(1e9, demod=cs.DeNBFM(KF), agc=AGC_DB, channels=1024, mix_channels=True, device=local)
        m = measure_config(torch, cs, dist, rank, world, local, ch, x, 5, args.warmup, seek_to=rank * steps_total(5) * n,
                           align=shard.frame_alignment(1024))
        entry("C4", "C4: 1 GS/s CF32 -> dcblock -> firpfbch 1024 channels -> per-channel AGC -40 dB + NBFM -> --mix (sum), 1 F32 output",
              8.0 + 4.0 / 1024, world * n, m, "weak",
              "frame-aligned time segments per rank; --mix is a per-rank sum over all 1024 channels, no collective", 5)
        ch.close(); del x, ch; cleanup()
    if "C4b" in want:
        # the channelizer the task names: liquid's 2x oversampled firpfbch2_crcf analyzer (frames of C/2 samples, every channel
        # at 2/C of the input rate) in place of the reference's firpfbch_crcf
        n = 1 << args.log2n_c4b
        x = channelizer_input(torch, n, 1024, 64, 1e-4, 7e-4, 3e-5, 400 + rank)
        ch = cs.Chain(1e9, demod=cs.DeNBFM(KF), agc=AGC_DB, channels=1024, mix_channels=True, device=local, channelizer=1)
        m = measure_config(torch, cs, dist, rank, world, local, ch, x, 5, args.warmup, seek_to=rank * steps_total(5) * n,
                           align=shard.frame_alignment(1024))
        entry("C4b", "C4b: 1 GS/s CF32 -> dcblock -> firpfbch2 1024 channels (2x oversampled) -> per-channel AGC -40 dB + NBFM -> --mix (sum), "
              "1 F32 output at twice C4's rate", 8.0 + 8.0 / 1024, world * n, m, "weak",
              "frame-aligned time segments per rank; --mix is a per-rank sum, no collective", 5)
        ch.close(); del x, ch; cleanup()
    if "C5" in want:
        n = 1 << args.log2n_c5
        total_streams = 256
        mine = shard.stream_shard(total_streams, world, rank)
        x = am_streams_input(torch, n, len(mine), first_stream=rank, stride=world)
        ch = cs.Chain(10e6, 1e6, 200e3, cs.DeAM(), agc=AGC_DB, nstreams=len(mine), device=local)
        m = measure_config(torch, cs, dist, rank, world, local, ch, x, 3, args.warmup)
        entry("C5", "C5: 256 streams x 10 MS/s CF32 -> mix 1 MHz -> msresamp 200 kHz -> dcblock -> AGC -40 dB -> ampmodem(0.8, DSB)",
              8.0 + 4.0 * 0.02, total_streams * n, m, "strong", f"256 streams dealt round-robin, {len(mine)} per rank", 3)
        ch.close(); del x, ch; cleanup()
    return out


def run_ours(args):
    import torch
    import composable_sdr_b200 as cs
    from composable_sdr_b200 import shard
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpus = pin_to_gpu_numa(torch, local)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cs.build.build()
    n = 1 << args.log2n
    x = device_input(torch, n, rank)
    chain = cs.Chain(SR, OFFSET, BW, cs.DeNBFM(KF), agc=AGC_DB, device=local)
    cap = chain.max_output(n)
    out = torch.empty(cap, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    # time-segment sharding (composable_sdr_b200/shard.py): the stream is world * (steps + warmup) chunks long and rank r
    # owns the r-th contiguous segment: seek to its start minus the warm-up, feed the warm-up history, drop its outputs
    seg_total = world * (args.steps + max(args.warmup, 3)) * n
    seek, first, stop, start = shard.shard_input_range(seg_total, world, rank, chain.warmup_len())
    if start > 0:
        chain.seek(seek)
        wl = min(start - first, n)
        chain.process_raw(x.data_ptr(), wl, wl, [out.data_ptr()], cap)      # overlap-save history (discarded)
    stream = torch.cuda.ExternalStream(chain.cuda_stream, device=torch.device("cuda", local))

    def step():
        return chain.process_raw(x.data_ptr(), n, n, [out.data_ptr()], cap)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    chain.profile(True)
    launches0 = cs.kernel_launches()
    try:
        uuid = str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
    except Exception:
        uuid = None
    clocks = ClockSampler(local, uuid)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ny = 0
    for _ in range(args.steps):
        ny = step()
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = cs.kernel_launches() - launches0
    fe_ms, fe_launches = chain.frontend_ms()
    chain.profile(False)
    fixups = chain.agc_fixups()
    plan = chain.agc_plan()
    ms_max = shard.max_over_ranks(dist, ms, device="cuda") if dist is not None else ms
    value = world * args.steps * n / (ms_max * 1e-3) / 1e6

    # ---- end to end through the public C ABI with HOST buffers (pinned): H2D + chain + D2H inside the timed region
    ne = 1 << min(args.log2n, 26)
    xh = cs.PinnedBuffer(ne, "complex64")
    xh.array[:] = x[:ne].cpu().numpy()
    del x, out
    e2e_chain = cs.Chain(SR, OFFSET, BW, cs.DeNBFM(KF), agc=AGC_DB, device=local)
    cap_e = e2e_chain.max_output(ne)
    oh = cs.PinnedBuffer(cap_e, "float32")
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        nye = e2e_chain.process_raw(xh.array.ctypes.data, ne, ne, [oh.array.ctypes.data], cap_e)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        nye = e2e_chain.process_raw(xh.array.ctypes.data, ne, ne, [oh.array.ctypes.data], cap_e)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dt_max = shard.max_over_ranks(dist, dt, device="cuda") if dist is not None else dt
    e2e_val = world * e2e_steps * ne / dt_max / 1e6
    checksum = float(oh.array[:nye].astype("float64").sum())
    e2e_chain.close()
    # the ceiling of that leg: plain pinned H2D copies on all ranks at once (GB/s per rank)
    h2d = h2d_bandwidth(torch, dist)
    if dist is not None:
        t = torch.tensor([h2d], dtype=torch.float64, device="cuda")
        allh = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allh, t)
        h2d_all = [round(float(v.item()), 2) for v in allh]
    else:
        h2d_all = [round(h2d, 2)]
    chain.close()
    del chain
    torch.cuda.empty_cache()

    peak, peak_kind = measured_peak()
    pc = per_config(torch, cs, dist, rank, world, local, args, peak) if args.configs else {}

    if rank == 0:
        # the chain may split a chunk into parts (one k_frontend launch each): bytes per launch = bytes per step / parts
        parts = max(1, fe_launches // max(args.steps, 1))
        ach = B_ALG * (n / parts) / (fe_ms / max(fe_launches, 1) * 1e-3) / 1e9 if fe_ms > 0 else None
        tr = ncu_traffic()
        ms_step = ms_max / args.steps
        chain_ach = B_ALG * n / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 2.56 MS/s CF32 -> mix 100 kHz -> msresamp 200 kHz -> dcblock -> AGC -40 dB -> NBFM",
                       "chunk_samples": n, "l2": f"inputs ({n * 8 / 2**30:g} GiB/step) larger than L2, no flush",
                       "sharding": "time segments of one stream per rank (seek + overlap-save warm-up), no collective"},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": ne * 8, "d2h_bytes_per_step": int(nye) * 4,
                    "steps": e2e_steps, "chunk_samples": ne, "checksum": checksum,
                    "h2d_gbs_per_rank_all_ranks_copying": h2d_all, "cpu_affinity": (f"{len(cpus)} cpus next to the GPU (NVML)" if cpus else None),
                    "h2d_ceiling_Msamples_s": sum(h2d_all) / 8.0 * 1e3},
            "roofline": {"bound": "hbm", "kernel": "k_frontend (mix + half-band cascade + arbitrary resampler)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                         "peak_kind": peak_kind, "bytes_per_sample": B_ALG, "launches": fe_launches,
                         "avg_launch_ms": fe_ms / max(fe_launches, 1),
                         "share_of_step": fe_ms / ms if ms > 0 else None,
                         # DRAM bytes of one launch from the committed ncu --set full capture, scaled from the
                         # captured chunk size to this run's chunk (traffic is linear in the samples streamed)
                         "traffic": (tr["dram_bytes_per_launch"] * (n / parts) / tr["chunk_samples"]) if tr else None,
                         "samples_per_launch": n // parts,
                         "traffic_source": (tr or {}).get("source"),
                         # the whole step (all kernels) against the same peak: the north star's ">= 60 % on the fused chain"
                         "chain_achieved": chain_ach, "chain_frac": chain_ach / peak,
                         "launches_per_step": launches / args.steps},
            "outputs_per_step": int(ny), "agc_fixups": int(fixups), "agc_plan(L,W)": list(plan),
            "per_config": pc,
        }
        if world == 1 and not args.no_cpu:
            v, cnt = cpu_port_throughput(seconds=args.cpu_seconds, threads=1)
            from oracle import real_liquid
            line["cpu_baseline"] = {"value": v, "unit": "Msamples/s", "cores": 1, "kind": "port",
                                    "sample": f"{cnt} samples: a 2^23-sample block of the config-2 signal, repeated for "
                                              f"~{args.cpu_seconds:.0f} s on one core (the reference is single-threaded); "
                                              "the port is gcc -O2 scalar C with memmove delay lines, not liquid's SIMD dotprod",
                                    "real_libliquid": real_liquid.find()}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28, help="log2 of the chunk size in samples per step per GPU")
    ap.add_argument("--configs", default="C1,C3,C4,C4b,C5", help="other configs measured in the same run (per_config); '' = none")
    ap.add_argument("--log2n-c3", type=int, default=28, help="log2 of the config-3 chunk (samples per step per GPU)")
    ap.add_argument("--log2n-c4", type=int, default=30, help="log2 of the config-4 chunk (samples per step per GPU)")
    ap.add_argument("--log2n-c4b", type=int, default=28, help="log2 of the firpfbch2 variant of config 4 (samples per step per GPU)")
    ap.add_argument("--log2n-c5", type=int, default=22, help="log2 of the config-5 chunk per stream (256 streams in all)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
