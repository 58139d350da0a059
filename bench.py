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


def device_input(torch, n, rank):
    """config-2 signal, generated on the GPU: keyed FM carrier at +100 kHz, interferer at -400 kHz, noise."""
    import math
    g = torch.Generator(device="cuda").manual_seed(0x5D2B200 + rank)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    blk = 1 << 24
    for i in range(0, n, blk):
        m = min(blk, n - i)
        k = torch.arange(i, i + m, device="cuda", dtype=torch.float64)
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


def run_ours(args):
    import torch
    import composable_sdr_b200 as cs
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
    # time-segment sharding: rank r owns samples [r*seg, (r+1)*seg) of the one stream
    seg = (args.steps + args.warmup) * n
    if rank > 0:
        warm = chain.warmup_len()
        chain.seek(rank * seg - warm)
        wl = min(warm, n)
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
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * args.steps * n / (ms_max * 1e-3) / 1e6

    # ---- end to end through the public C ABI with HOST buffers (pinned): H2D + chain + D2H inside the timed region
    ne = 1 << min(args.log2n, 26)
    xh = cs.PinnedBuffer(ne, "complex64")
    xh.array[:] = x[:ne].cpu().numpy()
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
    te = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_steps * ne / float(te.item()) / 1e6
    checksum = float(oh.array[:nye].astype("float64").sum())

    if rank == 0:
        peak, peak_kind = measured_peak()
        # the chain may split a chunk into parts (one k_frontend launch each): bytes per launch = bytes per step / parts
        parts = max(1, fe_launches // max(args.steps, 1))
        ach = B_ALG * (n / parts) / (fe_ms / max(fe_launches, 1) * 1e-3) / 1e9 if fe_ms > 0 else None
        tr = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 2.56 MS/s CF32 -> mix 100 kHz -> msresamp 200 kHz -> dcblock -> AGC -40 dB -> NBFM",
                       "chunk_samples": n, "l2": f"inputs ({n * 8 / 2**30:g} GiB/step) larger than L2, no flush",
                       "sharding": "time segments of one stream per rank (seek + overlap-save warm-up), no collective"},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": ne * 8, "d2h_bytes_per_step": int(nye) * 4,
                    "steps": e2e_steps, "chunk_samples": ne, "checksum": checksum},
            "roofline": {"bound": "hbm", "kernel": "k_frontend (mix + half-band cascade + arbitrary resampler)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                         "peak_kind": peak_kind, "bytes_per_sample": B_ALG, "launches": fe_launches,
                         "avg_launch_ms": fe_ms / max(fe_launches, 1),
                         "share_of_step": fe_ms / ms if ms > 0 else None,
                         # DRAM bytes of one launch from the committed ncu --set full capture, scaled from the
                         # captured chunk size to this run's chunk (traffic is linear in the samples streamed)
                         "traffic": (tr["dram_bytes_per_launch"] * (n / parts) / tr["chunk_samples"]) if tr else None,
                         "samples_per_launch": n // parts,
                         "traffic_source": (tr or {}).get("source")},
            "outputs_per_step": int(ny), "agc_fixups": int(fixups),
        }
        if world == 1 and not args.no_cpu:
            v, cnt = cpu_port_throughput(seconds=args.cpu_seconds, threads=1)
            line["cpu_baseline"] = {"value": v, "unit": "Msamples/s", "cores": 1, "kind": "port",
                                    "sample": f"{cnt} samples: a 2^23-sample block of the config-2 signal, repeated for "
                                              f"~{args.cpu_seconds:.0f} s on one core (the reference is single-threaded)"}
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
