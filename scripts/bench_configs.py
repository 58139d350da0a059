#!/usr/bin/env python
"""Device-resident throughput of the five BASELINE configs on one GPU (secondary numbers; bench.py is the contract).
Prints one JSON line per config."""
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import composable_sdr_b200 as cs  # noqa: E402


def run(name, chain, x, steps=5, warmup=2, b_alg=None):
    nx = x.shape[-1]
    cap = chain.max_output(nx)
    nptr = chain.nstreams * chain.nout
    dt = torch.float32 if chain.out_dtype.__name__ == "float32" else torch.complex64
    outs = [torch.empty(max(cap, 1), dtype=dt, device="cuda") for _ in range(nptr)]
    ptrs = [o.data_ptr() for o in outs]
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(chain.cuda_stream)
    for _ in range(warmup):
        chain.process_raw(x.data_ptr(), nx, nx, ptrs, cap)
    torch.cuda.synchronize()
    l0 = cs.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        n = chain.process_raw(x.data_ptr(), nx, nx, ptrs, cap)
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tot = x.numel()
    line = {"config": name, "input_samples_per_step": tot, "ms_per_step": ms, "Msamples_per_s": tot / ms / 1e3,
            "outputs_per_step": n, "launches_per_step": (cs.kernel_launches() - l0) / steps,
            "agc_counters(seq gain, seq fsm, refined)": chain.agc_counters()}
    if b_alg:
        line["hbm_frac_of_6540GBs"] = b_alg * tot / (ms * 1e-3) / 1e9 / 6540.2
    print(json.dumps(line), flush=True)


def sig(n, seed, scale=0.3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    k = torch.arange(n, device="cuda", dtype=torch.float64)      # float32 loses integer resolution above 2^24
    # tone at 0.15 rad/sample (inside the 200 kHz passband after the 100 kHz offset mix), slow FM
    ph = torch.remainder(0.15 * k + 2.0 * torch.sin(k * 2e-3), 2 * 3.141592653589793).to(torch.float32)
    del k
    x = scale * torch.polar(torch.ones(n, device="cuda"), ph)
    x = x + 0.02 * torch.complex(torch.randn(n, generator=g, device="cuda"), torch.randn(n, generator=g, device="cuda"))
    return x.to(torch.complex64)


def sig_am(n, seed, sr=10e6, fc=1e6 + 437.0, amp=0.4, index=0.8, tone=1e3):
    """config-5 stream: AM carrier (index 0.8, 1 kHz tone) 437 Hz off the mixer frequency, plus noise"""
    import math
    g = torch.Generator(device="cuda").manual_seed(seed)
    k = torch.arange(n, device="cuda", dtype=torch.float64)
    env = (amp * (1.0 + index * torch.cos(2 * math.pi * tone / sr * k))).to(torch.float32)
    ph = torch.remainder(2 * math.pi * fc / sr * k, 2 * math.pi).to(torch.float32)
    del k
    x = torch.polar(env, ph)
    x = x + 0.02 * torch.complex(torch.randn(n, generator=g, device="cuda"), torch.randn(n, generator=g, device="cuda"))
    return x.to(torch.complex64)


def main():
    n = 1 << 26
    x = sig(n, 1)
    torch.cuda.synchronize()       # the chains run on their own streams
    run("C1 mix+msresamp+dcblock (DeNo)", cs.Chain(2.56e6, 1e5, 200e3), x, b_alg=8 + 8 * 0.078125)
    run("C2 + AGC + NBFM", cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0), x, b_alg=8 + 4 * 0.078125)
    run("C2w + AGC + WBFM decim 4 (SURVEY 8f N2)", cs.Chain(2.56e6, 1e5, 200e3, cs.DeWBFM(4), agc=-40.0), x, b_alg=8 + 1 * 0.078125)
    x3 = sig(1 << 24, 3)
    run("C3 16-ch PFB + per-channel AGC + NBFM", cs.Chain(2.56e6, demod=cs.DeNBFM(0.3), agc=-40.0, channels=16), x3, b_alg=12)
    run("C3 16-ch PFB, DeNo, no AGC", cs.Chain(2.56e6, channels=16), x3, b_alg=16)
    run("C4 1024-ch PFB + AGC + NBFM + mix", cs.Chain(1e9, demod=cs.DeNBFM(0.3), agc=-40.0, channels=1024, mix_channels=True),
        sig(1 << 24, 4, 3e-4), b_alg=8.004)
    S = 256
    x5 = torch.stack([sig_am(1 << 18, 50 + s, fc=1e6 + 437.0 + 3.0 * (s % 64)) for s in range(S)])
    torch.cuda.synchronize()
    run("C5 256 streams x 2^18: mix+msresamp+AGC+AM", cs.Chain(10e6, 1e6, 200e3, cs.DeAM(), agc=-40.0, nstreams=S), x5,
        steps=3, warmup=1, b_alg=8.08)


if __name__ == "__main__":
    main()
