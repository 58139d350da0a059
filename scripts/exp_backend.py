"""bench-like timing of the whole chain for a few AGC warm-up settings (usage: exp_backend.py)"""
import sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import composable_sdr_b200 as cs
import bench
for log2n in (26, 27):
    n = 1 << log2n
    x = bench.device_input(torch, n, 0)
    torch.cuda.synchronize()       # the chain runs on its own stream
    for W in (384, 256, 192):
        cs.set_option(4, W)
        ch = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
        cap = ch.max_output(n)
        out = torch.empty(cap, dtype=torch.float32, device="cuda")
        for _ in range(5):
            ch.process_raw(x.data_ptr(), n, n, [out.data_ptr()], cap)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ch.profile(True)
        st = torch.cuda.ExternalStream(ch.cuda_stream)
        c0 = ch.agc_counters()
        e0.record(st)
        for _ in range(10):
            ch.process_raw(x.data_ptr(), n, n, [out.data_ptr()], cap)
        e1.record(st); e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fms, k = ch.frontend_ms()
        print(f"2^{log2n} W={W}: step {ms*1e3:.1f} us, front end {fms/k*1e3:.1f} us, rest {ms*1e3 - fms/k*1e3:.1f} us, counters over the 10 timed calls {tuple(a - b for a, b in zip(ch.agc_counters(), c0))}, (L, W) = {ch.agc_plan()}", flush=True)
        ch.close()
    del x
cs.set_option(4, 384)
