import sys, os, json
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import composable_sdr_b200 as cs
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_configs import sig
x = sig(1 << 24, 1)
for name, opts in [("default", {})]:
    for k, v in opts.items():
        cs.set_option(k, v)
    ch = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
    import time
    for it in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        y = ch.process(x)[0]
        torch.cuda.synchronize()
        print(name, it, ch.agc_counters(), round((time.perf_counter() - t0) * 1e3, 3), 'ms', flush=True)
    for k in opts:
        cs.set_option(k, {3: 512, 4: 384, 6: 0, 8: 0}[k])
    ch.close()
