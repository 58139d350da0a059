"""ncu target: config 5 (batch of streams, AM)"""
import sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import composable_sdr_b200 as cs
from bench_configs import sig_am
S, n = 64, 1 << 20
x = torch.stack([sig_am(n, 10 + s, fc=1e6 + 437.0 + 3.0 * s) for s in range(S)])
torch.cuda.synchronize()
ch = cs.Chain(10e6, 1e6, 200e3, cs.DeAM(), agc=-40.0, nstreams=S)
cap = ch.max_output(n)
outs = [torch.empty(max(cap, 1), dtype=torch.float32, device="cuda") for _ in range(S)]
ptrs = [o.data_ptr() for o in outs]
for _ in range(3):
    ch.process_raw(x.data_ptr(), n, n, ptrs, cap)
torch.cuda.synchronize()
