"""ncu target: a few front-end-only launches of one variant (usage: prof_variant.py VARIANT [log2n])"""
import sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import composable_sdr_b200 as cs
from bench_configs import sig
variant = int(sys.argv[1]); n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 26)
x = sig(n, 1)
cs.set_option(9, variant)
ch = cs.Chain(2.56e6, 1e5, 200e3)
cap = ch.max_output(n)
out = torch.empty(cap, dtype=torch.complex64, device="cuda")
for _ in range(5):
    ch.process_raw(x.data_ptr(), n, n, [out.data_ptr()], cap)
torch.cuda.synchronize()
