"""ncu target: config 3 / 4 chains (usage: prof_c3.py CHANNELS [agc])"""
import sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import composable_sdr_b200 as cs
if os.environ.get('CSDR_EXP_LIB'):
    from composable_sdr_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(os.environ['CSDR_EXP_LIB'])
from bench_configs import sig
C = int(sys.argv[1]); agc = len(sys.argv) > 2
n = 1 << int(os.environ.get('PROF_LOG2N', '24'))
x = sig(n, 3, 0.3 if C < 100 else 3e-4)
torch.cuda.synchronize()
ch = cs.Chain(2.56e6 if C < 100 else 1e9, demod=cs.DeNBFM(0.3) if agc else None, agc=-40.0 if agc else 0.0, channels=C, mix_channels=(C >= 100 and agc)) if agc else cs.Chain(2.56e6, channels=C)
cap = ch.max_output(n)
nptr = ch.nstreams * ch.nout
dt = torch.float32 if agc else torch.complex64
outs = [torch.empty(max(cap, 1), dtype=dt, device="cuda") for _ in range(nptr)]
ptrs = [o.data_ptr() for o in outs]
for _ in range(4):
    ch.process_raw(x.data_ptr(), n, n, ptrs, cap)
torch.cuda.synchronize()
