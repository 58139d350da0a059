"""time k_frontend for one variant (usage: exp_variant.py VARIANT(0|1) [label])"""
import sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import composable_sdr_b200 as cs
if os.environ.get("CSDR_EXP_LIB"):                 # a variant library built with CSDR_NVCC_EXTRA (see exp_build.sh)
    from composable_sdr_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(os.environ["CSDR_EXP_LIB"])
from bench_configs import sig
n = 1 << 27
x = sig(n, 1)
variant = int(sys.argv[1])          # 1 = k_frontend_direct (default), 0 = k_frontend_std
cs.set_option(9, variant)
if len(sys.argv) > 3: cs.set_option(10, int(sys.argv[3]))
ch = cs.Chain(2.56e6, 1e5, 200e3)
cap = ch.max_output(n)
out = torch.empty(cap, dtype=torch.complex64, device="cuda")
for _ in range(3):
    ch.process_raw(x.data_ptr(), n, n, [out.data_ptr()], cap)
ch.profile(True)
for _ in range(10):
    ch.process_raw(x.data_ptr(), n, n, [out.data_ptr()], cap)
ms, k = ch.frontend_ms()
print(f"{sys.argv[2] if len(sys.argv) > 2 else ''} variant {variant}: k_frontend {ms / k * 1e3:.1f} us per 2^27 samples, "
      f"{8.3125 * n / (ms / k * 1e-3) / 1e9 / 6540.2:.3f} of measured HBM roofline", flush=True)
