"""bench.py with library options set first: exp_opts.py OPT=VALUE[,OPT=VALUE...] [bench args]; then a config-2 parity check against the oracle"""
import os, sys, runpy, json, io, contextlib
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import composable_sdr_b200 as cs
opts = [kv.split("=") for kv in sys.argv[1].split(",") if kv]
for k, v in opts: cs.set_option(int(k), int(v))
# parity of config 2 and a 16-channel chain with these options
from oracle import oracle as orc
from util import assert_parity, REL_TOL_AFTER_DCBLOCK
n = 1 << 22
x = cs.synth.config2(n, keyed=None)
ref = orc.Chain(2.56e6, 1e5, 200e3, orc.DEMOD_NBFM, 0.3, -40.0).process(x)[0]
ch = cs.Chain(2.56e6, 1e5, 200e3, cs.DeNBFM(0.3), agc=-40.0)
y = np.concatenate([ch.process(x[i:i + (1 << 20)])[0] for i in range(0, n, 1 << 20)])
gate_mismatch = int(np.count_nonzero((y == 0) != (ref == 0)))
assert_parity(y, ref, rel=REL_TOL_AFTER_DCBLOCK, period=1 / 0.3, what="config 2")
print("options", sys.argv[1], "parity ok, gate mismatches", gate_mismatch, flush=True)
ch.close()
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    runpy.run_path(sys.argv[0], run_name="__main__")
for l in buf.getvalue().splitlines():
    if l.startswith("{"):
        d = json.loads(l)
        print("   C2 %.4f ms plan %s fixups %s |" % (d["ms_per_step"], d.get("agc_plan(L,W)"), d.get("agc_fixups")),
              {k: (round(v["ms_per_step"], 4), v["agc_plan(L,W)"], v["agc_counters"]) for k, v in d["per_config"].items()}, flush=True)
