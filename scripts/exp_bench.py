"""bench.py against a variant library (timing experiments): CSDR_EXP_LIB=exp/<name>.so python scripts/exp_bench.py [bench args]"""
import os, sys, runpy
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import composable_sdr_b200  # noqa: F401  (path shim)
from composable_sdr_b200 import _lib
if os.environ.get("CSDR_EXP_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["CSDR_EXP_LIB"])
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
