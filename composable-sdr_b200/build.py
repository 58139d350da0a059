"""Build libcsdr_b200.so (nvcc, sm_100a only) and the liquid-named alias library, in-tree."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcsdr_b200.so")
COMPAT = os.path.join(HERE, "libcsdr_liquid_compat.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libcsdr_b200.so can only be built with the CUDA toolkit (no CPU fallback)")
    return p


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    srcs.append(os.path.join(HERE, "..", "include", "csdr_b200.h"))
    if force or _stale(LIB, srcs):
        # CSDR_NVCC_EXTRA: extra -D flags of the timing experiments (scripts/exp_*.py); never set in a normal build
        cmd = [nvcc_path()] + NVCC_FLAGS + os.environ.get("CSDR_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB, os.path.join(CSRC, "csdr_b200.cu")]
        subprocess.check_call(cmd)
    compat_src = os.path.join(CSRC, "liquid_compat.c")
    if force or _stale(COMPAT, [compat_src, LIB]):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", COMPAT, compat_src,
                               "-L" + HERE, "-lcsdr_b200", "-ldl", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
