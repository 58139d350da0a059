"""Builds the TEST-ONLY CPU emulation of the kernel sources (see cuda_emu.h)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
CSRC = os.path.join(ROOT, "composable-sdr_b200", "csrc")
LIB = os.path.join(HERE, "libemu_kernels.so")


def build(force=False):
    srcs = [os.path.join(HERE, "emu_kernels.cpp"), os.path.join(HERE, "cuda_emu.h")] + \
           [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-DCSDR_EMU", "-fPIC", "-shared", "-pthread",
                               "-I" + HERE, "-I" + CSRC, "-o", LIB, os.path.join(HERE, "emu_kernels.cpp")])
    return LIB
