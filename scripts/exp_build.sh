#!/bin/bash
# build a variant of libcsdr_b200.so with extra -D flags into exp/<name>.so (timing experiments only)
# usage: scripts/exp_build.sh <name> <flags...>
name=$1; shift
mkdir -p exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -diag-suppress 177 "$@" \
     -o exp/$name.so composable-sdr_b200/csrc/csdr_b200.cu
