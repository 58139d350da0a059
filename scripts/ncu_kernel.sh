#!/bin/bash
# one `ncu --set full` capture of the kernels matching a regex inside a short bench run
# usage: scripts/ncu_kernel.sh <tag> <kernel regex> <skip> <count> [bench args]
tag=$1; re=$2; skip=$3; cnt=$4; shift 4
ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c $cnt -f -o gpurun_out/${tag} \
    python bench.py --steps 1 --warmup 3 --no-cpu "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
