#!/bin/bash
# per-kernel durations of the bench steps (ncu launch list, cold caches, serialised): SHARES, not absolutes
# usage: scripts/kernel_times.sh <tag> [bench args]
tag=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c ${NCU_COUNT:-600} --csv --log-file gpurun_out/${tag}_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu "$@" > gpurun_out/${tag}_ncu_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/${tag}_launches_raw.csv "ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu $*" > gpurun_out/${tag}_launches.csv
cat gpurun_out/${tag}_launches.csv
