#!/bin/bash
# Shorter variant of gpu_check.sh: GPU parity tests, smoke, a bench line (both arms), ncu launch list + one full capture of k_frontend.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 700 python -m pytest tests -q -m gpu --timeout=300 > gpurun_out/pytest_gpu_full.log 2>&1; grep -E '^E  .*Error|^FAILED|passed|failed' gpurun_out/pytest_gpu_full.log | tail -40
echo "== smoke" ; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 400 python bench.py --steps 20 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.json
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --log2n 26 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full capture of k_frontend"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frontend -s 3 -c 2 -o gpurun_out/prof_frontend -f \
    python bench.py --steps 2 --warmup 3 --log2n 26 --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "== reference arm" ; timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.json
ls -la gpurun_out
