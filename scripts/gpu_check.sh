#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a bench line, ncu launch list + one full capture of the top kernel.
# usage: scripts/gpu_check.sh [quick|full]
mode=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu --timeout=600 > gpurun_out/pytest_gpu_full.log 2>&1; grep -E '^E  .*Error|^FAILED|passed|failed' gpurun_out/pytest_gpu_full.log | tail -40
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.json
if [ "$mode" = "full" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --log2n 26 --no-cpu > gpurun_out/ncu_bench.log 2>&1
  echo "== ncu full capture of k_frontend"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_frontend -s 3 -c 2 -o gpurun_out/prof_frontend -f \
      python bench.py --steps 2 --warmup 3 --log2n 26 --no-cpu > gpurun_out/ncu_full.log 2>&1
  echo "== ncu full capture of the back-end kernels"
  for k in k_agc_chain k_be_emit k_be_prep k_dc_local; do
    timeout 300 ncu --set full --clock-control none -k regex:$k -s 4 -c 1 -o gpurun_out/prof_$k -f \
        python bench.py --steps 2 --warmup 3 --log2n 26 --no-cpu > gpurun_out/ncu_$k.log 2>&1
  done
  echo "== ncu full capture of the channelizer kernels (configs 3 and 4)"
  timeout 300 ncu --set full --clock-control none -k regex:k_pfb_tile -s 2 -c 1 -o gpurun_out/prof_k_pfb_tile -f \
      python scripts/prof_c3.py 16 agc > gpurun_out/ncu_k_pfb_tile.log 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:k_pfb_ring -s 2 -c 1 -o gpurun_out/prof_k_pfb_ring -f \
      python scripts/prof_c3.py 1024 agc > gpurun_out/ncu_k_pfb_ring.log 2>&1
  ls -la gpurun_out
fi
