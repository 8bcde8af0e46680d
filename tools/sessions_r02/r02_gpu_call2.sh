#!/usr/bin/env bash
# round 2, GPU call 2: parity of the event-stepped fifo engine (gs_tick2_kernel), then first numbers + ncu
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --tb=short 2>&1 | tail -40 > gpurun_out/r02_c2_parity.txt
tail -6 gpurun_out/r02_c2_parity.txt
if grep -q "failed\|error" gpurun_out/r02_c2_parity.txt; then
  echo "PARITY FAILED -- skipping the measurements"; exit 0
fi
timeout 600 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_parity.py 2>&1 | tail -15 > gpurun_out/r02_c2_rest.txt
tail -4 gpurun_out/r02_c2_rest.txt
for R in 4736 3552 2368; do
  timeout 300 python bench.py --replicas $R --steps 3 --warmup 3 --value-only > gpurun_out/r02_c2_value_R$R.json 2> gpurun_out/r02_c2_value_R$R.err
  cat gpurun_out/r02_c2_value_R$R.json
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c2_bench.json 2> gpurun_out/r02_c2_bench.err
tail -c 1500 gpurun_out/r02_c2_bench.json; tail -5 gpurun_out/r02_c2_bench.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gs_tick2 -s 3 -c 1 -f -o gpurun_out/r02_tick2_v1 \
    python bench.py --replicas 4736 --steps 1 --warmup 3 --value-only > gpurun_out/r02_c2_ncu.log 2>&1
tail -2 gpurun_out/r02_c2_ncu.log
