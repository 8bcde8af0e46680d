#!/usr/bin/env bash
# round 2, GPU call 8: V5 kernel (early wheel loads, take_lowest fast path) parity + value, config c5, host memory
set -u
mkdir -p gpurun_out
free -g > gpurun_out/r02_c8_free.txt 2>&1; cat gpurun_out/r02_c8_free.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q --tb=short 2>&1 | tail -8 > gpurun_out/r02_c8_parity.txt
tail -3 gpurun_out/r02_c8_parity.txt
if grep -q "failed\|error" gpurun_out/r02_c8_parity.txt; then echo "PARITY FAILED"; exit 0; fi
for R in 4144 3552; do
  timeout 200 python bench.py --replicas $R --steps 3 --warmup 3 --value-only --distinct 148 > gpurun_out/r02_c8_value_v5_R$R.json 2> gpurun_out/r02_c8_value_v5_R$R.err
  echo "v5 R=$R: $(cat gpurun_out/r02_c8_value_v5_R$R.json)"
done
timeout 900 python bench.py --config c5 --steps 2 --warmup 3 --no-extras > gpurun_out/r02_c8_bench_c5.json 2> gpurun_out/r02_c8_bench_c5.err
cut -c1-2200 gpurun_out/r02_c8_bench_c5.json; tail -3 gpurun_out/r02_c8_bench_c5.err | cut -c1-300
