#!/usr/bin/env bash
# round 2, GPU call 3: parity of the warp-uniform-store version + new entry points, register/occupancy variants,
# end-to-end path diagnostics
set -u
mkdir -p gpurun_out
cp gpuschedule_b200/libgsched.so /tmp/libgsched_default.so
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_switch.py -x -q --tb=short 2>&1 | tail -30 > gpurun_out/r02_c3_parity.txt
tail -5 gpurun_out/r02_c3_parity.txt
if grep -q "failed\|error" gpurun_out/r02_c3_parity.txt; then echo "PARITY FAILED"; exit 0; fi
for mb in 32 28 24 20; do
  cp tools/variants/libgsched_mb$mb.so gpuschedule_b200/libgsched.so
  R=$((148*mb))
  timeout 200 python bench.py --replicas $R --steps 3 --warmup 3 --value-only --distinct 148 > gpurun_out/r02_c3_value_mb$mb.json 2> gpurun_out/r02_c3_value_mb$mb.err
  echo "mb$mb: $(cat gpurun_out/r02_c3_value_mb$mb.json)"
done
cp /tmp/libgsched_default.so gpuschedule_b200/libgsched.so
i=0
for extra in "" "--no-numa" "--e2e-threads 32" "--e2e-stagger 0" "--e2e-threads 8"; do
  i=$((i+1))
  timeout 300 python bench.py --replicas 4736 --distinct 296 --steps 2 --warmup 3 --e2e-only --e2e-steps 3 $extra > gpurun_out/r02_c3_e2e_$i.json 2> gpurun_out/r02_c3_e2e_$i.err
  echo "e2e [$extra]: $(cut -c1-900 gpurun_out/r02_c3_e2e_$i.json)"
done
