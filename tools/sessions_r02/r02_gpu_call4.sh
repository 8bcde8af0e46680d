#!/usr/bin/env bash
# round 2, GPU call 4: parity of the V4 kernel (32-bit shared addressing, 32-bit masks) + arenas / strided copies,
# BASELINE configs at size, V3/V4 register variants, end-to-end path with one strided copy each way
set -u
mkdir -p gpurun_out
cp gpuschedule_b200/libgsched.so /tmp/libgsched_default.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --tb=short 2>&1 | tail -30 > gpurun_out/r02_c4_parity.txt
tail -5 gpurun_out/r02_c4_parity.txt
if grep -q "failed\|error" gpurun_out/r02_c4_parity.txt; then echo "PARITY FAILED"; exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_parity.py --durations=8 2>&1 | tail -30 > gpurun_out/r02_c4_rest.txt
tail -14 gpurun_out/r02_c4_rest.txt
for v in mb32 v4_mb32 v4_mb28 v4_mb24; do
  cp tools/variants/libgsched_$v.so gpuschedule_b200/libgsched.so
  case $v in *mb32) R=4736;; *mb28) R=4144;; *) R=3552;; esac
  timeout 200 python bench.py --replicas $R --steps 3 --warmup 3 --value-only --distinct 148 > gpurun_out/r02_c4_value_$v.json 2> gpurun_out/r02_c4_value_$v.err
  echo "$v: $(cat gpurun_out/r02_c4_value_$v.json)"
done
cp /tmp/libgsched_default.so gpuschedule_b200/libgsched.so
i=0
for extra in "" "--e2e-stagger 0" "--e2e-threads 8" "--e2e-threads 4"; do
  i=$((i+1))
  timeout 300 python bench.py --replicas 4736 --distinct 296 --steps 2 --warmup 3 --e2e-only --e2e-steps 4 $extra > gpurun_out/r02_c4_e2e_$i.json 2> gpurun_out/r02_c4_e2e_$i.err
  echo "e2e [$extra]: $(cut -c1-900 gpurun_out/r02_c4_e2e_$i.json)"; tail -2 gpurun_out/r02_c4_e2e_$i.err
done
