#!/usr/bin/env bash
# round 2, GPU call 11: compact spans / start-only job results + blocking waits: whole GPU suite, e2e, horus register variants
set -u
mkdir -p gpurun_out
cp gpuschedule_b200/libgsched.so /tmp/libgsched_default.so
timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -25 > gpurun_out/r02_c11_tests.txt
tail -5 gpurun_out/r02_c11_tests.txt
if grep -q "failed\|error" gpurun_out/r02_c11_tests.txt; then echo "TESTS FAILED"; exit 0; fi
timeout 200 python bench.py --steps 3 --warmup 3 --value-only --distinct 148 > gpurun_out/r02_c11_value.json 2> gpurun_out/r02_c11_value.err
echo "value: $(cat gpurun_out/r02_c11_value.json)"
for extra in "" "--e2e-steps 8"; do
  timeout 300 python bench.py --distinct 296 --steps 2 --warmup 3 --e2e-only $extra > gpurun_out/r02_c11_e2e.json 2> gpurun_out/r02_c11_e2e.err
  echo "e2e [$extra]: $(cut -c1-700 gpurun_out/r02_c11_e2e.json)"; tail -2 gpurun_out/r02_c11_e2e.err | cut -c1-300
done
for v in h16 h21 h25; do
  cp tools/variants/libgsched_$v.so gpuschedule_b200/libgsched.so
  timeout 300 python bench.py --mode horus --horus-replicas 9472 --horus-scalar-only > gpurun_out/r02_c11_horus_$v.json 2> gpurun_out/r02_c11_horus_$v.err
  echo "horus $v: $(cat gpurun_out/r02_c11_horus_$v.json)"; tail -1 gpurun_out/r02_c11_horus_$v.err | cut -c1-200
done
cp /tmp/libgsched_default.so gpuschedule_b200/libgsched.so
