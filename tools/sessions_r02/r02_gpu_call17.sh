#!/usr/bin/env bash
# round 2, GPU call 17: sjf with grouped admission (runs of identical list entries placed with one node walk): parity, then throughput
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_policies.py tests/test_gpu_configs.py -x -q --tb=short -k "sjf or mixed or invariants or cli or fixtures" 2>&1 | tail -25 > gpurun_out/r02_c17_tests.txt; tail -4 gpurun_out/r02_c17_tests.txt
if grep -q "failed\|error" gpurun_out/r02_c17_tests.txt; then echo "TESTS FAILED"; exit 0; fi
for cfg in "sjf 10000 2368" "sjf 10000 3108"; do
  set -- $cfg
  timeout 400 python bench.py --policy $1 --jobs $2 --replicas $3 --steps 1 --warmup 1 > gpurun_out/r02_c17_$1_$3.json 2> gpurun_out/r02_c17_$1_$3.err
  echo "$cfg: $(python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_c17_$1_$3.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms','replicas')}, 'frac', d['roofline']['frac'])" 2>&1 | tail -1)"; tail -1 gpurun_out/r02_c17_$1_$3.err | cut -c1-200
done
