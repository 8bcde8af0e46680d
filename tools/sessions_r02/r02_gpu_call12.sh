#!/usr/bin/env bash
# round 2, GPU call 12: event-driven policies at higher replica counts (they are latency bound per replica)
set -u
mkdir -p gpurun_out
for cfg in "sjf 10000 2368" "sjf 10000 4144" "dlas-gpu 100000 2368" "dlas-gpu 100000 4144" "gittins 100000 1184" "gittins 100000 2368"; do
  set -- $cfg
  timeout 400 python bench.py --policy $1 --jobs $2 --replicas $3 --steps 1 --warmup 1 > gpurun_out/r02_c12_$1_$3.json 2> gpurun_out/r02_c12_$1_$3.err
  echo "$cfg: $(python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_c12_$1_$3.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms','replicas')}, 'frac', d['roofline']['frac'])" 2>&1 | tail -1)"; tail -1 gpurun_out/r02_c12_$1_$3.err | cut -c1-200
done
