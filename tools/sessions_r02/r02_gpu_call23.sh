#!/usr/bin/env bash
# round 2, GPU call 23: event-driven policy kernels at 16 / 20 / 24 resident warps per SM (launch bounds), replicas = 148 x warps
set -u
mkdir -p gpurun_out
cp gpuschedule_b200/libgsched.so /tmp/libgsched_keep.so
for mb in 16 20 24; do
  if [ $mb != 16 ]; then cp tools/variants/libgsched_pm$mb.so gpuschedule_b200/libgsched.so; fi
  R=$((148 * mb))
  for cfg in "sjf 10000" "dlas-gpu 100000" "gittins 100000"; do
    set -- $cfg
    timeout 400 python bench.py --policy $1 --jobs $2 --replicas $R --steps 1 --warmup 1 > gpurun_out/r02_c23_$1_pm$mb.json 2> gpurun_out/r02_c23_$1_pm$mb.err
    echo "minblocks $mb $cfg x $R: $(python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_c23_$1_pm$mb.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms','replicas')}, 'frac', d['roofline']['frac'])" 2>&1 | tail -1)"
  done
done
cp /tmp/libgsched_keep.so gpuschedule_b200/libgsched.so
