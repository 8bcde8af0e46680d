#!/usr/bin/env bash
# round 2, GPU call 21: ncu capture of the horus kernel (one wave of replicas, one simulation per warp)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gs_horus_kernel -c 1 -f -o gpurun_out/r02_horus_kernel \
    python bench.py --mode horus --horus-replicas 2368 --horus-scalar-only > gpurun_out/r02_c21_ncu.log 2>&1
tail -3 gpurun_out/r02_c21_ncu.log | cut -c1-300
ls -la gpurun_out/r02_horus_kernel.ncu-rep
