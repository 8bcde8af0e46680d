#!/usr/bin/env bash
# round 2, GPU call 15: horus replica batches with 32 simulations per warp at 8 and 16 resident warps per SM
set -u
mkdir -p gpurun_out
for R in 37888 75776; do
  SECONDS=0
  timeout 600 python bench.py --mode horus --horus-replicas $R --horus-rows 12288 --horus-stream 6000000 --horus-both-mappings --horus-scalar-only \
      > gpurun_out/r02_c15_horus_$R.json 2> gpurun_out/r02_c15_horus_$R.err
  echo "horus $R (${SECONDS}s): $(cat gpurun_out/r02_c15_horus_$R.json)"; tail -2 gpurun_out/r02_c15_horus_$R.err | cut -c1-300
done
