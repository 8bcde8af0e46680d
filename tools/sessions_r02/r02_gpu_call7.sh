#!/usr/bin/env bash
# round 2, GPU call 7: end-to-end pipeline with two handles per thread, config c5, horus mapping experiment
set -u
mkdir -p gpurun_out
i=0
for extra in "" "--e2e-stagger 0" "--e2e-threads 8" "--e2e-steps 8"; do
  i=$((i+1))
  timeout 300 python bench.py --distinct 296 --steps 2 --warmup 3 --e2e-only $extra > gpurun_out/r02_c7_e2e_$i.json 2> gpurun_out/r02_c7_e2e_$i.err
  echo "e2e [$extra]: $(cut -c1-1100 gpurun_out/r02_c7_e2e_$i.json)"; tail -2 gpurun_out/r02_c7_e2e_$i.err | cut -c1-300
done
timeout 900 python bench.py --config c5 --steps 2 --warmup 3 --no-extras > gpurun_out/r02_c7_bench_c5.json 2> gpurun_out/r02_c7_bench_c5.err
cut -c1-1800 gpurun_out/r02_c7_bench_c5.json; tail -3 gpurun_out/r02_c7_bench_c5.err | cut -c1-300
timeout 600 python bench.py --mode horus --horus-replicas 9472 --horus-both-mappings > gpurun_out/r02_c7_horus_9472.json 2> gpurun_out/r02_c7_horus_9472.err
cut -c1-900 gpurun_out/r02_c7_horus_9472.json; tail -3 gpurun_out/r02_c7_horus_9472.err | cut -c1-300
