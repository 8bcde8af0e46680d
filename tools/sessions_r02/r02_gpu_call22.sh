#!/usr/bin/env bash
# round 2, GPU call 22: smoke, the whole GPU suite and the bench line on the final tree (grouped sjf admission, direct
# gittins table, exchange on request); the reference arm has not changed since call 16
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c22_smoke.txt 2>&1; tail -2 gpurun_out/r02_c22_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -25 > gpurun_out/r02_c22_tests.txt; tail -3 gpurun_out/r02_c22_tests.txt
SECONDS=0
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_c22_bench.json 2> gpurun_out/r02_c22_bench.err
echo "bench wall ${SECONDS}s"; grep "^{" gpurun_out/r02_c22_bench.json | tail -1 | cut -c1-600; tail -3 gpurun_out/r02_c22_bench.err | cut -c1-300
