#!/usr/bin/env bash
# round 2, GPU call 13: what the driver runs at round end, on the final tree: smoke, the GPU suite, both bench arms
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c13_smoke.txt 2>&1; tail -2 gpurun_out/r02_c13_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -6 > gpurun_out/r02_c13_tests.txt; tail -3 gpurun_out/r02_c13_tests.txt
SECONDS=0
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_c13_reference.json 2> gpurun_out/r02_c13_reference.err
echo "reference arm wall ${SECONDS}s: $(cut -c1-300 gpurun_out/r02_c13_reference.json)"
SECONDS=0
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_c13_bench.json 2> gpurun_out/r02_c13_bench.err
echo "bench wall ${SECONDS}s"; grep "^{" gpurun_out/r02_c13_bench.json | tail -1 | cut -c1-600; tail -3 gpurun_out/r02_c13_bench.err | cut -c1-300
