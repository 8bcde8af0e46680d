#!/usr/bin/env bash
# round 2, GPU call 16: what the driver runs at round end (smoke, the GPU suite, both bench arms) on the tree with the
# 24-byte records and the faster command line, plus the ncu capture of the tick kernel at the shipped replica count
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c16_smoke.txt 2>&1; tail -2 gpurun_out/r02_c16_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -25 > gpurun_out/r02_c16_tests.txt; tail -3 gpurun_out/r02_c16_tests.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gs_tick2 -s 3 -c 1 -f -o gpurun_out/r02_tick2_v5 \
    python bench.py --steps 1 --warmup 3 --value-only > gpurun_out/r02_c16_ncu.log 2>&1
tail -2 gpurun_out/r02_c16_ncu.log | cut -c1-300
SECONDS=0
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_c16_reference.json 2> gpurun_out/r02_c16_reference.err
echo "reference arm wall ${SECONDS}s: $(cut -c1-300 gpurun_out/r02_c16_reference.json)"
SECONDS=0
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_c16_bench.json 2> gpurun_out/r02_c16_bench.err
echo "bench wall ${SECONDS}s"; grep "^{" gpurun_out/r02_c16_bench.json | tail -1 | cut -c1-600; tail -3 gpurun_out/r02_c16_bench.err | cut -c1-300
