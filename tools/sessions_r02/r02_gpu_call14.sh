#!/usr/bin/env bash
# round 2, GPU call 14: 24-byte records + node events: smoke, whole GPU suite, value, end to end
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c14_smoke.txt 2>&1; tail -2 gpurun_out/r02_c14_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -25 > gpurun_out/r02_c14_tests.txt; tail -4 gpurun_out/r02_c14_tests.txt
if grep -q "failed\|error" gpurun_out/r02_c14_tests.txt; then echo "TESTS FAILED"; exit 0; fi
timeout 200 python bench.py --steps 3 --warmup 3 --value-only --distinct 148 > gpurun_out/r02_c14_value.json 2> gpurun_out/r02_c14_value.err
echo "value: $(cat gpurun_out/r02_c14_value.json)"
timeout 300 python bench.py --distinct 296 --steps 2 --warmup 3 --e2e-only > gpurun_out/r02_c14_e2e.json 2> gpurun_out/r02_c14_e2e.err
echo "e2e: $(cut -c1-800 gpurun_out/r02_c14_e2e.json)"; tail -2 gpurun_out/r02_c14_e2e.err | cut -c1-300
