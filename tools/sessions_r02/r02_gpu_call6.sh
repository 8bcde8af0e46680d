#!/usr/bin/env bash
# round 2, GPU call 6: the bench line of the shipped tree (both arms), ncu evidence at the shipped replica count,
# launch list of the same command, config c5
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --tb=short 2>&1 | tail -8 > gpurun_out/r02_c6_parity.txt
tail -3 gpurun_out/r02_c6_parity.txt
if grep -q "failed\|error" gpurun_out/r02_c6_parity.txt; then echo "PARITY FAILED"; exit 0; fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_c6_bench.json 2> gpurun_out/r02_c6_bench.err
tail -c 700 gpurun_out/r02_c6_bench.json; echo; tail -3 gpurun_out/r02_c6_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c6_reference.json 2> gpurun_out/r02_c6_reference.err
cut -c1-400 gpurun_out/r02_c6_reference.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gs_tick2 -s 3 -c 1 -f -o gpurun_out/r02_tick2_v4 \
    python bench.py --steps 1 --warmup 3 --value-only > gpurun_out/r02_c6_ncu.log 2>&1
tail -2 gpurun_out/r02_c6_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c6_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-sharded --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02_c6_launchrun.log 2>&1
tail -2 gpurun_out/r02_c6_launchrun.log | cut -c1-300
timeout 900 python bench.py --config c5 --steps 2 --warmup 3 --no-extras > gpurun_out/r02_c6_bench_c5.json 2> gpurun_out/r02_c6_bench_c5.err
cut -c1-1500 gpurun_out/r02_c6_bench_c5.json; tail -3 gpurun_out/r02_c6_bench_c5.err
