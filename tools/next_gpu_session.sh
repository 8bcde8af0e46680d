#!/usr/bin/env bash
# First GPU call of the next round (run under gpurun, one B200):
#   gpurun --timeout 900 -- 'bash tools/next_gpu_session.sh'
# 1. the whole GPU suite -- the horus+ / word-stream / batched-sweep / CLI tests of tests/test_gpu_widen_horus.py have
#    never run on a device (XPASS = they work; then drop the NOT_RUN_YET marks)
# 2. the utilisation-aware engine: all three kernel mappings (scalar x1 / x32 lanes, cooperative warp) + the horus+ device check, then a launch list and one
#    full ncu capture of gs_horus_kernel (where do the ~50 dependent loads per sample go?)
# 3. the main line again (nothing on that path changed since profiles/r01_bench_final.json)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -25 | tee gpurun_out/r02_gpu_tests.txt
timeout 200 python bench.py --mode horus --horus-replicas 2368 --horus-both-mappings > gpurun_out/r02_bench_horus.json 2> gpurun_out/r02_bench_horus.err
tail -c 1500 gpurun_out/r02_bench_horus.json
timeout 300 ncu --set full --clock-control none --import-source on -k gs_horus_kernel -c 1 -o gpurun_out/r02_horus_kernel \
    python bench.py --mode horus --horus-replicas 592 > gpurun_out/r02_ncu_horus.log 2>&1
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_main.json 2> gpurun_out/r02_bench_main.err
tail -c 600 gpurun_out/r02_bench_main.json
