#!/usr/bin/env bash
# round 2, GPU call 1: whole GPU suite (incl. the horus+ / coop / CLI tests that never ran), ncu baseline of the
# round-1 tick kernel at the shipped 3552 replicas, one ncu capture of the horus kernels
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_smi.txt 2>&1
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
lscpu | head -30 > gpurun_out/r02_lscpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -rxXs 2>&1 | tail -40 > gpurun_out/r02_gpu_tests_a.txt
tail -5 gpurun_out/r02_gpu_tests_a.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gs_tick -s 3 -c 1 -f -o gpurun_out/r02_tick_base \
    python bench.py --replicas 3552 --steps 1 --warmup 3 --value-only > gpurun_out/r02_ncu_tick_base.log 2>&1
tail -2 gpurun_out/r02_ncu_tick_base.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gs_horus -c 2 -f -o gpurun_out/r02_horus_base \
    python bench.py --mode horus --horus-replicas 1184 > gpurun_out/r02_ncu_horus_base.log 2>&1
tail -c 600 gpurun_out/r02_ncu_horus_base.log
