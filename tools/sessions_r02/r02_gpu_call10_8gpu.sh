#!/usr/bin/env bash
# round 2, 8-GPU call: end-to-end path at N=8 (the part of the scaling run that depends on the host), short
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_c10_topo8.txt 2>&1
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29575 \
    bench.py --gpus 8 --steps 2 --warmup 3 --e2e-only --distinct 296 > gpurun_out/r02_c10_e2e_8gpu.json 2> gpurun_out/r02_c10_e2e_8gpu.err
echo "wall ${SECONDS}s"
grep "^{" gpurun_out/r02_c10_e2e_8gpu.json | tail -1 | cut -c1-1500; tail -3 gpurun_out/r02_c10_e2e_8gpu.err | cut -c1-300
