#!/usr/bin/env bash
# round 2, fourth 2-GPU call: the C4 block alone after making its collectives failure-safe
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
    bench.py --gpus 2 --only-sharded > gpurun_out/r02_c25_sharded.json 2> gpurun_out/r02_c25_sharded.err
grep "^{" gpurun_out/r02_c25_sharded.json | tail -1 | cut -c1-2500; tail -3 gpurun_out/r02_c25_sharded.err | cut -c1-300
