#!/usr/bin/env bash
# round 2, third 2-GPU call (final tree: 24-byte records, direct gittins look-up): sharded test with the adaptive exchange, then the complete bench line at N=2 exactly as the
# driver launches it (replica throughput, e2e, sharded block incl. always / adaptive exchange on both traces)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q --tb=short 2>&1 | tail -12 > gpurun_out/r02_c20_sharded_test.txt
tail -4 gpurun_out/r02_c20_sharded_test.txt
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_c20_bench_2gpu.json 2> gpurun_out/r02_c20_bench_2gpu.err
echo "bench --gpus 2 wall: ${SECONDS}s"
grep "^{" gpurun_out/r02_c20_bench_2gpu.json | tail -1 | cut -c1-3000; tail -4 gpurun_out/r02_c20_bench_2gpu.err | cut -c1-300
