#!/usr/bin/env bash
# round 2, 2-GPU call: the sharded single simulation (gittins, config C4) -- bit identity test, then timings of the
# BASELINE trace (runnable list of a few dozen jobs) and of overloaded traces (runnable lists of thousands)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_c5_topo2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q --tb=short 2>&1 | tail -20 > gpurun_out/r02_c5_sharded_test.txt
tail -6 gpurun_out/r02_c5_sharded_test.txt
if grep -q "failed\|error" gpurun_out/r02_c5_sharded_test.txt; then echo "SHARDED TEST FAILED"; exit 0; fi
i=0
for cfg in "--sharded-jobs 100000 --sharded-rate 0.5" "--sharded-jobs 100000 --sharded-rate 3.0" "--sharded-jobs 30000 --sharded-rate 20.0"; do
  i=$((i+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
      bench.py --gpus 2 --only-sharded $cfg > gpurun_out/r02_c5_sharded_$i.json 2> gpurun_out/r02_c5_sharded_$i.err
  echo "sharded [$cfg]: $(cat gpurun_out/r02_c5_sharded_$i.json)"; tail -3 gpurun_out/r02_c5_sharded_$i.err
done
