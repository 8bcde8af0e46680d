#!/usr/bin/env bash
# round 2, GPU call 24: policy tests on the shipped build (20 resident warps per SM)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_policies.py tests/test_gpu_configs.py -x -q --tb=short 2>&1 | tail -8 > gpurun_out/r02_c24_tests.txt; tail -3 gpurun_out/r02_c24_tests.txt
