#!/usr/bin/env bash
# Host build of gs_horus.cu (stand-in CUDA runtime) under AddressSanitizer and UBSan: bounds of the per-replica slab,
# alignment of every record, signed overflow.  Not part of the suite (needs LD_PRELOAD of the sanitizer runtime).
#   bash tests/emu/run_sanitizers.sh
set -eu
cd "$(dirname "$0")/../.."
for san in address undefined; do
  lib=/tmp/libhorus_abi_$san.so
  g++ -O1 -g -fsanitize=$san -fno-sanitize-recover=all -fno-omit-frame-pointer -fPIC -std=c++17 -ffp-contract=off -shared -x c++ \
      -I tests/emu/fake_cuda -I include -I gpuschedule_b200/csrc -o $lib gpuschedule_b200/csrc/gs_horus.cu
  rt=$(g++ -print-file-name=lib$([ $san = address ] && echo asan || echo ubsan).so)
  echo "== $san"
  LD_PRELOAD=$rt ASAN_OPTIONS=detect_leaks=0 python tests/emu/sanitizer_scenarios.py $lib
done
