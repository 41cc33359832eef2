#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --maxfail=10 2>&1 | tail -30 > gpurun_out/r2o_kernels.log
grep -E "passed|failed|^FAILED|Error" gpurun_out/r2o_kernels.log | tail -20
for ldg in 1 0; do
  if [ $ldg = 1 ]; then export TDVC_B200_CONV_LDG=1; else unset TDVC_B200_CONV_LDG; fi
  echo "== LDG=$ldg"
  for args in "64 64 3 1024 1920 2 10" "64 64 3 1024 1920 2 10 1 1 0 1 1" "64 64 3 1024 1920 2 10 1 1 0 1 0 0 1" "128 128 3 512 960 2 10" "128 128 3 512 960 2 10 1 1 0 1 0 0 1" "128 64 3 1024 1920 2 10" "4 64 3 1024 1920 2 10" "64 32 7 1024 1920 2 10" "32 64 7 1024 1920 2 10" "64 216 3 1024 1920 2 10 1 1 1"; do
    timeout 120 python tools/conv_bench.py $args 2>&1 | tail -1
  done
done
