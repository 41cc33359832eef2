#!/bin/bash
for lib in "" "$PWD/tdvc_b200/libtdvc_b200_ab.so" ""  "$PWD/tdvc_b200/libtdvc_b200_ab.so"; do
  echo "== lib: ${lib:-default}"
  for args in "256 64 1 1024 1920 2 10" "192 64 1 1024 1920 2 10" "128 64 1 1024 1920 2 10" "128 128 1 512 960 2 10 1 1 0 0 1 1" "128 128 1 256 480 2 10 1 1 0 0 1 2"; do
    TDVC_B200_LIB=$lib python tools/conv_bench.py $args 2>&1 | tail -1
  done
done
