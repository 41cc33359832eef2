#!/bin/bash
for lib in "$PWD/tdvc_b200/libtdvc_b200_ab.so" ""; do
  echo "== lib: ${lib:-new}"
  for sg in 2.0 0.3 6.0; do TDVC_B200_LIB=$lib python tools/dcn_bench.py 1024 1920 $sg; done
done
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k dcn 2>&1 | tail -3
