#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_coding.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r3c_coding.log
cat gpurun_out/r3c_coding.log
./gpu_r3b.sh
