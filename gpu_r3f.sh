#!/bin/bash
mkdir -p gpurun_out
python tools/ar_bench.py 8 3 2>&1 | grep -v Warn
python tools/ar_bench.py 16 3 2>&1 | grep -v Warn
timeout 900 python -m pytest tests/test_coding.py -x -q -m gpu 2>&1 | tail -15
