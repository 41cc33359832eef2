#!/bin/bash
mkdir -p gpurun_out
python tools/ar_bench.py 8 2 2>&1 | grep -v Warn


ls -la gpurun_out/
