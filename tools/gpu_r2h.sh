#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=8 2>&1 | tail -60 > gpurun_out/r2h_tests.log
grep -E "passed|failed|Error|assert " gpurun_out/r2h_tests.log | tail -20
timeout 900 python bench.py --steps 22 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
head -c 400 gpurun_out/r2h_bench.json
