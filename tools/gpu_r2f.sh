#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=8 2>&1 | tail -80 > gpurun_out/r2f_tests.log
grep -E "passed|failed|Error|assert " gpurun_out/r2f_tests.log | tail -30
