#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=8 -s 2>&1 | tail -150 > gpurun_out/r2d_tests.log
timeout 900 python bench.py --steps 22 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
timeout 900 python bench.py --steps 22 --warmup 3 --no-cpu-baseline --no-eager-baseline --precision mixed > gpurun_out/r2d_bench_mixed.json 2> gpurun_out/r2d_bench_mixed.err
timeout 600 python tools/sync_ab.py exact > gpurun_out/r2d_sync_ab.log 2>&1
grep -E "passed|failed|identical|recon max|frame [0-9]:" gpurun_out/r2d_tests.log | tail -30
head -c 300 gpurun_out/r2d_bench.json; echo; head -c 300 gpurun_out/r2d_bench_mixed.json; echo; cat gpurun_out/r2d_sync_ab.log
