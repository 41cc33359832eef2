#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/fullres_diag.py > gpurun_out/r2b_diag.log 2>&1
timeout 600 python bench.py --steps 22 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 600 python bench.py --steps 22 --warmup 3 --no-cpu-baseline --no-eager-baseline --precision mixed > gpurun_out/r2b_bench_mixed.json 2> gpurun_out/r2b_bench_mixed.err
cat gpurun_out/r2b_diag.log | tail -70
head -c 400 gpurun_out/r2b_bench.json; echo; head -c 400 gpurun_out/r2b_bench_mixed.json; echo
