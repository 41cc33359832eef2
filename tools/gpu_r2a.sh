#!/bin/bash
# round-2 GPU validation: kernel tests, parity tests (full-size fixtures included), bench in both precision modes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py -q -m gpu --maxfail=8 -x 2>&1 | tail -40 > gpurun_out/r2a_kernels.log
timeout 2400 python -m pytest tests/test_gpu_parity.py -q -m gpu --maxfail=6 -s 2>&1 | tail -120 > gpurun_out/r2a_parity.log
timeout 900 python bench.py --steps 22 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 900 python bench.py --steps 22 --warmup 3 --no-cpu-baseline --no-eager-baseline --precision mixed > gpurun_out/r2a_bench_mixed.json 2> gpurun_out/r2a_bench_mixed.err
tail -5 gpurun_out/r2a_kernels.log gpurun_out/r2a_parity.log
head -c 1500 gpurun_out/r2a_bench.json; echo; head -c 600 gpurun_out/r2a_bench_mixed.json; echo
tail -5 gpurun_out/r2a_bench.err gpurun_out/r2a_bench_mixed.err
