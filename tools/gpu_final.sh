#!/bin/bash
# final record of the round: default bench line (with CPU and GPU-eager baselines) and the reference arm
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
head -c 400 gpurun_out/r02_bench_1gpu.json; echo; head -c 300 gpurun_out/r02_bench_reference_arm.json; echo
