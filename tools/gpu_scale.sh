#!/bin/bash
# bench.py on N GPUs of one box (N = number of visible GPUs): python -m torch.distributed.run, one rank per GPU
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 22 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
head -c 300 gpurun_out/r02_bench_${N}gpu.json; echo; tail -2 gpurun_out/r02_bench_${N}gpu.err
