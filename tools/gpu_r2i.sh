#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 22 --warmup 3 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
head -c 900 gpurun_out/r2i_bench_2gpu.json; tail -3 gpurun_out/r2i_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2i_ref_2gpu.json 2> gpurun_out/r2i_ref_2gpu.err; head -c 300 gpurun_out/r2i_ref_2gpu.json
