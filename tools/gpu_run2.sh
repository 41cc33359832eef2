python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -5
python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8
python tools/stage_parity.py 128 192 2 0 2>&1 | tail -40
python tools/conv_bench.py 64 64 3 1024 1920 2 10
python bench.py --steps 11 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
