set -x
timeout 300 python tools/tc_check.py 2>&1 | tail -20
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8
python tools/conv_bench.py 64 64 3 1024 1920 2 10
python tools/conv_bench.py 128 128 3 512 960 2 10
python tools/conv_bench.py 128 128 1 512 960 2 10
python tools/conv_bench.py 32 64 7 1024 1920 2 10
python tools/conv_bench.py 64 128 3 1024 1920 2 10 1 2
python bench.py --steps 11 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
tail -3 gpurun_out/bench_r1e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1e.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['stats'])
for k,v in list(d['kernels'].items())[:24]: print(k,v)
PY
