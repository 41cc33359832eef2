set -x
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -8
python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -8
./tools/mma_probe 100 2>&1 | tee gpurun_out/mma_probe.log
python tools/stage_parity.py 128 192 2 0 2>&1 | grep -E "mcnet|prediction1|recon|bpp|y_hat"
python tools/conv_bench.py 64 64 3 1024 1920 2 10
python tools/conv_bench.py 128 128 3 512 960 2 10
python bench.py --steps 11 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
tail -3 gpurun_out/bench_r1d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1d.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for k,v in list(d['kernels'].items())[:14]: print(k,v)
PY
