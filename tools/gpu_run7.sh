python tools/tc_check.py | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/conv_bench.py 64 216 3 1024 1920 2 5 1 1 1 0
python tools/conv_bench.py 64 3 3 1024 1920 2 5 1 1 0 3
python tools/conv_bench.py 16 2 7 1024 1920 2 5 1 1 0 0
python tools/conv_bench.py 64 64 3 1024 1920 2 5 1 1 0 2 1
python tools/conv_bench.py 64 64 3 1024 1920 2 5 1 1 0 1 0
python bench.py --steps 11 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -2 gpurun_out/bench_r1g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1g.json').read().strip().splitlines()[-1])
ks=d.pop('kernels')
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in list(ks.items())[:40]: print(k,v)
PY
