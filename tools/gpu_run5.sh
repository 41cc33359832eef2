set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 22 --warmup 3 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -2 gpurun_out/bench_r1f.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1f.json 2> gpurun_out/bench_ref_r1f.err; tail -2 gpurun_out/bench_ref_r1f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1f.json').read().strip().splitlines()[-1])
ks=d.pop('kernels')
print(json.dumps(d))
for k,v in list(ks.items())[:16]: print(k,v)
print(open('gpurun_out/bench_ref_r1f.json').read())
PY
