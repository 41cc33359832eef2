# Round validation on a GPU box (run under gpurun): GPU test suite, smoke, both bench arms.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
for k in ('kernels', 'memory_bound_kernels'):
    d.pop(k, None)
print(json.dumps(d))
print(open('gpurun_out/bench_ref.json').read())
PY
