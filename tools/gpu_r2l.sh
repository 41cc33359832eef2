#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=8 2>&1 | tail -40 > gpurun_out/r2l_tests.log
grep -E "passed|failed|Error|assert " gpurun_out/r2l_tests.log | tail -20
timeout 900 python bench.py --steps 22 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
python -c "
import json; p=json.load(open('gpurun_out/r2l_bench.json')); print(p['ms_per_step'], 1000/p['e2e']['value'], p['exact_precision_ms_per_step'])
for k,v in p['memory_bound_kernels'].items():
    if 'spynet' in k or 'se_' in k: print(k, v)"
