#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=8 2>&1 | tail -60 > gpurun_out/r2j_tests.log
grep -E "passed|failed|Error|assert " gpurun_out/r2j_tests.log | tail -20
B="python bench.py --steps 22 --warmup 3 --no-cpu-baseline --no-eager-baseline"
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export TDVC_B200_NO_OVERLAP_CACHE=1; else unset TDVC_B200_NO_OVERLAP_CACHE; fi
  timeout 600 $B > gpurun_out/r2j_bench_$v.json 2> gpurun_out/r2j_bench.err
  python -c "
import json; p=json.load(open('gpurun_out/r2j_bench_$v.json')); print('no_overlap_cache=$v', p['ms_per_step'], 1000/p['e2e']['value'], p['exact_precision_ms_per_step'], p['clocks']['sm_mhz'])"
done
