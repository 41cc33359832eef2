#!/bin/bash
B="python bench.py --steps 22 --no-cpu-baseline --no-eager-baseline"
for w in 3 25 3 47; do
  timeout 600 $B --warmup $w > gpurun_out/r2n_bench.json 2> gpurun_out/r2n.err
  python -c "
import json; p=json.load(open('gpurun_out/r2n_bench.json')); print('warmup $w', p['ms_per_step'], 1000/p['e2e']['value'], p['exact_precision_ms_per_step'], p['clocks']['sm_mhz'])"
done
