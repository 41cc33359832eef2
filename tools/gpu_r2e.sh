#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 22 --warmup 3 --no-cpu-baseline --no-eager-baseline"
for ms in 0 100 1000 0 100; do
  TDVC_BENCH_SMI_MS=$ms timeout 600 $B > gpurun_out/r2e_bench_smi$ms.json 2> gpurun_out/r2e.err
  python -c "
import json; p=json.load(open('gpurun_out/r2e_bench_smi$ms.json')); print('smi $ms', p['ms_per_step'], 1000/p['e2e']['value'], p['clocks'])"
done
timeout 600 python tools/sync_ab.py exact 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k fullres 2>&1 | tail -3
