#!/bin/bash
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err ) 2> gpurun_out/r2g_time.log
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err ) 2>> gpurun_out/r2g_time.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
cat gpurun_out/r2g_time.log gpurun_out/r2g_smoke.log | tail -12; head -c 600 gpurun_out/r2g_bench_default.json; echo; cat gpurun_out/r2g_bench_reference.json | head -c 1200; nproc
