#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sync_ab.py exact > gpurun_out/r2c_sync_ab.log 2>&1
timeout 600 python tools/sync_ab.py mixed >> gpurun_out/r2c_sync_ab.log 2>&1
timeout 1200 python tools/fullres_diag.py > gpurun_out/r2c_diag.log 2>&1
cat gpurun_out/r2c_sync_ab.log; tail -40 gpurun_out/r2c_diag.log
