#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "one_product or chan_sum or absmax or gdn or spynet or dcn_backward or planar or tcgen05" > gpurun_out/r2m_memcheck.log 2>&1
echo "exit $?"; tail -15 gpurun_out/r2m_memcheck.log
