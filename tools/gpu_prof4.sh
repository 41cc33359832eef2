set -x
python tools/conv_bench.py 64 64 3 1024 1920 2 3 1 1 0 2 1 > gpurun_out/plain_res.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_res python tools/conv_bench.py 64 64 3 1024 1920 2 3 1 1 0 2 1 > gpurun_out/ncu_res.log 2>&1
