set -x
python tools/conv_bench.py 64 32 7 1024 1920 2 3 > gpurun_out/plain_7.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_7x7 python tools/conv_bench.py 64 32 7 1024 1920 2 3 > gpurun_out/ncu_7.log 2>&1
