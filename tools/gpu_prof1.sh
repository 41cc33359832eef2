set -x
python tools/conv_bench.py 64 64 3 1024 1920 2 3 > gpurun_out/plain_conv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 2 -f -o gpurun_out/prof_conv64_v2 python tools/conv_bench.py 64 64 3 1024 1920 2 3 > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/ncu_conv.log
