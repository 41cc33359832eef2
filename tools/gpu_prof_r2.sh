#!/bin/bash
# Round-2 profile capture (run under gpurun): launch list of one steady-state frame + full captures of the dominant kernels.
mkdir -p gpurun_out
set -x
python tools/profile_frame.py > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/profile_frame.py > gpurun_out/r02_prof_ncu1.log 2>&1
tail -2 gpurun_out/r02_prof_ncu1.log
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -f -o gpurun_out/r02_dcn_tc_full python tools/profile_frame.py > gpurun_out/r02_prof_ncu2.log 2>&1
python tools/conv_bench.py 64 64 3 1024 1920 2 3 > gpurun_out/r02_plain_conv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_conv3x3_64to64_1080p python tools/conv_bench.py 64 64 3 1024 1920 2 3 > gpurun_out/r02_ncu_conv.log 2>&1
python tools/conv_bench.py 64 64 3 1024 1920 2 3 1 1 0 1 0 0 1 > gpurun_out/r02_plain_conv_p1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_conv3x3_64to64_p1 python tools/conv_bench.py 64 64 3 1024 1920 2 3 1 1 0 1 0 0 1 > gpurun_out/r02_ncu_conv_p1.log 2>&1
python tools/conv_bench.py 128 128 3 512 960 2 3 1 1 0 1 0 0 1 > gpurun_out/r02_plain_conv_p1b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_conv3x3_128to128_p1 python tools/conv_bench.py 128 128 3 512 960 2 3 1 1 0 1 0 0 1 > gpurun_out/r02_ncu_conv_p1b.log 2>&1
cat gpurun_out/r02_plain_conv.log gpurun_out/r02_plain_conv_p1.log gpurun_out/r02_plain_conv_p1b.log gpurun_out/r02_prof_plain.log
