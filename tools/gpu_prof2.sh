# Round-1 profile capture (run under gpurun): launch list of one frame + full captures of the dominant kernels.
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/prof2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/prof2_ncu1.log 2>&1
tail -2 gpurun_out/prof2_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -f -o gpurun_out/r01_dcn_tc_full $CMD > gpurun_out/prof2_ncu3.log 2>&1
python tools/conv_bench.py 64 64 3 1024 1920 2 3 > gpurun_out/plain_conv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/r01_conv3x3_64to64_1080p python tools/conv_bench.py 64 64 3 1024 1920 2 3 > gpurun_out/ncu_conv.log 2>&1
python tools/conv_bench.py 128 128 3 512 960 2 3 > gpurun_out/plain_conv2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/r01_conv3x3_128to128_split python tools/conv_bench.py 128 128 3 512 960 2 3 > gpurun_out/ncu_conv2.log 2>&1
./tools/mma_probe 100 > gpurun_out/mma_probe.log 2>&1
