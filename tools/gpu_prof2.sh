# Round-1 profile capture (run under gpurun): launch list of one frame + full capture of the dominant conv kernel.
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/prof2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/prof2_ncu1.log 2>&1
tail -2 gpurun_out/prof2_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 60 -c 3 -f -o gpurun_out/r01_conv_tc_full $CMD > gpurun_out/prof2_ncu2.log 2>&1
tail -2 gpurun_out/prof2_ncu2.log
ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -f -o gpurun_out/r01_dcn_tc_full $CMD > gpurun_out/prof2_ncu3.log 2>&1
tail -2 gpurun_out/prof2_ncu3.log
