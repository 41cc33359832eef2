set -x
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline"
$CMD > gpurun_out/prof6_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dcn_tc_kernel -c 1 -f -o gpurun_out/r01_dcn_tc_full $CMD > gpurun_out/prof6_ncu.log 2>&1
