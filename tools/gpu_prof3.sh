# ncu full capture of one conv_bench shape (run under gpurun): bash tools/gpu_prof3.sh <name> <conv_bench args...>
set -x
NAME=$1; shift
python tools/conv_bench.py "$@" > gpurun_out/plain_$NAME.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o gpurun_out/$NAME python tools/conv_bench.py "$@" > gpurun_out/ncu_$NAME.log 2>&1
cat gpurun_out/plain_$NAME.log
