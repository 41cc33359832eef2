# Developer A/B run on a GPU box (under gpurun): kernel tests, then the bench (optionally a second time with an
# experiment switch given as "VAR=value" in $1).
set -x
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 11 --warmup 3 --no-cpu-baseline > gpurun_out/exp_base.json 2> gpurun_out/exp_base.err; tail -2 gpurun_out/exp_base.err
if [ -n "$1" ]; then
  env "$1" python bench.py --steps 11 --warmup 3 --no-cpu-baseline > gpurun_out/exp_alt.json 2> gpurun_out/exp_alt.err; tail -2 gpurun_out/exp_alt.err
fi
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
