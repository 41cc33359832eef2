B=$PWD/tdvc_b200/libtdvc_b200_B.so
python tools/tc_check.py | tail -1
for i in 1 2; do
python tools/conv_bench.py 64 64 3 1024 1920 2 20 1 1 0 2 1 | sed 's/^/A /'
TDVC_B200_LIB=$B python tools/conv_bench.py 64 64 3 1024 1920 2 20 1 1 0 2 1 | sed 's/^/B /'
python tools/conv_bench.py 128 128 3 512 960 2 20 1 1 0 2 1 | sed 's/^/A /'
TDVC_B200_LIB=$B python tools/conv_bench.py 128 128 3 512 960 2 20 1 1 0 2 1 | sed 's/^/B /'
python tools/conv_bench.py 64 128 3 1024 1920 2 20 1 2 | sed 's/^/A /'
TDVC_B200_LIB=$B python tools/conv_bench.py 64 128 3 1024 1920 2 20 1 2 | sed 's/^/B /'
python bench.py --steps 11 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('A frame', d['ms_per_step'], d['e2e']['value'])"
TDVC_B200_LIB=$B python bench.py --steps 11 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B frame', d['ms_per_step'], d['e2e']['value'])"
done
