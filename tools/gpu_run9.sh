python tools/tc_check.py | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/conv_bench.py 64 64 3 1024 1920 2 20 1 1 0 2 1
python tools/conv_bench.py 64 64 3 1024 1920 2 20 1 1 0 1 0
python tools/conv_bench.py 128 128 3 512 960 2 20 1 1 0 2 1
python tools/conv_bench.py 128 128 3 512 960 2 20 1 1 0 1 0
python tools/conv_bench.py 128 128 1 512 960 2 20 1 1 0 0 1
python bench.py --steps 22 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('frame', d['ms_per_step'], d['e2e']['value'], d['stats'])"
