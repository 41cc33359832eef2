python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for i in 1 2; do
python bench.py --steps 22 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('A pdl ', d['ms_per_step'], d['e2e']['value'])"
TDVC_B200_NO_PDL=1 python bench.py --steps 22 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B nopdl', d['ms_per_step'], d['e2e']['value'])"
done
