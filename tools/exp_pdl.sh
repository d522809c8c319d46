python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do
NBE_NO_PDL=1 python bench.py --steps 50 --warmup 5 --no-incumbent 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NO_PDL', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'])"
python bench.py --steps 50 --warmup 5 --no-incumbent 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL   ', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'])"
done
