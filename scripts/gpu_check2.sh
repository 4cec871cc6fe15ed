timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for args in "--workload cfg2-oneoff --tasks 100000 --nodes 10000" "--workload cfg3-oneoff"; do
  timeout 300 python bench.py $args --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t); print('$args', {k:d[k] for k in ('value','ms_per_step','split_ms_per_step','paths')}, d['e2e']['value'])
except Exception as e: print('$args', 'FAILED', t[-400:])"
done
