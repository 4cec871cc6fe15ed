#!/bin/bash
# round 2: resolve v4 (cooperative pre-filter + light rounds, concurrent commit); cluster 8 vs 16
mkdir -p gpurun_out
echo "== headline + parity suites"; (time timeout 900 python -m pytest tests/test_headline_gpu.py tests/test_parity_gpu.py -q -x 2>&1 | tail -6) 2>&1
for cl in 8 16; do
  echo "== bench cluster=$cl"
  PE_PLACE_CLUSTER=$cl timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --latency-ticks 0 > gpurun_out/r2d_bench_cl$cl.json 2> gpurun_out/r2d_bench_cl$cl.err; echo "rc=$?"; tail -c 400 gpurun_out/r2d_bench_cl$cl.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2d_bench_cl$cl.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_full')})
    print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch']); print(d['place'])
except Exception as e: print('bench parse failed', e)
PY
done
