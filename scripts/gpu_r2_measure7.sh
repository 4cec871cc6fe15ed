#!/bin/bash
# round 2: scan at 3 CTAs per SM, failure counters fetched on demand
mkdir -p gpurun_out
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== bench"
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 3000 --latency-ticks 0 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "rc=$?"; tail -c 300 gpurun_out/r2g_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'parity_full', 'parity_prefix')})
    print({k: d['e2e'][k] for k in ('value', 'ms_per_step', 'h2d_bytes_per_step', 'd2h_bytes_per_step')}); print(d['roofline']['frac'], d['roofline']['ms_per_launch'])
except Exception as e: print('bench parse failed', e)
PY
