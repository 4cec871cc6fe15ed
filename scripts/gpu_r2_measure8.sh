#!/bin/bash
# round 2: final build -- full GPU suite, other configurations, cfg4 at 1M nodes
mkdir -p gpurun_out
echo "== gpu suite"; (time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) 2>&1
for wl in cfg3-oneoff cfg5-grouped cfg3-grouped cfg2-oneoff cfg2-grouped cfg1 cfg4-oneoff; do
  echo "== bench $wl"
  case $wl in cfg5-grouped) extra="--tasks 500000 --no-cpu";; cfg2-*) extra="--tasks 100000 --nodes 10000 --no-cpu";; cfg4-oneoff) extra="--tasks 100000 --nodes 1000000 --cpu-sample 300 --steps 1 --warmup 1";; cfg3-oneoff) extra="--cpu-sample 3000";; *) extra="--no-cpu";; esac
  timeout 400 python bench.py --steps 3 --warmup 2 --latency-ticks 0 --workload $wl $extra 2>gpurun_out/r2h_$wl.err | tail -1 > gpurun_out/r2h_bench_$wl.json; tail -c 300 gpurun_out/r2h_$wl.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2h_bench_$wl.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_full', 'parity_prefix')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
except Exception as e: print('bench parse failed', e)
PY
done
