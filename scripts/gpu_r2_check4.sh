#!/bin/bash
# round 2: k_groups (task groups on the whole machine) -- full GPU suite, then grouped benches
mkdir -p gpurun_out
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for wl in cfg3-grouped cfg2-grouped; do
  echo "== bench $wl"
  if [ $wl = cfg2-grouped ]; then extra="--tasks 100000 --nodes 10000"; else extra=""; fi
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --latency-ticks 0 --workload $wl $extra 2>&1 | tail -2 | tee gpurun_out/r2q_bench_$wl.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:600]); continue
    print({k: d[k] for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_full') if k in d}, d['e2e']['value'])
"
done
