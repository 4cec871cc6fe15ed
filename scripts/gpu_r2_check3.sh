#!/bin/bash
# round 2: full GPU suite, then the bench at two batch sizes (parity, latency arm included)
mkdir -p gpurun_out
echo "== gpu suite"; timeout 900 python -m pytest tests/test_headline_gpu.py -q -x 2>&1 | tail -6
for mb in 16384; do
  echo "== bench max_batch=$mb"
  timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 3000 --max-batch $mb 2>&1 | tail -2 | tee gpurun_out/r2o_bench_mb$mb.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: d[k] for k in ('value', 'ms_per_step', 'split_ms_per_step', 'place', 'paths', 'parity_full', 'parity_prefix') if k in d}, d['e2e'], d['roofline']['frac'])
"
done
