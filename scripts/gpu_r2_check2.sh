#!/bin/bash
# round 2: placement step v2 (slot-only commit, shared-memory touched bitmap, floor-rank tails) -- parity, then a batch-size sweep
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_headline_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -15
for mb in 16384; do
  echo "== bench max_batch=$mb"
  timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu --max-batch $mb 2>&1 | tail -2 | tee gpurun_out/r2l_bench_mb$mb.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:300]); continue
    print({k: d[k] for k in ('value', 'ms_per_step', 'split_ms_per_step', 'place', 'paths', 'parity_full') if k in d}, d['e2e']['value'], d['roofline']['frac'])
"
done
