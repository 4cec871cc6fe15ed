#!/bin/bash
# round 2: warp-local resolve of k_place; cluster of 8 vs 16 CTAs
mkdir -p gpurun_out
echo "== gpu suite (cluster 16 where available)"; (time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) 2>&1
for cl in 8 16; do
  echo "== bench cluster=$cl"
  PE_PLACE_CLUSTER=$cl timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --latency-ticks 0 > gpurun_out/r2b_bench_cl$cl.json 2> gpurun_out/r2b_bench_cl$cl.err; echo "rc=$?"; tail -c 400 gpurun_out/r2b_bench_cl$cl.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2b_bench_cl$cl.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_full')})
    print(d['e2e']['value'], d['roofline']['frac']); print(d['place'])
except Exception as e: print('bench parse failed', e)
PY
done
echo "== cluster 8 headline tests"; PE_PLACE_CLUSTER=8 timeout 600 python -m pytest tests/test_headline_gpu.py -q -x 2>&1 | tail -4
