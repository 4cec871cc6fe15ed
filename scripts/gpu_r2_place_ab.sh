#!/bin/bash
# round 2: A-B runs of k_place variants (PE_PLACE_CLUSTER = 8 / 16): headline parity tests, then the bench line with the per-phase counters
mkdir -p gpurun_out
echo skip-suite
for cc in 16; do
  echo "== bench cluster=$cc"
  PE_PLACE_CLUSTER=$cc timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --latency-ticks 0 > gpurun_out/r2e_bench_cc$cc.json 2> gpurun_out/r2e_bench_cc$cc.err; echo "rc=$?"; tail -c 400 gpurun_out/r2e_bench_cc$cc.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2e_bench_cc$cc.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'parity_full')})
    print(d['place'])
except Exception as e: print('bench parse failed', e)
PY
done
