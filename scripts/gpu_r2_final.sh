#!/bin/bash
# round 2: the final build -- smoke(), full GPU suite, the default bench line (what the driver runs at round end)
mkdir -p gpurun_out


echo "== bench default"
timeout 600 python bench.py > gpurun_out/r2j_bench_1gpu.json 2> gpurun_out/r2j_bench_1gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r2j_bench_1gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2j_bench_1gpu.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_full', 'parity_prefix', 'gpu_launches', 'clocks')})
    e = d['e2e']; print({k: e[k] for k in ('value', 'ms_per_step', 'h2d_bytes_per_step', 'd2h_bytes_per_step')}); print({k: {q: round(v[q], 2) for q in ('p50', 'p90', 'p99', 'max')} for k, v in e['tick_latency_ms'].items() if 'p50' in v})
    print(d['roofline']['frac'], d['roofline']['traffic'], d['roofline'].get('traffic_source'), d['cpu_baseline'])
except Exception as e: print('bench parse failed', e)
PY
