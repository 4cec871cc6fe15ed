#!/bin/bash
# round 2: sequential candidate-parallel resolve + concurrent commit; cluster 8 vs 16; cfg5 storm; other configs
mkdir -p gpurun_out
echo "== gpu suite"; (time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) 2>&1
for cl in 8 16; do
  echo "== bench cluster=$cl"
  PE_PLACE_CLUSTER=$cl timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --latency-ticks 0 > gpurun_out/r2c_bench_cl$cl.json 2> gpurun_out/r2c_bench_cl$cl.err; echo "rc=$?"; tail -c 400 gpurun_out/r2c_bench_cl$cl.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2c_bench_cl$cl.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_full')})
    print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch']); print(d['place'])
except Exception as e: print('bench parse failed', e)
PY
done
for wl in cfg5-grouped cfg3-grouped cfg2-oneoff cfg2-grouped; do
  echo "== bench $wl"
  case $wl in cfg5-grouped) extra="--tasks 500000";; cfg2-*) extra="--tasks 100000 --nodes 10000";; *) extra="";; esac
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --latency-ticks 0 --workload $wl $extra 2>gpurun_out/r2c_$wl.err | tail -1 > gpurun_out/r2c_bench_$wl.json; tail -c 300 gpurun_out/r2c_$wl.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2c_bench_$wl.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_full')}, d['e2e']['value'])
except Exception as e: print('bench parse failed', e)
PY
done
