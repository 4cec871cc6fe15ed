#!/bin/bash
# round 2, re-entry: status of the committed build -- GPU suite, default bench line, launch list, full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== gpu suite"; (time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) 2>&1
echo "== bench default"
timeout 600 python bench.py > gpurun_out/r2a_bench_1gpu.json 2> gpurun_out/r2a_bench_1gpu.err; echo "rc=$?"; tail -c 600 gpurun_out/r2a_bench_1gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2a_bench_1gpu.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_full', 'parity_prefix', 'gpu_launches', 'clocks')})
    print(d['e2e']); print(d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline']['bytes_per_launch']); print(d['place'])
except Exception as e: print('bench parse failed', e)
PY
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches_cfg3_1m.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu --latency-ticks 0 > gpurun_out/r2a_launches.log 2>&1; echo "rc=$?"
echo "== full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_merge|k_place|k_rows" -s 80 -c 5 -o gpurun_out/r2a_batch -f \
  python bench.py --steps 1 --warmup 0 --no-cpu --latency-ticks 0 > gpurun_out/r2a_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out/
