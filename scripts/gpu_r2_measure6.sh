#!/bin/bash
# round 2: full GPU suite, the default bench line, the reference arm, launch list and full captures of the final build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== gpu suite"; (time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) 2>&1
echo "== bench default"
timeout 600 python bench.py > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r2f_bench_1gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2f_bench_1gpu.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_full', 'parity_prefix', 'gpu_launches', 'clocks')})
    print(d['e2e']); print(d['roofline']['frac'], d['roofline']['ms_per_launch']); print(d['place'])
except Exception as e: print('bench parse failed', e)
PY
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2f_bench_reference_arm.json 2>&1; tail -c 600 gpurun_out/r2f_bench_reference_arm.json
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2f_launches_cfg3_1m.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu --latency-ticks 0 > gpurun_out/r2f_launches.log 2>&1; echo "rc=$?"
echo "== full captures (one batch in the middle of the 1M-task tick)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_scan|k_merge|k_place|k_rows" -s 150 -c 5 -o gpurun_out/r2f_batch -f \
  python bench.py --steps 1 --warmup 0 --no-cpu --latency-ticks 0 > gpurun_out/r2f_full.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_groups" -s 1 -c 1 -o gpurun_out/r2f_groups -f \
  python bench.py --steps 1 --warmup 1 --no-cpu --latency-ticks 0 --workload cfg3-grouped > gpurun_out/r2f_full_groups.log 2>&1; echo "rc=$?"
ls -la gpurun_out/ | head -30
