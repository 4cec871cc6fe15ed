#!/bin/bash
echo "== gpu suite"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
bash scripts/gpu_r2_groups_prof.sh
timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu --latency-ticks 0 --workload cfg1 2>&1 | tail -1 | cut -c1-400
