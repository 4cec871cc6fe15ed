#!/bin/bash
for wl in cfg3-grouped cfg2-grouped; do
  if [ $wl = cfg2-grouped ]; then extra="--tasks 100000 --nodes 10000"; else extra=""; fi
  timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu --latency-ticks 0 --workload $wl $extra 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l[:600]); continue
    print('$wl', d['value'], d['ms_per_step'], d['parity_full'], 'phase cycles (eval, thresh, count, offsets, compact, stage, fill, wb):', d['sequencer_cycles']['prof'][8:16])
"
done
