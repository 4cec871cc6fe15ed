#!/bin/bash
# N ranks on one box (gpurun --gpus N): the node-heavy configuration (cfg4 one-off, 100k tasks x 1M nodes), where the scan dominates
N=${1:-2}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --latency-ticks 0 --workload cfg4-oneoff --tasks 100000 --nodes 1000000 --cpu-sample 300 > gpurun_out/mg_cfg4_$N.log 2>&1
grep "^{" gpurun_out/mg_cfg4_$N.log | tail -1 > gpurun_out/r2h_bench_cfg4-oneoff_${N}gpu.json
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2h_bench_cfg4-oneoff_${N}gpu.json').read())
    print($N, {k: d.get(k) for k in ('value', 'ms_per_step', 'split_ms_per_step', 'paths', 'parity_prefix')})
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/mg_cfg4_$N.log').read()[-1500:])
PY
