# GPU parity suite + a short bench line (used through gpurun during development)
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --tasks ${1:-200000} --steps 2 --warmup 1 --no-cpu ${2:-} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','split_ms_per_step','paths')}, d['e2e'], {k:d['sequencer_cycles'][k] for k in ('fast','medium','generic','fast_exits')})"
