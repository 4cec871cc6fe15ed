timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for mb in 0 9472 18944; do
timeout 200 python bench.py --tasks 200000 --steps 2 --warmup 1 --no-cpu --max-batch $mb 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($mb, {k:d[k] for k in ('value','split_ms_per_step','paths')}, {k:d['sequencer_cycles'][k] for k in ('fast','medium','generic','fast_exits','ordered_warp_wait','ordered_warp_work','prof')})"
done
