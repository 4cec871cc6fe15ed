#!/bin/bash
# round 2, first GPU contact of the chunked parallel placement step: targeted parity, the whole GPU suite, a short bench
mkdir -p gpurun_out
nvidia-smi -L
echo "== targeted"; timeout 600 python -m pytest tests/test_headline_gpu.py -x -q -k "60k_tasks" 2>&1 | tail -30
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 5000 2>&1 | tail -5 | tee gpurun_out/r2a_bench.json
