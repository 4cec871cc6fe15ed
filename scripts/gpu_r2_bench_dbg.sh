#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "
import faulthandler, sys, runpy
faulthandler.dump_traceback_later(150, exit=True)
sys.argv = ['bench.py', '--steps', '2', '--warmup', '2', '--no-cpu', '--max-batch', '16384']
runpy.run_path('bench.py', run_name='__main__')
" 2>&1 | tail -40 | cut -c1-2500
