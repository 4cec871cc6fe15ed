#!/bin/bash
# ncu source-level capture of one k_place launch in the middle of a 300k-task tick
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_place -s 30 -c 1 -o gpurun_out/r2_place -f \
  python bench.py --steps 1 --warmup 0 --no-cpu --tasks 300000 > gpurun_out/r2_place_ncu.log 2>&1
tail -3 gpurun_out/r2_place_ncu.log
ls -la gpurun_out/
