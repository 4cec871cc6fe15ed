# Round-end evidence on one B200 (through gpurun): smoke, the default bench line, the reference arm,
# the ncu launch list and full captures of the scan / sequencer kernels.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --tasks 100000 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_scan -s 12 -c 2 -o gpurun_out/scan_r1e -f python bench.py --tasks 60000 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_scan.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_sequencer -s 6 -c 1 -o gpurun_out/seq_r1e -f python bench.py --tasks 60000 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_seq.log 2>&1
ls -la gpurun_out/
