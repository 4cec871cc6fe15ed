# Two ranks on one box (gpurun --gpus 2): parity of the node-sharded group against the oracle, then a bench line.
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/mg_check.log 2>&1
grep -v "^W0\|Setting OMP\|^\*\*\*" gpurun_out/mg_check.log | tail -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu --latency-ticks 0 > gpurun_out/mg_bench.log 2>&1
grep "^{" gpurun_out/mg_bench.log | tail -1 | cut -c1-1500; grep -i "error\|Traceback" -A8 gpurun_out/mg_bench.log | head -30
