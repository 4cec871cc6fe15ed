"""Node-sharded engine group against the CPU oracle (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Every rank mirrors all nodes and submits the same ticks; each scans its slice of the node axis
(SURVEY 8e).  Checks: rank 0's placements / failure counters / final node state equal the oracle's,
and every rank produced exactly rank 0's placements.  Prints MULTI_GPU_CHECK PASS/FAIL on rank 0.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from swarmkit_b200 import PlacementEngine, workloads as W   # noqa: E402
from swarmkit_b200.engine import nccl_unique_id              # noqa: E402
from tests import randwork as R                               # noqa: E402
from tests.oracle_lib import OracleEngine                     # noqa: E402


def main():
    rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [
        ("cfg3-oneoff", W.cfg3("oneoff", n_nodes=20000, n_tasks=30000, n_services=100), 100),
        ("cfg2-oneoff", W.cfg2("oneoff", n_nodes=3000, n_tasks=20000, n_services=20), 1),
        ("cfg2-tight", W.cfg2("oneoff", n_nodes=300, n_tasks=6000, n_services=12), 1),
        ("cfg4-oneoff", W.cfg4("oneoff", n_nodes=12000, n_tasks=12000, n_services=60), 60),
        ("cfg1", W.cfg1(), 1),
    ]
    ok = True
    # the headline shape at its real node count and batch size, against the committed golden vector (no oracle run needed)
    from tests.golden import make_golden_big as GB
    wh = GB.workload("big_cfg3_oneoff_1m_100k")
    gold, _ = GB.load("big_cfg3_oneoff_1m_100k")
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    eng = PlacementEngine(node_capacity=wh.n_nodes, device=local, rank=rank, world_size=world, nccl_id=box[0])
    eng.node_upsert(wh.nodes)
    eng.set_node_count(wh.n_nodes)
    sub = wh.tick.slice_groups(0, 40_000)
    out_node, _ = eng.schedule(sub)
    bad = int((out_node != gold[:sub.n_tasks]).sum())
    flag = torch.tensor([bad], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if rank == 0:
        st = eng.stats()
        print(f"cfg3-headline-40k: mismatches(max over ranks)={int(flag.item())} place_tasks={st['place_tasks']} handed_over={st['place_cuts']}", flush=True)
        ok = ok and int(flag.item()) == 0
    eng.close()
    dist.barrier()
    for name, w, n_svc in cases:
        box = [nccl_unique_id() if rank == 0 else None]      # one communicator (one id) per engine group
        dist.broadcast_object_list(box, src=0)
        eng = PlacementEngine(node_capacity=w.n_nodes, device=local, rank=rank, world_size=world, nccl_id=box[0])
        eng.node_upsert(w.nodes)
        eng.set_node_count(w.n_nodes)
        out_node, out_fail = eng.schedule(w.tick)
        mine = torch.from_numpy(out_node.astype(np.int64)).cuda()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        same = bool((mine == ref).all().item())
        flag = torch.tensor([1 if same else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            cpu = OracleEngine(node_capacity=w.n_nodes)
            cpu.node_upsert(w.nodes)
            cpu.set_node_count(w.n_nodes)
            exp = cpu.schedule(w.tick)
            try:
                R.compare_results(w.tick, (out_node, out_fail), exp, name)
                R.compare_state(eng, cpu, w.n_nodes, n_svc, 0, 0, name)
                st = eng.stats()
                print(f"{name}: ok  ranks_agree={bool(flag.item())} placed={int((out_node != 0xFFFFFFFF).sum())} "
                      f"scan_launches={st['scan_launches']} fast={st['fast_path']} medium={st['medium_path']} slow={st['slow_path']}", flush=True)
            except AssertionError as e:
                ok = False
                print(f"{name}: MISMATCH vs oracle: {str(e)[:300]}", flush=True)
            ok = ok and bool(flag.item())
        eng.close()
        dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok or rank != 0 else 1)


if __name__ == "__main__":
    main()
