#!/bin/bash
timeout 200 python - <<PY 2>&1 | tail -12
import faulthandler, time, sys, os, numpy as np
faulthandler.dump_traceback_later(170, exit=True)
sys.path.insert(0, '.')
from swarmkit_b200 import PlacementEngine
from tests.golden import make_golden_big as GB
w = GB.workload("big_cfg3_oneoff_1m_100k")
eng = PlacementEngine(node_capacity=w.n_nodes, max_batch=16384)
tt = 10000
eng.node_upsert(w.nodes); eng.set_node_count(w.n_nodes)
subs = [w.tick.slice_groups(i * tt, (i + 1) * tt) for i in range(100)]
for i in range(1000):
    if i % 100 == 0: sys.stderr.write(f"tick #{i}\n")
    out, _ = eng.schedule(subs[i % 100])
    if (out == 0xFFFFFFFF).any():
        print("tick", i, "unplaced", int((out == 0xFFFFFFFF).sum()), flush=True)
st = eng.stats()
print("done", {k: st[k] for k in ("place_tasks", "place_cuts", "place_amb", "slow_path", "fast_path", "medium_path", "seq_prof")}, flush=True)
PY
