#!/bin/bash
timeout 200 python - <<PY 2>&1 | tail -30
import faulthandler, time, sys, os, numpy as np
faulthandler.dump_traceback_later(150, exit=True)
sys.path.insert(0, '.')
from swarmkit_b200 import PlacementEngine
from tests.golden import make_golden_big as GB
w = GB.workload("big_cfg3_oneoff_1m_100k"); gold, _ = GB.load("big_cfg3_oneoff_1m_100k")
eng = PlacementEngine(node_capacity=w.n_nodes, max_batch=16384)
for tt in (1000, 10000):
    eng.node_upsert(w.nodes); eng.set_node_count(w.n_nodes)
    t_all = time.time()
    for i in range(1000000 // tt):
        s = w.tick.slice_groups(i * tt, (i + 1) * tt)
        t0 = time.time()
        if time.time() - t_all > 40 and not os.environ.get("PE_DEBUG_SYNC"):
            pass
        sys.stderr.write(f"tick {tt} #{i}\n") if (i % 50 == 0) else None
        out, _ = eng.schedule(s)
        dt = time.time() - t0
        bad = int((out != gold[i * tt:(i + 1) * tt]).sum())
        if bad or dt > 1.0:
            print("tick", tt, i, "dt", round(dt, 3), "mismatches", bad, flush=True)
    print("done", tt, round(time.time() - t_all, 2), "s", flush=True)
PY
