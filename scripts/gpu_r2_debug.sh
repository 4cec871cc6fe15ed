#!/bin/bash
mkdir -p gpurun_out
for mb in 0 16384; do
echo "== full tick max_batch=$mb"
timeout 120 python - <<PY
import time, numpy as np, sys
sys.path.insert(0, '.')
from swarmkit_b200 import PlacementEngine
from tests.golden import make_golden_big as GB
w = GB.workload("big_cfg3_oneoff_1m_100k"); gold, _ = GB.load("big_cfg3_oneoff_1m_100k")
eng = PlacementEngine(node_capacity=w.n_nodes, max_batch=$mb)
for n in (200000, 1000000):
    eng.node_upsert(w.nodes); eng.set_node_count(w.n_nodes)
    sub = w.tick.slice_groups(0, n)
    t0 = time.time(); out, _ = eng.schedule(sub); dt = time.time() - t0
    st = eng.stats()
    print(n, "tasks", round(dt, 3), "s mism", int((out != gold[:n]).sum()), {k: st[k] for k in ("place_tasks", "place_cuts", "place_chunks", "place_amb", "place_tails", "seq_prof")}, flush=True)
    eng.stats_reset()
PY
done
echo "== small ticks"
timeout 120 python - <<PY
import time, numpy as np, sys
sys.path.insert(0, '.')
from swarmkit_b200 import PlacementEngine
from tests.golden import make_golden_big as GB
w = GB.workload("big_cfg3_oneoff_1m_100k")
eng = PlacementEngine(node_capacity=w.n_nodes)
eng.node_upsert(w.nodes); eng.set_node_count(w.n_nodes)
for tt in (1000, 10000):
    lat = []
    for i in range(30):
        s = w.tick.slice_groups(i * tt, (i + 1) * tt)
        t0 = time.time(); eng.schedule(s); lat.append(time.time() - t0)
    print(tt, "per tick ms", [round(1e3 * x, 2) for x in lat[:5]], "median", round(1e3 * float(np.median(lat)), 3), flush=True)
PY
