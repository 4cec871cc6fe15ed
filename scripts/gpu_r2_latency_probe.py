"""Which 10k-task ticks are slow, and why (hand-overs to the ordered sequencer? tails?): per-tick latency + engine counters."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swarmkit_b200 import PlacementEngine, workloads as W

w = W.cfg3("oneoff")
eng = PlacementEngine(node_capacity=w.n_nodes)
eng.node_upsert(w.nodes); eng.set_node_count(w.n_nodes)
tt = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
subs = [w.tick.slice_groups(i * tt, (i + 1) * tt) for i in range(w.tick.n_groups // tt)]
rows = []
prev = eng.stats()
for i in range(1000):
    s = subs[i % len(subs)]
    t0 = time.perf_counter(); eng.schedule(s); dt = 1e3 * (time.perf_counter() - t0)
    st = eng.stats()
    rows.append((i, dt, st["place_cuts"] - prev["place_cuts"], st["slow_path"] - prev["slow_path"], st["medium_path"] - prev["medium_path"],
                 st["place_tails"] - prev["place_tails"], st["place_chunks"] - prev["place_chunks"], st["scan_ms"] - prev["scan_ms"], st["place_ms"] - prev["place_ms"], st["sequencer_ms"] - prev["sequencer_ms"], st["prep_ms"] - prev["prep_ms"]))
    prev = st
rows.sort(key=lambda r: -r[1])
print("tick ms cuts slow medium tails chunks scan_ms place_ms seq_ms prep_ms")
for r in rows[:14]: print(" ".join(f"{x:.2f}" if isinstance(x, float) else str(x) for x in r))
print("...median")
for r in rows[len(rows) // 2 - 2: len(rows) // 2 + 2]: print(" ".join(f"{x:.2f}" if isinstance(x, float) else str(x) for x in r))
print("cuts total", sum(r[2] for r in rows), "ticks with cuts", sum(1 for r in rows if r[2]))
by = sorted(rows, key=lambda r: r[0])
for lo in range(0, 1000, 100):
    seg = by[lo:lo + 100]
    print(f"ticks {lo}-{lo+99}: median {np.median([r[1] for r in seg]):.2f} ms  max {max(r[1] for r in seg):.2f}  cuts {sum(r[2] for r in seg)}  slow {sum(r[3] for r in seg)}  tails {sum(r[5] for r in seg)}  place_ms {np.median([r[8] for r in seg]):.2f}  seq_ms {np.median([r[9] for r in seg]):.2f} max_seq {max(r[9] for r in seg):.2f}")
