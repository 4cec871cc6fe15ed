#!/usr/bin/env python
"""Regenerates tests/golden/big_*.npz: placements of the FULL-SIZE BASELINE
configurations, computed once by the pinned CPU oracle (oracle/flat_oracle.cpp,
single thread; cfg3-oneoff 1M x 100k takes ~7 CPU-minutes).

The tick is sequential (scheduler.go:464-469: one group after the other, every
placement visible to the next), so the placements of the first n tasks of a tick
are the placements of the n-task tick: one full-size vector pins every prefix
length, which is how tests/test_headline_gpu.py and bench.py (`parity_prefix`,
`parity_full`) use it.

Usage: python tests/golden/make_golden_big.py [name ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from swarmkit_b200 import workloads as W  # noqa: E402
from tests.oracle_lib import OracleEngine  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (workload factory kwargs); sizes are BASELINE.json's where one CPU core finishes in minutes
CASES = {
    "big_cfg3_oneoff_1m_100k": lambda: W.cfg3("oneoff", n_nodes=100_000, n_tasks=1_000_000),
    "big_cfg2_oneoff_100k_10k": lambda: W.cfg2("oneoff", n_nodes=10_000, n_tasks=100_000),
    "big_cfg2_grouped_100k_10k": lambda: W.cfg2("grouped", n_nodes=10_000, n_tasks=100_000),
    "big_cfg3_grouped_1m_100k": lambda: W.cfg3("grouped", n_nodes=100_000, n_tasks=1_000_000),
    "big_cfg4_oneoff_60k_200k": lambda: W.cfg4("oneoff", n_nodes=200_000, n_tasks=60_000),
    "big_cfg5_storm_500k_100k": lambda: W.cfg5(n_nodes=100_000, n_tasks=500_000, n_services=5000),
}


def path(name: str) -> str:
    return os.path.join(HERE, name + ".npz")


_CACHE = {}


def workload(name: str):
    """The seeded workload of a case (cached: the generators take seconds at these sizes; callers must not mutate it)."""
    if name not in _CACHE:
        _CACHE[name] = CASES[name]()
    return _CACHE[name]


def load(name: str):
    """out_node (uint32, PE_NONE = unplaced) and the final per-node task totals of the golden run."""
    z = np.load(path(name))
    lo, hi = z["node_lo16"], z["node_hi8"]
    out = lo.astype(np.uint32) | (hi.astype(np.uint32) << 16)
    out[out == 0xFFFFFF] = 0xFFFFFFFF
    return out, z["total"]


def main():
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        w = workload(nm)
        o = OracleEngine(node_capacity=w.n_nodes)
        o.node_upsert(w.nodes)
        o.set_node_count(w.n_nodes)
        t0 = time.time()
        out_node, _ = o.schedule(w.tick)
        dt = time.time() - t0
        state = o.snapshot(0, w.n_nodes)
        assert w.n_nodes < 0xFFFFFF
        v = np.where(out_node == 0xFFFFFFFF, 0xFFFFFF, out_node).astype(np.uint32)
        # two byte planes compress better than one u32 array of 17-bit values
        np.savez_compressed(path(nm), node_lo16=(v & 0xFFFF).astype(np.uint16), node_hi8=(v >> 16).astype(np.uint8),
                            total=state["total_tasks"].astype(np.uint32))
        print(nm, "tasks", out_node.size, "placed", int((out_node != 0xFFFFFFFF).sum()), f"{dt:.1f}s",
              os.path.getsize(path(nm)), "bytes", flush=True)


if __name__ == "__main__":
    main()
