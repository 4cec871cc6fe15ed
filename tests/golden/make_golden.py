#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: placements + failure counters of small seeded
workloads, computed by the pinned CPU oracle (oracle/flat_oracle.cpp, which is
itself pinned to the reference's known-answer tests -- DESIGN.md section 2).
The reference is Go and cannot run here, so these vectors are the oracle's, and
the fixtures exist so the GPU parity tests do not depend on rebuilding the
oracle.  Usage: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from swarmkit_b200 import workloads as W  # noqa: E402
from tests import randwork as R  # noqa: E402
from tests.oracle_lib import OracleEngine  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    yield "cfg1", W.cfg1(), None
    yield "cfg2_oneoff", W.cfg2("oneoff", n_nodes=300, n_tasks=2000, n_services=10), None
    yield "cfg2_grouped", W.cfg2("grouped", n_nodes=300, n_tasks=2000, n_services=10), None
    yield "cfg3_oneoff", W.cfg3("oneoff", n_nodes=2000, n_tasks=3000, n_services=30), None
    yield "cfg3_grouped", W.cfg3("grouped", n_nodes=2000, n_tasks=3000, n_services=30), None
    yield "cfg4_oneoff", W.cfg4("oneoff", n_nodes=1500, n_tasks=2000, n_services=40), None
    for seed in (11, 12, 13):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(100, 600))
        nodes = R.random_nodes(rng, n, tight=bool(seed % 2))
        tick = R.random_tick(rng, n, 150)
        yield f"random_{seed}", None, (nodes, tick, n)


def workload_of(name):
    for nm, w, raw in cases():
        if nm == name:
            return (w.nodes, w.tick, w.n_nodes) if w is not None else raw
    raise KeyError(name)


def names():
    return [nm for nm, _, _ in cases()]


def main():
    for nm, w, raw in cases():
        nodes, tick, n = (w.nodes, w.tick, w.n_nodes) if w is not None else raw
        o = OracleEngine(node_capacity=n)
        o.node_upsert(nodes)
        o.set_node_count(n)
        out_node, out_fail = o.schedule(tick)
        state = o.snapshot(0, n)
        np.savez_compressed(os.path.join(HERE, nm + ".npz"), out_node=out_node, out_fail=out_fail,
                            total=state["total_tasks"], cpu=state["cpu_avail"], mem=state["mem_avail"])
        print(nm, "tasks", out_node.size, "placed", int((out_node != 0xFFFFFFFF).sum()))


if __name__ == "__main__":
    main()
