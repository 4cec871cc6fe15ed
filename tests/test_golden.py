"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them.  GPU: the CUDA engine reproduces them
bit-exactly without the oracle being involved at all."""
import os

import numpy as np
import pytest

from swarmkit_b200 import abi
from tests.golden import make_golden as G

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check(engine, name):
    nodes, tick, n = G.workload_of(name)
    gold = np.load(os.path.join(HERE, name + ".npz"))
    engine.node_upsert(nodes)
    engine.set_node_count(n)
    out_node, out_fail = engine.schedule(tick)
    assert (out_node == gold["out_node"]).all(), f"{name}: placements differ from the golden vector"
    g = tick.groups
    for i in range(g.size):
        t0, k = int(g[i]["task_off"]), int(g[i]["n_tasks"])
        if k and (out_node[t0:t0 + k] == abi.PE_NONE).any():
            assert (out_fail[i] == gold["out_fail"][i]).all(), f"{name}: failure counters of group {i}"
    st = engine.snapshot(0, n)
    assert (st["total_tasks"] == gold["total"]).all() and (st["cpu_avail"] == gold["cpu"]).all() and (st["mem_avail"] == gold["mem"]).all()


@pytest.mark.parametrize("name", G.names())
def test_oracle_matches_golden(name):
    from tests.oracle_lib import OracleEngine
    check(OracleEngine(), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", G.names())
def test_engine_matches_golden(name):
    from swarmkit_b200 import PlacementEngine
    check(PlacementEngine(), name)


@pytest.mark.parametrize("name,n", [("big_cfg3_oneoff_1m_100k", 1500), ("big_cfg2_oneoff_100k_10k", 20000),
                                    ("big_cfg4_oneoff_60k_200k", 300), ("big_cfg3_grouped_1m_100k", 3000)])
def test_oracle_reproduces_big_golden_prefix(name, n):
    """The full-size vectors (tests/golden/make_golden_big.py) are the oracle's: a prefix is recomputed here (the tick
    is sequential, so the first n tasks of the golden tick are the golden n-task tick)."""
    from tests.golden import make_golden_big as GB
    from tests.oracle_lib import OracleEngine
    w = GB.workload(name)
    gold, total = GB.load(name)
    assert gold.size == w.tick.n_tasks and total.size == w.n_nodes
    g = w.tick.groups
    n_groups = int(np.searchsorted(g["task_off"].astype(np.int64) + g["n_tasks"], n, side="right"))
    sub = w.tick.slice_groups(0, max(n_groups, 1))
    o = OracleEngine(node_capacity=w.n_nodes)
    o.node_upsert(w.nodes)
    o.set_node_count(w.n_nodes)
    out, _ = o.schedule(sub)
    assert (out == gold[:sub.n_tasks]).all()
