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
