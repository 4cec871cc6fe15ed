"""The chunked parallel placement algorithm (stage / resolve / commit) restated on the CPU
(oracle/place_model.cpp) against the sequential oracle: placements, failure counters and the whole node table
must be identical for every chunk / candidate / list size.  This pins the ALGORITHM the CUDA kernel
(swarmkit_b200/csrc/kernel_place.cuh) implements; the kernel itself is compared with the oracle on the GPU
(tests/test_parity_gpu.py, tests/test_headline_gpu.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from swarmkit_b200 import workloads as W
from swarmkit_b200.abi import FlatABI
from tests import oracle_lib, randwork as R
from tests.oracle_lib import OracleEngine


class ModelEngine(FlatABI):
    def __init__(self, node_capacity: int = 0):
        super().__init__(oracle_lib._make("libplace_model.so"), "mpe_", node_capacity=node_capacity)
        self.lib.mpe_model_counters.argtypes = [C.c_void_p]
        self.lib.mpe_model_counters.restype = None

    def counters(self):
        out = np.zeros(4, np.uint64)
        self.lib.mpe_model_counters(out.ctypes.data)
        return dict(zip(("parallel", "cuts", "ambiguous", "tails"), (int(x) for x in out)))


def set_knobs(batch, chunk, k, listcap, group=32):
    os.environ.update(PM_BATCH=str(batch), PM_CHUNK=str(chunk), PM_K=str(k), PM_LISTCAP=str(listcap), PM_GROUP=str(group))


def run_pair(nodes, ticks, n_nodes, what):
    m, o = ModelEngine(n_nodes), OracleEngine(n_nodes)
    for e in (m, o):
        e.node_upsert(nodes)
        e.set_node_count(n_nodes)
    for t in ticks:
        R.compare_results(t, m.schedule(t), o.schedule(t), what)
    return m, o


# (batch, chunk, candidates per task [0 = position in the chunk + 1, what the kernel uses], list cap[, lanes per group])
KNOBS = [(512, 128, 0, 1024), (64, 8, 4, 16), (300, 32, 3, 1024), (97, 5, 2, 7), (1000, 64, 64, 64, 4), (200, 16, 0, 24, 8)]


@pytest.mark.parametrize("knobs", KNOBS)
@pytest.mark.parametrize("seed", range(6))
def test_random_streams(knobs, seed):
    set_knobs(*knobs)
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(60, 500))
    nodes = R.random_nodes(rng, n, tight=bool(seed % 2))
    ticks = [R.random_tick(rng, n, 250, p_oneoff=0.93, rotate=bool(seed % 3 == 0)) for _ in range(2)]
    m, o = run_pair(nodes, ticks, n, f"model knobs={knobs} seed={seed}")
    R.compare_state(m, o, n, 4, 3, 40, "model state")


@pytest.mark.parametrize("knobs", KNOBS)
def test_cfg3_shape(knobs):
    """Many tasks per descriptor, classes consumed inside a batch: the tail / ambiguity paths."""
    set_knobs(*knobs)
    w = W.cfg3("oneoff", n_nodes=1500, n_tasks=6000, n_services=12)
    m, o = run_pair(w.nodes, [w.tick], w.n_nodes, f"cfg3 knobs={knobs}")
    R.compare_state(m, o, w.n_nodes, 12, 0, 0, "cfg3 state")
    assert m.counters()["parallel"] > 0


@pytest.mark.parametrize("knobs", KNOBS[:3])
def test_cfg2_and_cfg4_shape(knobs):
    set_knobs(*knobs)
    w = W.cfg2("oneoff", n_nodes=400, n_tasks=3000)
    m, o = run_pair(w.nodes, [w.tick], w.n_nodes, "cfg2")
    R.compare_state(m, o, w.n_nodes, 1, 0, 0, "cfg2 state")
    w = W.cfg4("oneoff", n_nodes=1200, n_tasks=3000, n_services=30)
    m, o = run_pair(w.nodes, [w.tick], w.n_nodes, "cfg4")
    R.compare_state(m, o, w.n_nodes, 30, 3, 64, "cfg4 state")


def test_paths_are_exercised():
    """The interesting paths must actually run in this suite (tails, ambiguous lanes, cuts)."""
    set_knobs(256, 32, 6, 1024)
    w = W.cfg3("oneoff", n_nodes=800, n_tasks=8000, n_services=10)
    m, o = run_pair(w.nodes, [w.tick], w.n_nodes, "paths")
    c = m.counters()
    assert c["tails"] > 0 and c["ambiguous"] > 0, c
