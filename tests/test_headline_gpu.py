"""GPU parity at the sizes BASELINE.json quotes, against the committed full-size golden vectors
(tests/golden/big_*.npz, made once by tests/golden/make_golden_big.py with the pinned CPU oracle).  The oracle is not
involved at run time.  The tick is sequential, so the golden placements of the first n tasks are the golden
placements of the n-task tick: one vector pins every prefix length.

What these cover that the small parity tests cannot: default-size batches (4736 tasks against 100k nodes), hundreds
of them in one tick, ~1000 scan rows per batch, classes consumed inside a batch at that density."""
import numpy as np
import pytest

from swarmkit_b200 import PlacementEngine, abi
from tests.golden import make_golden_big as GB

pytestmark = pytest.mark.gpu


def run_prefix(name, n_tasks=None, flags=0, max_batch=0):
    w = GB.workload(name)
    gold_node, gold_total = GB.load(name)
    tick = w.tick if n_tasks is None else w.tick.slice_groups(0, n_tasks)
    eng = PlacementEngine(node_capacity=w.n_nodes, flags=flags, max_batch=max_batch)
    eng.node_upsert(w.nodes)
    eng.set_node_count(w.n_nodes)
    out_node, _ = eng.schedule(tick)
    gold = gold_node[:tick.n_tasks]
    bad = np.flatnonzero(out_node != gold)
    assert bad.size == 0, f"{name}: {bad.size} of {tick.n_tasks} placements differ from the golden vector, first at task {bad[:5]}: {out_node[bad[:5]]} vs {gold[bad[:5]]}"
    if n_tasks is None:
        st = eng.snapshot(0, w.n_nodes)
        assert (st["total_tasks"] == gold_total).all(), f"{name}: per-node task totals differ"
    return eng.stats()


def test_cfg3_oneoff_60k_tasks_round1_batches():
    """12 full batches of round 1's size (4736) of the headline workload; the chunked parallel step places them."""
    st = run_prefix("big_cfg3_oneoff_1m_100k", 60_000, max_batch=4736)
    assert st["scan_launches"] >= 12
    assert st["place_tasks"] > 0.9 * 60_000, st


def test_cfg3_oneoff_60k_tasks_default_batches():
    """The default batch size (a sixth of the nodes, at most 16384)."""
    st = run_prefix("big_cfg3_oneoff_1m_100k", 60_000)
    assert 3 <= st["scan_launches"] <= 5
    assert st["place_tasks"] > 0.9 * 60_000, st


def test_cfg3_oneoff_60k_tasks_ordered_sequencer():
    """Same prefix through the ordered sequencer alone (the fall-back of the parallel step)."""
    st = run_prefix("big_cfg3_oneoff_1m_100k", 60_000, flags=abi.PE_CFG_ORDERED_ONLY)
    assert st["place_tasks"] == 0


def test_cfg3_oneoff_full_1m_x_100k():
    """The whole headline tick: 1M one-off tasks x 100k nodes, every placement and the final node totals."""
    run_prefix("big_cfg3_oneoff_1m_100k")


@pytest.mark.parametrize("max_batch", [0, 2048, 12000, 24000])
def test_cfg3_oneoff_batch_sizes(max_batch):
    """Batch size changes how often classes are consumed inside a batch, never a placement."""
    run_prefix("big_cfg3_oneoff_1m_100k", 150_000, max_batch=max_batch)


def test_cfg2_oneoff_full_100k_x_10k():
    run_prefix("big_cfg2_oneoff_100k_10k")


def test_cfg2_oneoff_bigger_batches():
    run_prefix("big_cfg2_oneoff_100k_10k", max_batch=2000)


def test_cfg2_grouped_full_100k_x_10k():
    run_prefix("big_cfg2_grouped_100k_10k")


def test_cfg3_grouped_full_1m_x_100k():
    run_prefix("big_cfg3_grouped_1m_100k")


def test_cfg4_oneoff_60k_x_200k():
    """Full filter chain (resources, generic kinds, host ports, max replicas) at 200k nodes."""
    run_prefix("big_cfg4_oneoff_60k_200k")


def test_cfg5_reschedule_storm_500k_x_100k():
    """BASELINE configs[4]: 10k of 100k nodes drained, their ~500k tasks re-placed in ONE tick (5000 groups of ~100)
    on top of 4.5M running tasks; every placement and the final node totals."""
    run_prefix("big_cfg5_storm_500k_100k")
