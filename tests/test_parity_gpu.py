"""GPU parity: the CUDA engine against the CPU oracle through the same C ABI,
on the same seeded inputs -- placements bit-exact, failure counters, and the
whole node mirror afterwards."""
import numpy as np
import pytest

from swarmkit_b200 import PlacementEngine, abi, workloads as W
from tests import randwork as R
from tests.oracle_lib import OracleEngine

pytestmark = pytest.mark.gpu


def run_both(nodes, tick, n_nodes, flags=0, max_batch=0, ticks=None):
    gpu = PlacementEngine(node_capacity=n_nodes, flags=flags, max_batch=max_batch)
    cpu = OracleEngine(node_capacity=n_nodes)
    for e in (gpu, cpu):
        e.node_upsert(nodes)
        e.set_node_count(n_nodes)
    results = []
    for t in (ticks or [tick]):
        results.append((t, gpu.schedule(t), cpu.schedule(t)))
    return gpu, cpu, results


@pytest.mark.parametrize("flags", [0, abi.PE_CFG_NO_SPECULATION])
def test_cfg1_swarm_bench_shape(flags):
    w = W.cfg1()
    gpu, cpu, res = run_both(w.nodes, w.tick, w.n_nodes, flags)
    for t, a, b in res:
        R.compare_results(t, a, b, "cfg1")
    assert (np.bincount(res[0][1][0], minlength=100) == 10).all()   # 1000 identical replicas over 100 nodes


@pytest.mark.parametrize("mode", ["grouped", "oneoff"])
@pytest.mark.parametrize("flags", [0, abi.PE_CFG_NO_SPECULATION])
def test_cfg2_small(mode, flags):
    n_tasks = 4000 if flags else 20000
    w = W.cfg2(mode, n_nodes=2000, n_tasks=n_tasks, n_services=20)
    gpu, cpu, res = run_both(w.nodes, w.tick, w.n_nodes, flags)
    for t, a, b in res:
        R.compare_results(t, a, b, f"cfg2-{mode}")
    R.compare_state(gpu, cpu, w.n_nodes, 20 if mode == "grouped" else 1, 0, 0, f"cfg2-{mode}")


@pytest.mark.parametrize("mode", ["grouped", "oneoff"])
def test_cfg2_tight_resources(mode):
    # demand exceeds capacity: tasks go unplaced, failure counters must match
    w = W.cfg2(mode, n_nodes=300, n_tasks=12000, n_services=12)
    gpu, cpu, res = run_both(w.nodes, w.tick, w.n_nodes)
    for t, a, b in res:
        R.compare_results(t, a, b, f"cfg2-tight-{mode}")
        assert (a[0] == abi.PE_NONE).any()
    R.compare_state(gpu, cpu, w.n_nodes, 12 if mode == "grouped" else 1, 0, 0, "cfg2-tight")


@pytest.mark.parametrize("mode", ["grouped", "oneoff"])
def test_cfg3_small(mode):
    w = W.cfg3(mode, n_nodes=20000, n_tasks=30000, n_services=100)
    gpu, cpu, res = run_both(w.nodes, w.tick, w.n_nodes)
    for t, a, b in res:
        R.compare_results(t, a, b, f"cfg3-{mode}")
    R.compare_state(gpu, cpu, w.n_nodes, 100, 0, 0, f"cfg3-{mode}")
    if mode == "oneoff":
        st = gpu.stats()
        assert st["scan_launches"] > 0 and st["fast_path"] > 0.5 * 30000


def test_cfg3_rotated_ties():
    w = W.cfg3("oneoff", n_nodes=10000, n_tasks=20000, n_services=50)
    rng = np.random.default_rng(7)
    w.tick.groups["tie_start"] = rng.integers(0, w.n_nodes, w.tick.n_groups)
    gpu, cpu, res = run_both(w.nodes, w.tick, w.n_nodes)
    for t, a, b in res:
        R.compare_results(t, a, b, "cfg3-rot")


@pytest.mark.parametrize("mode", ["grouped", "oneoff"])
def test_cfg4_small(mode):
    w = W.cfg4(mode, n_nodes=20000, n_tasks=20000, n_services=100)
    gpu, cpu, res = run_both(w.nodes, w.tick, w.n_nodes)
    for t, a, b in res:
        R.compare_results(t, a, b, f"cfg4-{mode}")
    R.compare_state(gpu, cpu, w.n_nodes, 100, 3, 64, f"cfg4-{mode}")


@pytest.mark.parametrize("seed", range(12))
def test_random_all_filters(seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(40, 1500))
    nodes = R.random_nodes(rng, n, tight=bool(seed % 2))
    ticks = [R.random_tick(rng, n, int(rng.integers(20, 200))) for _ in range(3)]
    flags = abi.PE_CFG_NO_SPECULATION if seed % 4 == 3 else 0
    gpu, cpu, res = run_both(nodes, None, n, flags=flags, max_batch=(64 if seed % 3 == 0 else 0), ticks=ticks)
    for i, (t, a, b) in enumerate(res):
        R.compare_results(t, a, b, f"random seed {seed} tick {i}")
    R.compare_state(gpu, cpu, n, 4, 3, 40, f"random seed {seed}")


@pytest.mark.parametrize("seed", range(4))
def test_random_oneoff_heavy(seed):
    # long runs of k=1 groups so the batched scan + fast path carry the load
    rng = np.random.default_rng(2000 + seed)
    n = int(rng.integers(500, 6000))
    nodes = R.random_nodes(rng, n)
    tick = R.random_tick(rng, n, 3000, p_oneoff=1.0, feature_p=0.25, rotate=bool(seed % 2))
    gpu, cpu, res = run_both(nodes, tick, n)
    for t, a, b in res:
        R.compare_results(t, a, b, f"oneoff-heavy seed {seed}")
    R.compare_state(gpu, cpu, n, 4, 3, 40, f"oneoff-heavy seed {seed}")
    assert gpu.stats()["scan_launches"] > 0


@pytest.mark.parametrize("seed", range(6))
def test_random_leaf_visits(seed):
    """Groups that are one leaf visit of a placement-preference tree (pe_group.leaf_cnt): the node set is cut down to
    the leaf (nodeset.go:59-101) on every path a group can take (one task: one-CTA path; k > 1: k_groups)."""
    rng = np.random.default_rng(3000 + seed)
    n = int(rng.integers(60, 3000))
    nodes = R.random_nodes(rng, n, tight=bool(seed % 2))
    ticks = [R.random_tick(rng, n, int(rng.integers(20, 150)), p_leaf=0.4, p_oneoff=0.4) for _ in range(3)]
    assert any((t.groups["leaf_cnt"] > 0).any() for t in ticks)
    gpu, cpu, res = run_both(nodes, None, n, flags=(abi.PE_CFG_ORDERED_ONLY if seed == 5 else 0), ticks=ticks)
    for i, (t, a, b) in enumerate(res):
        R.compare_results(t, a, b, f"leaf visits seed {seed} tick {i}")
    R.compare_state(gpu, cpu, n, 4, 3, 40, f"leaf visits seed {seed}")


@pytest.mark.parametrize("seed", range(3))
def test_pref_leaves(seed):
    """nodeSet.tree's leaves and their task sums (pe_pref_leaves) against the oracle, before and after a tick."""
    rng = np.random.default_rng(3100 + seed)
    n = int(rng.integers(50, 5000))
    nodes = R.random_nodes(rng, n)
    tick = R.random_tick(rng, n, 60)
    gpu, cpu, res = run_both(nodes, tick, n)

    def leaves(e, svc, cols):
        vals, tasks = e.pref_leaves(svc, cols, cap=n + 1)
        return sorted((tuple(v.tolist()), int(t)) for v, t in zip(vals, tasks))

    first = abi.PE_ATTR_FIRST_LABEL
    for svc in range(4):
        for cols in ([first], [first + 1, first], [first, first + 1, first + 2], [abi.PE_ATTR_ROLE, first + 2], []):
            a, b = leaves(gpu, svc, cols), leaves(cpu, svc, cols)
            assert a == b, f"seed {seed} service {svc} levels {cols}"
            assert len(a) <= 4 ** max(len(cols), 1)
    with pytest.raises(abi.EngineError):
        gpu.pref_leaves(0, [first, first + 1], cap=2)       # more leaves than the caller's buffers hold


@pytest.mark.parametrize("seed", range(3))
def test_match_matrix(seed):
    """The (service, node) predicate matrix for the constraint enforcer / global orchestrator (pe_match_matrix):
    constraint.NodeMatches alone, and all four node-attribute filters, against the oracle."""
    rng = np.random.default_rng(3200 + seed)
    n = int(rng.integers(40, 4000))
    nodes = R.random_nodes(rng, n)
    tick = R.random_tick(rng, n, 80, feature_p=0.6)
    gpu, cpu = PlacementEngine(node_capacity=n), OracleEngine(node_capacity=n)
    for e in (gpu, cpu):
        e.node_upsert(nodes)
        e.set_node_count(n)
    for only_constraints in (True, False):
        t = tick.slice_groups(0, tick.n_groups)
        g = t.groups.copy()
        if only_constraints:
            g["filter_mask"] &= 1 << abi.PE_F_CONSTRAINT
        t.groups = g
        a, b = gpu.match_matrix(t, n), cpu.match_matrix(t, n)
        assert a.shape == (tick.n_groups, n) and (a == b).all(), f"seed {seed} only_constraints={only_constraints}: {(a != b).sum()} pairs differ"
        assert 0 < a.sum() < a.size


def test_large_group_global_path():
    # k > 2048 candidates: the sequencer's global-memory sort / fill path
    w = W.cfg1(n_nodes=5000, n_tasks=12000)
    gpu, cpu, res = run_both(w.nodes, w.tick, w.n_nodes)
    for t, a, b in res:
        R.compare_results(t, a, b, "large-group")


def test_fit_and_deltas():
    rng = np.random.default_rng(5)
    n = 400
    nodes = R.random_nodes(rng, n)
    tick = R.random_tick(rng, n, 150, p_oneoff=1.0, rotate=False)
    idx = rng.integers(0, n + 5, tick.n_groups).astype(np.uint32)
    gpu = PlacementEngine(node_capacity=n)
    cpu = OracleEngine(node_capacity=n)
    for e in (gpu, cpu):
        e.node_upsert(nodes)
        e.set_node_count(n)
    a_ok, a_f = gpu.fit(tick, idx)
    b_ok, b_f = cpu.fit(tick, idx)
    assert (a_ok == b_ok).all() and (a_f == b_f).all()
    d = np.zeros(50, abi.task_delta_dt)
    d["node_idx"] = rng.integers(0, n, 50)
    d["svc_id"] = rng.integers(0, 4, 50)
    d["sign"] = rng.choice([-1, 1], 50)
    d["counts"] = 1
    d["cpu"] = rng.integers(0, 3, 50) * 10**9
    d["mem"] = rng.integers(0, 3, 50) * 2**30
    # keep counters non-negative: only remove where something is there
    tot = cpu.snapshot(0, n)["total_tasks"]
    for i in range(50):
        s = cpu.snapshot_service(int(d[i]["svc_id"]), 0, n)
        if d[i]["sign"] < 0 and (tot[d[i]["node_idx"]] == 0 or s[d[i]["node_idx"]] == 0):
            d[i]["sign"] = 1
    # make node indices unique so the order of application cannot matter for the check
    _, first = np.unique(d["node_idx"], return_index=True)
    d = d[np.sort(first)]
    for e in (gpu, cpu):
        e.node_task_delta(d)
        e.node_remove(np.array([3, 7], np.uint32))
    R.compare_state(gpu, cpu, n, 4, 3, 40, "after fit+delta")
    t2 = R.random_tick(rng, n, 100)
    R.compare_results(t2, gpu.schedule(t2), cpu.schedule(t2), "after fit+delta")


def test_engine_requires_extension():
    # the product binding must be backed by the in-tree CUDA library
    import os
    from swarmkit_b200 import engine_library_path
    assert os.path.exists(engine_library_path())


def test_cfg3_many_nodes_touched_bitmap_in_global_memory():
    # more than 12288 bitmap words: the sequencer keeps `touched` in global memory instead of shared memory
    w = W.cfg3("oneoff", n_nodes=450_000, n_tasks=3000, n_services=30)
    gpu, cpu, res = run_both(w.nodes, w.tick, w.n_nodes)
    for t, a, b in res:
        R.compare_results(t, a, b, "cfg3-450k-nodes")
    st = gpu.stats()
    assert st["scan_launches"] > 0 and st["fast_path"] > 0


def test_cfg3_more_nodes_than_the_placement_steps_shared_bitmap_holds():
    # more than 18432 bitmap words (~590k nodes): k_place keeps no copy of the touched bitmap in shared memory and asks the
    # chunk's set of taken nodes for every candidate (the BM = false instantiation of its resolve phase)
    w = W.cfg3("oneoff", n_nodes=620_000, n_tasks=2500, n_services=25)
    gpu, cpu, res = run_both(w.nodes, w.tick, w.n_nodes)
    for t, a, b in res:
        R.compare_results(t, a, b, "cfg3-620k-nodes")
    st = gpu.stats()
    assert st["scan_launches"] > 0 and st["place_tasks"] > 0
