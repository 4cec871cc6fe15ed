"""CPU-only checks of the host logic (no GPU): the product's host shim compiled
against the oracle's ABI (oracle/_build/libshim_on_oracle.so, test-only) must
(a) replay the reference's known-answer tests and (b) agree decision-for-decision
with the object-level oracle on random event streams.  This pins the encoded
(flat) oracle to the object-level oracle, and exercises the shim's encoder."""
import inspect

import pytest

from tests import known_answers as KA
from tests.oracle_lib import build_sched, build_shim_on_oracle
from tests.sched_harness import Cluster, JsonScheduler, comparable
from tests.test_mirror_gpu import _rand_node, _rand_task  # generators only; nothing GPU is touched at import

import random


def make_shim():
    return JsonScheduler(build_shim_on_oracle(), "sso")


def make_oracle():
    return JsonScheduler(build_sched(), "so")


SCENARIOS = [(n, f) for n, f in inspect.getmembers(KA, inspect.isfunction) if n.startswith("scenario_")]


@pytest.mark.parametrize("name,fn", SCENARIOS, ids=[n for n, _ in SCENARIOS])
def test_known_answer_through_shim(name, fn):
    params = list(inspect.signature(fn).parameters)
    if "use_spec_version" in params:
        for v in (False, True):
            fn(make_shim, v)
    else:
        fn(make_shim)


@pytest.mark.parametrize("seed", range(24))
def test_random_event_stream_flat_vs_object(seed):
    rng = random.Random(seed)
    n_nodes, n_services = rng.randint(5, 60), rng.randint(2, 8)
    nodes = [_rand_node(rng, i) for i in range(n_nodes)]
    tasks = [_rand_task(rng, i, n_services) for i in range(rng.randint(10, 120))]
    services = [(f"svc{s}", (2 if s % 3 else None)) for s in range(n_services)]
    cm = Cluster(make_shim(), nodes=nodes, tasks=tasks, services=services)
    co = Cluster(make_oracle(), nodes=nodes, tasks=tasks, services=services)
    next_task = len(tasks)
    for step in range(10):
        dm, do = cm.run(), co.run()
        assert comparable(dm, cm) == comparable(do, co), f"seed {seed} step {step}: decisions differ"
        sm, so = cm.s.apply({"op": "device_check"}), co.snapshot()
        assert sm["mismatch"] == []
        assert sm["nodes"] == so["nodes"] and sm["unassigned"] == so["unassigned"]
        for c in (cm, co):
            r2 = random.Random(seed * 100 + step)
            for j in range(r2.randint(3, 25)):
                c.create_task(_rand_task(r2, next_task + j, n_services))
            running = [t for t in c.tasks.values() if t["status"]["state"] == "ASSIGNED"]
            for t in r2.sample(running, min(len(running), r2.randint(0, 6))):
                t = dict(t)
                t["status"] = dict(t["status"], state=r2.choice(["FAILED", "SHUTDOWN", "RUNNING"]))
                c.update_task(t)
            if r2.random() < 0.5:
                c.update_node(_rand_node(r2, r2.randrange(n_nodes)))
            if r2.random() < 0.3:
                c.create_node(_rand_node(r2, n_nodes + step))
            if r2.random() < 0.2 and len(c.nodes) > 3:
                c.delete_node(r2.choice(sorted(c.nodes)))
            c.advance(r2.choice([1, 30, 200]))
        next_task += 25


def test_drain_storm_uploads_only_the_changed_rows():
    from tests.sched_harness import drain_storm_scenario
    drain_storm_scenario(make_shim)


def test_service_ids_are_recycled_under_churn():
    """The engine keeps one dense counter column per service id and never frees one: the shim hands the ids of services
    nobody counts or references any more out again, so the columns are bounded by the services alive at one time."""
    from tests.sched_harness import description, node, resources, task
    nodes = [node(f"n{i:02d}", description=description(resources=resources(64 * 10**9, 2**36))) for i in range(12)]
    cm, co = Cluster(make_shim(), nodes=nodes), Cluster(make_oracle(), nodes=nodes)
    cm.s.apply({"op": "init", "now_ns": cm.now, "nodes": nodes, "tasks": [], "services": [], "svc_recycle_at": 8})
    high = 0
    for rnd in range(40):
        live = [f"svc{rnd}-{j}" for j in range(3)]
        for c in (cm, co):
            for sid in live:
                c.set_service(sid, 1)
                for r in range(4):
                    c.create_task(task(f"t{rnd:02d}-{sid}-{r}", service_id=sid, spec_version=1, reservations=resources(10**8, 2**20)))
        dm, do = cm.run(), co.run()
        assert comparable(dm, cm) == comparable(do, co), f"round {rnd}: decisions differ"
        sm = cm.s.apply({"op": "device_check"})
        assert sm["mismatch"] == [], f"round {rnd}: {sm['mismatch']}"
        assert sm["nodes"] == co.snapshot()["nodes"]
        high = max(high, sm["service_id_high_water"])
        # the services of this round go away: their tasks are shut down and deleted
        for c in (cm, co):
            for t in [t for t in c.tasks.values() if t["service_id"] in live]:
                c.update_task(dict(t, status=dict(t["status"], state="SHUTDOWN")))
                c.delete_task(t["id"])
    assert high <= 16, high          # 120 services came and went; at most a handful of ids were ever alive together


@pytest.mark.parametrize("network_constraints", [False, True])
def test_reference_benchmark_shape(network_constraints):
    """benchScheduler's store contents (scheduler_test.go:3378-3468) at 300 nodes x 600 tasks: nodes with an engine but no
    resource description, every third one with a `network` Network plugin; one-off tasks without a service, with a network
    attachment on that driver in the *Constraints* variants.  Shim and object oracle agree; every task is assigned; the
    spread is level over the eligible nodes."""
    from collections import Counter
    from tests.sched_harness import description, engine, node, task
    nodes = [node(f"n{i:04d}", description=description(engine=engine(plugins=[("Network", "network")] if i % 3 == 0 else []))) for i in range(300)]
    tasks = [task(f"task{i:04d}", networks=["network"] if network_constraints else ()) for i in range(600)]
    out = []
    for mk in (make_shim, make_oracle):
        c = Cluster(mk(), nodes=nodes)
        for t in tasks:
            c.create_task(t)
        out.append(c.run())
    assert out[0] == out[1]
    per_node = Counter(d["node_id"] for d in out[0].values())
    assert all(d["state"] == "ASSIGNED" for d in out[0].values())
    eligible = [n["id"] for i, n in enumerate(nodes) if i % 3 == 0] if network_constraints else [n["id"] for n in nodes]
    assert set(per_node) <= set(eligible) and len(per_node) == len(eligible) and max(per_node.values()) - min(per_node.values()) <= 1
