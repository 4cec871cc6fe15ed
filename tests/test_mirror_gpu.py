"""The drop-in check: the product's host shim (C++ mirror of the Go scheduler,
swarmkit_b200/csrc/scheduler_host.cpp) driving the CUDA engine through the C
ABI, replaying the reference's known-answer tests and, on random event streams,
agreeing decision-for-decision with the object-level CPU oracle."""
import inspect
import random

import pytest

from swarmkit_b200 import _build
from tests import known_answers as KA
from tests.oracle_lib import build_sched
from tests.sched_harness import (Cluster, JsonScheduler, comparable, description, discrete, engine, host_port, named, node, placement,
                                 resources, task)

pytestmark = pytest.mark.gpu

# (round 1 refused placement preferences; they are walked leaf by leaf now: scheduler_host.cpp::schedulePreferenceGroup)
UNSUPPORTED = set()


def make_mirror():
    return JsonScheduler(_build.build_scheduler_shim(), "ss")


def make_oracle():
    return JsonScheduler(build_sched(), "so")


SCENARIOS = [(n, f) for n, f in inspect.getmembers(KA, inspect.isfunction) if n.startswith("scenario_")]


@pytest.mark.parametrize("name,fn", SCENARIOS, ids=[n for n, _ in SCENARIOS])
def test_known_answer_on_gpu(name, fn):
    params = list(inspect.signature(fn).parameters)
    if "use_spec_version" in params:
        for v in (False, True):
            fn(make_mirror, v)
    else:
        fn(make_mirror)


def _rand_node(rng, i):
    labels = {k: rng.choice(["a", "b", "c", "A"]) for k in ("zone", "disk") if rng.random() < 0.8}
    plugins = [("Volume", rng.choice(["p1", "p2", "p1:latest"])), ("Log", rng.choice(["default", "json"]))] if rng.random() < 0.7 else []
    gen = []
    if rng.random() < 0.6:
        gen.append(discrete("apple", rng.randint(0, 6)))
    if rng.random() < 0.5:
        gen += [named("gpu", f"g{j}") for j in range(rng.randint(1, 4))]
    desc = description(hostname=f"host-{i % 7}", platform={"os": rng.choice(["linux", "windows"]), "arch": rng.choice(["amd64", "x86_64", "arm64", "aarch64"])}
                       if rng.random() < 0.85 else None,
                       resources=resources(rng.randint(1, 8) * 10**9, rng.randint(1, 16) * 2**30, gen) if rng.random() < 0.9 else None,
                       engine=engine(labels={"os": rng.choice(["ubuntu", "Ubuntu", "rhel"])}, plugins=plugins) if rng.random() < 0.8 else None)
    return node(f"n{i:04d}", state=rng.choice(["READY"] * 8 + ["DOWN", "UNKNOWN"]), availability=rng.choice(["ACTIVE"] * 8 + ["DRAIN", "PAUSE"]),
                labels=labels if rng.random() < 0.9 else None, role=rng.choice(["WORKER", "WORKER", "MANAGER"]),
                addr=rng.choice(["10.0.0.%d" % (i % 250), "10.1.%d.7" % (i % 4), "2001:db8::%x" % i, "", "bogus"]), description=desc)


def _rand_task(rng, i, n_services):
    s = rng.randrange(n_services)
    rs = random.Random(1000 + s)   # everything a filter reads is a function of the service (one spec per group)
    cons = []
    for _ in range(rs.randint(0, 3)):
        cons.append(rs.choice(["node.labels.zone==a", "node.labels.zone!=b", "node.labels.disk == C", "engine.labels.os==ubuntu",
                               "node.role==worker", "node.role != manager", "node.platform.os==linux", "node.platform.arch!=arm64",
                               "node.hostname!=host-3", "node.ip==10.0.0.0/24", "node.ip!=10.1.2.7", "node.ip==2001:db8::/64",
                               "node.bogus==x", "node.id!=n0003"]))
    res = resources(rs.choice([0, 5, 10]) * 10**8, rs.choice([0, 1, 2]) * 2**29,
                    ([discrete("apple", rs.randint(0, 2))] if rs.random() < 0.3 else []) + ([discrete("gpu", rs.randint(0, 2))] if rs.random() < 0.3 else [])) \
        if rs.random() < 0.6 else None
    prefs = rs.choice([[], [], [], ["node.labels.zone"], ["node.labels.zone", "engine.labels.os"], ["Node.Labels.disk", "node.hostname", "node.labels.zone"]])
    pl = placement(constraints=cons, platforms=rs.choice([[], [], [("amd64", "linux")], [("", "linux"), ("arm64", "")], [("aarch64", "linux")]]),
                   max_replicas=rs.choice([0, 0, 1, 2]), preferences=prefs)
    ports = rs.choice([None, None, None, [host_port(rs.choice([80, 443]), rs.choice(["TCP", "UDP"]))]])
    mounts = rs.choice([[], [], [{"type": "VOLUME", "driver": "p1"}], [{"type": "BIND", "driver": None}]])
    return task(f"t{i:05d}", service_id=f"svc{s}", spec_version=(1 + s % 2) if s % 3 else None, reservations=res, placement=pl, ports=ports,
                mounts=mounts, log_driver=rs.choice([None, None, "json", "none"]),
                desired_state=rng.choice(["RUNNING"] * 12 + ["SHUTDOWN"]))


@pytest.mark.parametrize("seed", range(6))
def test_random_event_stream_matches_oracle(seed):
    rng = random.Random(seed)
    n_nodes, n_services = rng.randint(5, 60), rng.randint(2, 8)
    nodes = [_rand_node(rng, i) for i in range(n_nodes)]
    tasks = [_rand_task(rng, i, n_services) for i in range(rng.randint(10, 120))]
    services = [(f"svc{s}", (2 if s % 3 else None)) for s in range(n_services)]
    cm = Cluster(make_mirror(), nodes=nodes, tasks=tasks, services=services)
    co = Cluster(make_oracle(), nodes=nodes, tasks=tasks, services=services)
    next_task = len(tasks)
    for step in range(12):
        dm, do = cm.run(), co.run()
        assert comparable(dm, cm) == comparable(do, co), f"seed {seed} step {step}: decisions differ"
        sm, so = cm.s.apply({"op": "device_check"}), co.snapshot()
        assert sm["mismatch"] == [], f"seed {seed} step {step}: device mirror diverged on {sm['mismatch']}"
        for a, b in zip(sm["nodes"], so["nodes"]):
            assert a == b, f"seed {seed} step {step}: NodeInfo differs for {a['id']}"
        assert sm["unassigned"] == so["unassigned"]
        # mutate: new tasks, failures, deletions, node churn -- the same events to both
        for c in (cm, co):
            r2 = random.Random(seed * 100 + step)
            for _ in range(r2.randint(3, 25)):
                c.create_task(_rand_task(r2, next_task + _, n_services))
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


def test_commit_failure_rolls_back():
    # scheduler.go:472-487: a decision the store rejects is undone and retried
    nodes = [node("a", description=description(resources=resources(2 * 10**9, 2**31))), node("b", description=description(resources=resources(2 * 10**9, 2**31)))]
    tasks = [task(f"t{i}", service_id="s", spec_version=1, reservations=resources(10**9, 2**30)) for i in range(3)]
    for make in (make_mirror, make_oracle):
        c = Cluster(make(), nodes=nodes, tasks=tasks, services=[("s", 1)])
        d = c.tick(fail_commit=["t1"])
        assert "t1" not in d and len(d) == 2
        snap = c.snapshot()
        assert sum(n["active_tasks"] for n in snap["nodes"]) == 2 and "t1" in snap["unassigned"]
        d = c.tick()
        assert d["t1"]["state"] == "ASSIGNED"
    # and the device mirror followed the rollback
    c = Cluster(make_mirror(), nodes=nodes, tasks=tasks, services=[("s", 1)])
    c.tick(fail_commit=["t1"])
    assert c.s.apply({"op": "device_check"})["mismatch"] == []


def test_drain_storm_uploads_only_the_changed_rows():
    """cfg5's flow through the shim on the GPU: the drained nodes' rows are the only ones that cross the ABI again."""
    from tests.sched_harness import drain_storm_scenario
    drain_storm_scenario(make_mirror)
