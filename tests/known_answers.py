"""The reference's own known-answer tests for the scheduler hot path, replayed
through the JSON event protocol.  Each function cites the test it ports
(manager/scheduler/scheduler_test.go unless noted) and asserts exactly what that
test asserts (count distributions / "one of the tied nodes" / error strings).
They run against the CPU oracle (pinning it) and against the product's host
shim + CUDA engine (the drop-in check)."""
from __future__ import annotations

from tests.sched_harness import (Cluster, count_by_node, description, discrete, engine, host_port, named, node,
                                 placement, resources, task)


def _ready_nodes(ids, **kw):
    return [node(i, **kw) for i in ids]


# scheduler_test.go:22-369 TestScheduler
def scenario_scheduler(make):
    c = Cluster(make(), nodes=_ready_nodes(["id1", "id2", "id3"]),
                tasks=[task("id1", state="ASSIGNED", node_id="id1"), task("id2"), task("id3")], services=[""])
    a = c.assignments(c.run())
    assert a["id2"] in ("id2", "id3") and a["id3"] in ("id2", "id3") and a["id2"] != a["id3"]      # :125-133
    for n in list(c.nodes.values()):
        c.update_node(n)                                                                            # :135-143
    c.delete_task("id1")
    c.create_task(task("id4"))
    assert c.assignments(c.run())["id4"] == "id1"                                                   # :167
    c.update_task(task("id4"))                                                                      # unassign -> rescheduled
    assert c.assignments(c.run())["id4"] == "id1"                                                   # :190
    c.create_node(node("removednode", state="DOWN"))
    c.delete_node("removednode")
    c.create_task(task("removednode"))
    assert c.assignments(c.run())["removednode"] != "removednode"                                   # :226
    c.create_node(node("id4"))
    c.create_task(task("id5"))
    assert c.assignments(c.run())["id5"] == "id4"                                                   # :261
    c.create_node(node("id5", state="DOWN"))
    c.create_task(task("id6"))
    assert c.assignments(c.run())["id6"] != "id5"                                                   # :296
    c.update_node(node("id5"))
    c.create_task(task("id7"))
    assert c.assignments(c.run())["id7"] == "id5"                                                   # :331
    c.create_node(node("id6"))
    c.update_node(node("id6", state="DOWN"))
    c.create_task(task("id8"))
    assert c.assignments(c.run())["id8"] != "id6"                                                   # :368


# scheduler_test.go:371-653 TestHA
def scenario_ha(make, use_spec_version):
    sv = 1 if use_spec_version else None
    t1 = lambda i: task(f"t1id{i}", service_id="service1", spec_version=sv)   # noqa: E731
    t2 = lambda i: task(f"t2id{i}", service_id="service2", spec_version=sv)   # noqa: E731
    c = Cluster(make(), nodes=_ready_nodes([f"id{i}" for i in range(1, 6)]), tasks=[t1(i) for i in range(18)],
                services=["service1", "service2"])
    t1a = count_by_node(c.assignments(c.run()), "t1")
    assert len(t1a) == 5 and sorted(t1a.values()) == [3, 3, 4, 4, 4]                                # :481-496
    for i in range(2):
        c.create_task(t2(i))
    t2a = count_by_node(c.assignments(c.run()), "t2")
    assert len(t2a) == 2 and all(t1a[n] == 3 for n in t2a)                                          # :520-524
    for i in range(18, 21):
        c.create_task(t1(i))
    new = c.assignments(c.run())
    shared = []
    for tid, nid in new.items():
        assert t1a[nid] != 5                                                                        # :541-543
        t1a[nid] += 1
        if t2a.get(nid, 0):
            shared.append(nid)
    assert len(shared) == 2 and shared[0] != shared[1]                                              # :556-558
    assert sorted(t1a.values()) == [4, 4, 4, 4, 5]                                                  # :576-577
    c.create_task(t2(4))
    a = c.assignments(c.run())
    assert t2a.get(a["t2id4"], 0) == 0 and t1a[a["t2id4"]] != 5                                     # :595-600
    t2a[a["t2id4"]] = 1
    for tid in [k for k, t in c.tasks.items() if t["node_id"] == "id1"]:
        c.delete_task(tid)
    t1a["id1"] = 0
    t2a["id1"] = 0
    for i in range(22, 26):
        c.create_task(t1(i))
    for i in range(5, 7):
        c.create_task(t2(i))
    a = c.assignments(c.run())
    assert count_by_node(a, "t1").get("id1", 0) == 4 and count_by_node(a, "t2").get("id1", 0) == 1  # :646-647


# scheduler_test.go:655-806 TestPreferences
def scenario_preferences(make, use_spec_version):
    sv = 1 if use_spec_version else None
    nodes = [node("id1", labels={"az": "az1"})] + [node(f"id{i}", labels={"az": "az2"}) for i in range(2, 6)]
    tasks = [task(f"t1id{i}", service_id="service1", spec_version=sv, placement=placement(preferences=["node.labels.az"]))
             for i in range(8)]
    c = Cluster(make(), nodes=nodes, tasks=tasks, services=["service1"])
    a = count_by_node(c.assignments(c.run()), "t1")
    assert a == {"id1": 4, "id2": 1, "id3": 1, "id4": 1, "id5": 1}                                  # :795-800


def _mp_nodes():
    def n(i, az, rack, mem, apples):
        return node(f"id{i}", labels={"az": az, "rack": rack},
                    description=description(resources=resources(1e9, mem, [discrete("apple", apples)])))
    return [n(0, "az1", "rack1", 1e8, 1), n(1, "az1", "rack1", 1e9, 10), n(2, "az2", "rack1", 1e9, 6), n(3, "az2", "rack1", 1e9, 6),
            n(4, "az2", "rack1", 1e9, 6), n(5, "az2", "rack2", 1e9, 6), n(6, "az2", "rack2", 1e9, 6)]


# scheduler_test.go:808-1111 TestMultiplePreferences
def scenario_multiple_preferences(make, use_spec_version):
    sv = 1 if use_spec_version else None
    tasks = [task(f"t1id{i}", service_id="service1", spec_version=sv,
                  placement=placement(preferences=["node.labels.az", "node.labels.rack"]),
                  reservations=resources(0, 2e8, [discrete("apple", 2)])) for i in range(12)]
    c = Cluster(make(), nodes=_mp_nodes(), tasks=tasks, services=["service1"])
    a = count_by_node(c.assignments(c.run()), "t1")
    g = lambda k: a.get(k, 0)   # noqa: E731
    assert g("id0") == 0 and g("id1") == 5                                                          # :1063,1068
    r1 = g("id2") + g("id3") + g("id4")
    if r1 == 4:
        assert sorted([g("id2"), g("id3"), g("id4")]) == [1, 1, 2] and sorted([g("id5"), g("id6")]) == [1, 2]
    elif r1 == 3:
        assert [g("id2"), g("id3"), g("id4"), g("id5"), g("id6")] == [1, 1, 1, 2, 2]
    else:
        raise AssertionError(f"unexpected task layout {a}")


# scheduler_test.go:1115-1261 TestMultiplePreferencesScaleUp (termination + id12+id21 == 2)
def scenario_multiple_preferences_scale_up(make):
    pl = placement(preferences=["node.labels.az", "node.labels.rack"])
    nodes = [node("id11", labels={"az": "dc1", "rack": "r1"}), node("id12", labels={"az": "dc1", "rack": "r2"}),
             node("id21", labels={"az": "dc2", "rack": "r1"})]
    tasks = [task(f"t1id{i}", service_id="service1", spec_version=1, placement=pl) for i in range(2)]
    for nid, cnt in (("id11", 3), ("id12", 1), ("id21", 3)):
        for i in range(cnt):
            tasks.append(task(f"t1running-{nid}-{i}", service_id="service1", spec_version=1, placement=pl, node_id=nid,
                              state="RUNNING"))
    c = Cluster(make(), nodes=nodes, tasks=tasks, services=["service1"])
    a = count_by_node(c.assignments(c.run()), "t1id")
    assert sum(a.values()) == 2 and a.get("id12", 0) + a.get("id21", 0) == 2                        # :1254,1260


# scheduler_test.go:1263-1323 TestSchedulerNoReadyNodes
def scenario_no_ready_nodes(make):
    c = Cluster(make(), tasks=[task("id1", service_id="serviceID1")], services=["serviceID1"])
    d = c.run()
    assert d["id1"]["err"] == "no suitable node" and d["id1"]["state"] == "PENDING"                 # :1300
    c.create_node(node("newnode"))
    assert c.assignments(c.run())["id1"] == "newnode"                                               # :1322


# scheduler_test.go:1325-1458 TestSchedulerFaultyNode
def scenario_faulty_node(make):
    rep = lambda tid, **kw: task(tid, service_id="service1", **kw)                                  # noqa: E731
    pre = lambda tid, **kw: task(tid, service_id="service2", node_id="id1", **kw)                   # noqa: E731
    c = Cluster(make(), nodes=_ready_nodes(["id1", "id2"]),
                tasks=[rep("id1", node_id="id1", state="RUNNING"), pre("id2", state="RUNNING")],
                services=["service1", "service2"])
    c.run()

    def failures(nid, key):
        for n in c.snapshot()["nodes"]:
            if n["id"] == nid:
                return n["failures"].get(key, 0)
        raise KeyError(nid)

    for i in range(8):
        rid, pid = f"rep{i:02d}", f"pre{i:02d}"
        c.create_task(rep(rid))
        a = c.assignments(c.run())
        assert a[rid] == ("id2" if i < 5 else "id1")                                                # :1423-1431
        assert failures("id2", "service1@0") == min(i, 5)                                           # :1435-1439
        assert failures("id1", "service1@0") == max(i - 5, 0)                                       # :1443-1448
        c.create_task(pre(pid))
        a = c.assignments(c.run())
        assert a[pid] == "id1"                                                                      # :1461
        assert failures("id1", "service2@0") == 0                                                   # :1467
        for tid in (rid, pid):
            t = dict(c.tasks[tid])
            t["status"] = dict(t["status"], state="FAILED")
            c.update_task(t)


# scheduler_test.go:1484-1618 TestSchedulerFaultyNodeSpecVersion
def scenario_faulty_node_spec_version(make):
    c = Cluster(make(), nodes=_ready_nodes(["id1", "id2"]),
                tasks=[task("id1", service_id="service1", spec_version=1, node_id="id1", state="RUNNING")],
                services=[("service1", 1)])
    c.run()

    def failures(nid, ver):
        for n in c.snapshot()["nodes"]:
            if n["id"] == nid:
                return n["failures"].get(f"service1@{ver}", 0)
        raise KeyError(nid)

    for i in range(15):
        tid = f"new{i:02d}"
        c.create_task(task(tid, service_id="service1", spec_version=2 if i > 5 else 1))
        a = c.assignments(c.run())
        assert a[tid] == ("id2" if (i < 5 or 5 < i < 11) else "id1")                                # :1575-1585
        e11, e12, e21, e22 = 0, 0, i, 0
        if i > 5:
            e11, e21, e22 = 1, 5, i - 6
        if i > 11:
            e12, e22 = i - 11, 5
        assert (failures("id1", 1), failures("id1", 2), failures("id2", 1), failures("id2", 2)) == (e11, e12, e21, e22)  # :1604-1607
        t = dict(c.tasks[tid])
        t["status"] = dict(t["status"], state="FAILED")
        c.update_task(t)


def _rc_node(nid, state, cpu, mem, oranges, apples):
    gen = [named("orange", o) for o in oranges] + [discrete("apple", apples)]
    return node(nid, state=state, description=description(resources=resources(cpu, mem, gen)))


# scheduler_test.go:1620-1778 TestSchedulerResourceConstraint
def scenario_resource_constraint(make):
    nodes = [_rc_node("underprovisioned", "READY", 1e9, 1e9, ["blue"], 1),
             _rc_node("nonready1", "UNKNOWN", 2e9, 2e9, ["blue", "red"], 2),
             _rc_node("nonready2", "UNKNOWN", 2e9, 2e9, ["blue", "red"], 2)]
    t = task("id1", service_id="serviceID1", reservations=resources(0, 2e9, [discrete("orange", 2), discrete("apple", 2)]))
    c = Cluster(make(), nodes=nodes, tasks=[t], services=["serviceID1"])
    d = c.run()
    assert d["id1"]["err"] == "no suitable node (2 nodes not available for new tasks; insufficient resources on 1 node)"  # :1745
    c.create_node(_rc_node("bignode", "READY", 4e9, 8e9, ["blue", "red", "green"], 4))
    assert c.assignments(c.run())["id1"] == "bignode"                                               # :1777


# scheduler_test.go:1780-1920 TestSchedulerResourceConstraintHA
def scenario_resource_constraint_ha(make):
    res = resources(0, 5e8, [discrete("apple", 1)])
    nodes = [node("id1", description=description(resources=resources(0, 1e9, [discrete("apple", 2)]))),
             node("id2", description=description(resources=resources(0, 1e11, [discrete("apple", 5)])))]
    tasks = [task("id1", node_id="id1", state="RUNNING", reservations=res)]
    tasks += [task(f"id{i}", node_id="id2", state="RUNNING", reservations=res) for i in (2, 3, 4)]
    tasks += [task("id5", reservations=res), task("id6", reservations=res)]
    c = Cluster(make(), nodes=nodes, tasks=tasks, services=[""])
    a = c.assignments(c.run())
    assert sorted([a["id5"], a["id6"]]) == ["id1", "id2"]                                           # :1915-1919


# scheduler_test.go:1922-2026 TestSchedulerResourceConstraintDeadTask
def scenario_resource_constraint_dead_task(make):
    res = resources(0, 8e8, [discrete("apple", 3)])
    n = node("id1", description=description(resources=resources(1e9, 1e9, [discrete("apple", 4)])))
    c = Cluster(make(), nodes=[n], tasks=[task("id1", service_id="serviceID1", reservations=res)], services=["serviceID1"])
    assert c.assignments(c.run()) == {"id1": "id1"}                                                 # :1996-1997
    c.create_task(task("id2", service_id="serviceID1", reservations=res))
    d = c.run()
    assert d["id2"]["err"] == "no suitable node (insufficient resources on 1 node)"                 # :2011
    t = dict(c.tasks["id1"])
    t["status"] = dict(t["status"], state="SHUTDOWN")
    c.update_task(t)
    assert c.assignments(c.run()) == {"id2": "id1"}                                                 # :2024-2025


# scheduler_test.go:2028-2110 TestSchedulerPreexistingDeadTask
def scenario_preexisting_dead_task(make):
    res = resources(0, 8e8, [discrete("apple", 1)])
    n = node("id1", description=description(resources=resources(1e9, 1e9, [discrete("apple", 1)])))
    c = Cluster(make(), nodes=[n], tasks=[task("id1", node_id="id1", state="SHUTDOWN", reservations=res)], services=[""])
    c.run()
    c.create_task(task("id2", node_id="", reservations=res))
    assert c.assignments(c.run()) == {"id2": "id1"}                                                 # :2108-2109


# scheduler_test.go:2112-2343 TestSchedulerCompatiblePlatform
def scenario_compatible_platform(make):
    plat = lambda arch, os_: {"os": os_, "arch": arch}                                              # noqa: E731
    nodes = [node("node1", description=description(platform=plat("x86_64", "linux"))),
             node("node2", description=description(platform=plat("amd64", "windows"))),
             node("node3", description=description())]
    t = lambda tid, plats=None: task(tid, service_id="serviceID1",                                 # noqa: E731
                                     placement=placement(platforms=plats) if plats is not None else None)
    c = Cluster(make(), nodes=nodes, tasks=[t("id1", [("amd64", "linux")])], services=["serviceID1"])
    assert c.assignments(c.run())["id1"] == "node1"                                                 # :2305
    c.create_task(t("id2", [("arm", "linux")]))
    assert c.run()["id2"]["err"] == "no suitable node (unsupported platform on 3 nodes)"           # :2314
    c.create_task(t("id3"))
    assert c.assignments(c.run())["id3"] in ("node2", "node3")                                      # :2323
    c.create_task(t("id4", [("", "linux")]))
    assert c.assignments(c.run())["id4"] == "node1"                                                 # :2332
    c.create_task(t("id5", [("amd64", "linux"), ("x86_64", "windows")]))
    assert c.assignments(c.run())["id5"] in ("node1", "node2")                                      # :2341


# scheduler_test.go:2345-2422 TestSchedulerUnassignedMap
def scenario_unassigned_map(make):
    n = node("node1", description=description(platform={"os": "linux", "arch": "x86_64"}))
    c = Cluster(make(), nodes=[n], tasks=[task("id1", service_id="serviceID1", placement=placement(platforms=[("amd64", "windows")]))],
                services=["serviceID1"])
    c.tick()
    assert "id1" in c.snapshot()["unassigned"]                                                      # :2407
    c.s.apply({"op": "delete_service", "id": "serviceID1"})
    c.tick()
    assert "id1" not in c.snapshot()["unassigned"]                                                  # :2421


# scheduler_test.go:2424-2520 TestPreassignedTasks
def scenario_preassigned_tasks(make):
    c = Cluster(make(), nodes=_ready_nodes(["node1", "node2"]),
                tasks=[task("task1"), task("task2", node_id="node1"), task("task3", node_id="node1")], services=[""])
    pre = c.assignments(c.preassigned())
    assert pre == {"task2": "node1", "task3": "node1"}                                              # :2498-2509
    assert c.assignments(c.tick()) == {"task1": "node2"}                                            # :2513-2514


# scheduler_test.go:2522-2630 TestIgnoreTasks
def scenario_ignore_tasks(make):
    c = Cluster(make(), nodes=_ready_nodes(["node1"]),
                tasks=[task("task1"), task("task2", node_id="node1", desired_state="SHUTDOWN"),
                       task("task3", node_id="node1", desired_state="REMOVE")], services=[""])
    assert c.assignments(c.run()) == {"task1": "node1"}                                             # :2617-2618


# scheduler_test.go:2632-2833 TestUnscheduleableTask
def scenario_unscheduleable_task(make):
    pl = placement(max_replicas=1)
    n = node("nodeid1", description=description())
    mk = lambda tid, ver, **kw: task(tid, service_id="serviceid1", spec_version=ver, placement=pl, **kw)   # noqa: E731
    c = Cluster(make(), nodes=[n], tasks=[mk("taskid1", 0), mk("taskid2", 0)], services=[("serviceid1", 0)])
    d = c.run()
    assigned = [k for k, v in d.items() if v["state"] == "ASSIGNED"]
    failed = [k for k, v in d.items() if v["state"] == "PENDING"]
    assert len(assigned) == 1 and len(failed) == 1
    assert d[failed[0]]["err"] == "no suitable node (max replicas per node limit exceed)"           # :2757
    c.set_service("serviceid1", 1)
    c.create_task(mk("taskid1update", 1))
    t = dict(c.tasks[assigned[0]])
    t["status"] = dict(t["status"], state="RUNNING")
    c.update_task(t)
    t = dict(c.tasks[failed[0]])
    t["desired_state"] = "SHUTDOWN"
    c.update_task(t)
    d = c.run()
    assert d[failed[0]]["state"] == "SHUTDOWN"                                                      # :2814-2817


# scheduler_test.go:2868-3336 TestSchedulerPluginConstraint
def scenario_plugin_constraint(make):
    def n(i, plugins):
        return node(f"node{i}_ID", description=description(engine=engine(plugins=plugins)))
    n1 = n(1, [("Volume", "plugin1"), ("Log", "default")])
    n2 = n(2, [("Volume", "plugin1"), ("Volume", "plugin2"), ("Log", "default")])
    n3 = n(3, [("Volume", "plugin1"), ("Network", "plugin1"), ("Log", "default")])
    n4 = n(4, [("Log", "plugin1")])
    vol = lambda d: {"type": "VOLUME", "driver": d}                                                 # noqa: E731
    T = lambda tid, **kw: task(tid, service_id="serviceID1", **kw)                                  # noqa: E731
    c = Cluster(make(), nodes=[n1], tasks=[T("task1_ID", mounts=[vol("plugin1")])], services=["serviceID1"])
    assert c.assignments(c.run())["task1_ID"] == "node1_ID"                                         # :3228
    c.create_task(T("task0_ID", mounts=[{"type": "BIND", "driver": None}]))
    assert c.assignments(c.run())["task0_ID"] == "node1_ID"                                         # :3239-3240
    c.create_task(T("task2_ID", mounts=[vol("plugin1"), vol("plugin2")]))
    assert c.run()["task2_ID"]["err"] == "no suitable node (missing plugin on 1 node)"             # :3248
    c.create_node(n2)
    assert c.assignments(c.run())["task2_ID"] == "node2_ID"                                         # :3259-3260
    c.create_task(T("task3_ID", mounts=[vol("plugin1")], networks=["plugin1"]))
    assert c.run()["task3_ID"]["err"] == "no suitable node (missing plugin on 2 nodes)"            # :3271
    c.create_node(n3)
    assert c.assignments(c.run())["task3_ID"] == "node3_ID"                                         # :3282-3283
    c.create_task(T("task4_ID", log_driver="plugin1"))
    assert c.run()["task4_ID"]["err"] == "no suitable node (missing plugin on 3 nodes)"            # :3295
    c.create_node(n4)
    assert c.assignments(c.run())["task4_ID"] == "node4_ID"                                         # :3306-3307
    c.create_task(T("task5_ID", log_driver="plugin1"))
    assert c.assignments(c.run())["task5_ID"] == "node4_ID"                                         # :3316-3317
    c.create_task(T("task6_ID", log_driver="none"))
    assert c.assignments(c.run())["task6_ID"] != ""                                                 # :3326-3327
    c.create_task(T("task7_ID", log_driver=""))
    assert c.assignments(c.run())["task7_ID"] != ""                                                 # :3335-3336


# scheduler_test.go:3470-3629 TestSchedulerHostPort
def scenario_host_port(make):
    T = lambda tid, ports: task(tid, service_id="serviceID1", ports=ports)                          # noqa: E731
    c = Cluster(make(), tasks=[T("id1", [host_port(58, "TCP")]), T("id2", [host_port(58, "UDP")])], services=["serviceID1"])
    d = c.run()
    assert all(v["state"] == "PENDING" and v["err"] for v in d.values()) and len(d) == 2            # :3598-3599
    c.create_node(node("nodeid1"))
    c.create_node(node("nodeid2"))
    a = c.assignments(c.run())
    assert a["id1"] != a["id2"]                                                                     # :3613-3615
    c.create_task(T("id3", [host_port(58, "UDP"), host_port(58, "TCP")]))
    assert c.run()["id3"]["err"] == "no suitable node (host-mode port already in use on 2 nodes)"  # :3628


# scheduler_test.go:3631-3882 TestSchedulerMaxReplicas
def scenario_max_replicas(make):
    T = lambda tid, pl: task(tid, service_id="serviceID1", placement=pl)                            # noqa: E731
    one = placement(max_replicas=1)
    c = Cluster(make(), tasks=[T("id1", one), T("id2", one)], services=["serviceID1"])
    d = c.run()
    assert all(v["state"] == "PENDING" for v in d.values()) and len(d) == 2                         # :3731-3732
    c.create_node(node("nodeid1"))
    c.create_node(node("nodeid2"))
    a = c.assignments(c.run())
    assert a["id1"] != a["id2"]                                                                     # :3746-3748
    c.create_task(T("id3", one))
    assert c.run()["id3"]["err"] == "no suitable node (max replicas per node limit exceed)"        # :3764
    c.create_node(node("nodeid3"))
    assert c.assignments(c.run())["id3"] == "nodeid3"
    three = placement(constraints=["node.hostname==node1"], max_replicas=3)
    for i in (4, 5, 6):
        c.create_task(T(f"id{i}", three))
    c.run()
    c.create_task(T("id7", three))
    assert c.run()["id7"]["err"] == "no suitable node (scheduling constraints not satisfied on 3 nodes)"   # :3881
