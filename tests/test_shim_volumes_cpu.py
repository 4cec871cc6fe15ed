"""Cluster (CSI) volumes through the PRODUCT's host shim (scheduler_host.cpp::scheduleVolumeGroup, fit_run) on the oracle
ABI, against the object-level oracle: the reference's scheduler-level scenarios (scheduler_ginkgo_test.go) and random event
streams with volumes of every access mode.  CPU only: on the GPU the same shim source drives the CUDA engine through the
primitives the GPU suite already covers (leaf-restricted groups, pe_fit, row upserts)."""
import random

import pytest

import tests.test_oracle_volumes_scheduler as ref_scenarios
from tests.oracle_lib import build_sched, build_shim_on_oracle
from tests.sched_harness import (Cluster, JsonScheduler, cluster_mount, comparable, csi_volume, description, discrete, host_port, named, node, placement,
                                 resources, task)


def make_shim():
    return JsonScheduler(build_shim_on_oracle(), "sso")


def make_oracle():
    return JsonScheduler(build_sched(), "so")


SCENARIOS = [n for n in dir(ref_scenarios) if n.startswith("test_")]


@pytest.mark.parametrize("name", SCENARIOS)
def test_reference_volume_scenarios_through_the_shim(name, monkeypatch):
    monkeypatch.setattr(ref_scenarios, "make_oracle", make_shim)
    getattr(ref_scenarios, name)()


ZONES = ["z1", "z2", "z3"]


def _node(rng, i):
    csi = []
    if rng.random() < 0.8:
        csi.append(("plugA", f"a{i}", {"zone": rng.choice(ZONES)} if rng.random() < 0.8 else None))
    if rng.random() < 0.4:
        csi.append(("plugB", f"b{i}", {"zone": rng.choice(ZONES), "rack": rng.choice(["r1", "r2"])}))
    lr = random.Random(i * 7919 + 13)
    labels = {k: lr.choice(v) for k, v in (("az", ["a", "b", "c"]), ("rack", ["r1", "r2"])) if lr.random() < 0.85}
    gen = ([discrete("apple", lr.randint(1, 4))] if lr.random() < 0.6 else []) + ([named("gpu", f"g{j}") for j in range(lr.randint(1, 3))] if lr.random() < 0.4 else [])
    return node(f"n{i:03d}", state=rng.choice(["READY"] * 9 + ["DOWN"]), labels=labels,
                description=description(resources=resources(rng.randint(2, 8) * 10**9, 2**34, gen), csi_info=csi))


def _volume(rng, i):
    scope = rng.choice(["SINGLE_NODE", "MULTI_NODE"])
    sharing = rng.choice(["NONE", "READ_ONLY", "ONE_WRITER", "ALL"])
    top = None if rng.random() < 0.3 else [{"zone": rng.choice(ZONES)} for _ in range(rng.randint(1, 2))]
    return csi_volume(f"vol{i:02d}", f"data{i}", group=rng.choice(["", "g1", "g2"]), driver=rng.choice(["plugA", "plugA", "plugB"]), scope=scope, sharing=sharing,
                      volume_id=(f"csi{i}" if rng.random() < 0.9 else ""), accessible_topology=top, availability=rng.choice(["ACTIVE"] * 6 + ["PAUSE"]))


def _mounts(rng, n_vol):
    ms = []
    for _ in range(rng.randint(1, 2)):
        src = rng.choice([f"data{rng.randrange(n_vol)}", f"data{rng.randrange(n_vol)}", "group:g1", "group:g2", "nosuch"])
        ms.append(cluster_mount(src, rng.choice(["/a", "/b", "/c"]), read_only=rng.random() < 0.4))
    if rng.random() < 0.3:
        ms.append({"type": "BIND", "driver": None, "source": "/x", "target": "/y", "read_only": False})
    return ms


@pytest.mark.parametrize("seed", range(16))
def test_random_event_stream_with_volumes(seed):
    rng = random.Random(9000 + seed)
    n_nodes, n_vol = rng.randint(4, 30), rng.randint(2, 8)
    nodes = [_node(rng, i) for i in range(n_nodes)]
    vols = [_volume(rng, i) for i in range(n_vol)]
    cm = Cluster(make_shim(), nodes=nodes, volumes=vols, services=["svc"])
    co = Cluster(make_oracle(), nodes=nodes, volumes=vols, services=["svc"])
    next_id = 0
    for step in range(10):
        r2 = random.Random(seed * 1000 + step)
        new = []
        for _ in range(r2.randint(1, 8)):
            kind = r2.random()
            res = resources(r2.choice([0, 5, 10]) * 10**8, 0) if r2.random() < 0.5 else None
            if kind < 0.65:      # one-off task with cluster mounts (a third of them spread over a label as well)
                pl = placement(preferences=["node.labels.az"]) if r2.random() < 0.33 else None
                new.append(task(f"t{next_id:04d}", service_id="svc", reservations=res, mounts=_mounts(r2, n_vol), placement=pl))
            elif kind < 0.8:     # preassigned (global-mode) task with cluster mounts
                new.append(task(f"t{next_id:04d}", service_id="svc", node_id=f"n{r2.randrange(n_nodes):03d}", reservations=res, mounts=_mounts(r2, n_vol)))
            else:                # plain task
                new.append(task(f"t{next_id:04d}", service_id="svc", reservations=res))
            next_id += 1
        for c in (cm, co):
            for t in new:
                c.create_task(t)
        dm, do = cm.run(), co.run()
        assert comparable(dm, cm) == comparable(do, co), f"seed {seed} step {step}: decisions differ"
        assert cm.s.apply({"op": "volume_usage"})["volumes"] == co.s.apply({"op": "volume_usage"})["volumes"], f"seed {seed} step {step}: volume usage differs"
        sm, so = cm.s.apply({"op": "device_check"}), co.snapshot()
        assert sm["mismatch"] == [] and sm["nodes"] == so["nodes"] and sm["unassigned"] == so["unassigned"]
        for c in (cm, co):
            r3 = random.Random(seed * 77 + step)
            running = [t for t in c.tasks.values() if t["status"]["state"] == "ASSIGNED"]
            for t in r3.sample(running, min(len(running), r3.randint(0, 3))):
                c.update_task(dict(t, status=dict(t["status"], state="SHUTDOWN")))
                c.delete_task(t["id"])
            if r3.random() < 0.4:
                c.update_volume(_volume(r3, r3.randrange(n_vol + 2)))
            if r3.random() < 0.3:
                c.update_node(_node(r3, r3.randrange(n_nodes)))
            c.advance(5)


def _static_volume(rng, i):
    """Volumes whose availability no placement of the group can change (VolumeBook::staticFor)."""
    top = None if rng.random() < 0.3 else [{"zone": rng.choice(ZONES)} for _ in range(rng.randint(1, 2))]
    return csi_volume(f"vol{i:02d}", f"data{i}", group=rng.choice(["", "g1", "g2"]), driver=rng.choice(["plugA", "plugA", "plugB"]), scope="MULTI_NODE",
                      sharing=rng.choice(["READ_ONLY", "ONE_WRITER", "ALL"]), volume_id=(f"csi{i}" if rng.random() < 0.9 else ""), accessible_topology=top,
                      availability=rng.choice(["ACTIVE"] * 6 + ["PAUSE"]))


@pytest.mark.parametrize("seed", range(16))
def test_random_replicated_services_on_static_volumes(seed):
    """Groups of k > 1 identical tasks whose cluster mounts are read-only or on share-all volumes: one engine group on the
    node set the host-side VolumesFilter allows; half the services also spread over node labels (every leaf visit of
    their preference tree carries the volume term too)."""
    rng = random.Random(9500 + seed)
    n_nodes, n_vol = rng.randint(4, 30), rng.randint(2, 6)
    nodes = [_node(rng, i) for i in range(n_nodes)]
    vols = [_static_volume(rng, i) for i in range(n_vol)]
    svcs = [(f"svc{j}", 1) for j in range(4)]
    cm = Cluster(make_shim(), nodes=nodes, volumes=vols, services=svcs)
    co = Cluster(make_oracle(), nodes=nodes, volumes=vols, services=svcs)
    next_id = 0
    for step in range(6):
        r2 = random.Random(seed * 1000 + step)
        new = []
        for j in r2.sample(range(4), r2.randint(1, 3)):
            r_svc = random.Random(seed * 31 + j)          # the spec of a service is fixed: its tasks form one group
            mounts = []
            for _ in range(r_svc.randint(1, 2)):
                vol_ro = r_svc.random() < 0.6
                if vol_ro:
                    src = r_svc.choice([f"data{r_svc.randrange(n_vol)}", "group:g1", "group:g2"])
                else:         # a writer: only on volumes that do not count their writers (the other kind is refused, below)
                    ok = [v["name"] for v in vols if v["sharing"] != "ONE_WRITER"]
                    src, vol_ro = (r_svc.choice(ok), False) if ok else ("group:g1", True)
                mounts.append(cluster_mount(src, r_svc.choice(["/a", "/b"]), read_only=vol_ro))
            res = resources(r_svc.choice([0, 5, 10]) * 10**8, 0)
            pl = placement(preferences=r_svc.choice([["node.labels.az"], ["node.labels.az", "node.labels.rack"]])) if r_svc.random() < 0.5 else None
            for _ in range(r2.randint(2, 9)):
                new.append(task(f"t{next_id:04d}", service_id=f"svc{j}", spec_version=1, reservations=res, mounts=mounts, placement=pl))
                next_id += 1
        for c in (cm, co):
            for t in new:
                c.create_task(t)
        dm, do = cm.run(), co.run()
        assert comparable(dm, cm) == comparable(do, co), f"seed {seed} step {step}: decisions differ"
        assert cm.s.apply({"op": "volume_usage"})["volumes"] == co.s.apply({"op": "volume_usage"})["volumes"]
        sm, so = cm.s.apply({"op": "device_check"}), co.snapshot()
        assert sm["mismatch"] == [] and sm["nodes"] == so["nodes"] and sm["unassigned"] == so["unassigned"]
        for c in (cm, co):
            c.advance(5)


@pytest.mark.parametrize("seed", range(16))
def test_random_replicated_services_on_volumes_that_count_their_users(seed):
    """Groups of k > 1 identical tasks on volumes of every scope / sharing, writers included: availability moves with
    every placement, the reference re-runs VolumesFilter inside its fill loop (scheduler.go:912-920) and the shim walks
    that loop one engine question at a time (scheduleVolumeGroupStepwise); half the services spread over node labels as
    well, so the walk happens inside the leaves of their preference tree (fillLeafStepwise)."""
    rng = random.Random(9900 + seed)
    n_nodes, n_vol = rng.randint(3, 24), rng.randint(2, 8)
    nodes = [_node(rng, i) for i in range(n_nodes)]
    vols = [_volume(rng, i) for i in range(n_vol)]
    svcs = [(f"svc{j}", 1) for j in range(4)]
    cm = Cluster(make_shim(), nodes=nodes, volumes=vols, services=svcs)
    co = Cluster(make_oracle(), nodes=nodes, volumes=vols, services=svcs)
    next_id = 0
    for step in range(6):
        r2 = random.Random(seed * 1000 + step)
        new = []
        for j in r2.sample(range(4), r2.randint(1, 3)):
            r_svc = random.Random(seed * 31 + j)          # the spec of a service is fixed: its tasks form one group
            mounts = [cluster_mount(r_svc.choice([f"data{r_svc.randrange(n_vol)}", "group:g1", "group:g2"]), r_svc.choice(["/a", "/b"]),
                                    read_only=r_svc.random() < 0.4) for _ in range(r_svc.randint(1, 2))]
            res = resources(r_svc.choice([0, 5, 10, 20]) * 10**8, 0, [discrete(r_svc.choice(["apple", "gpu"]), 1)] if r_svc.random() < 0.4 else [])
            pl = placement(constraints=r_svc.choice([[], [], ["node.labels.az != c"], ["node.labels.rack == r1"]]), max_replicas=r_svc.choice([0, 0, 1, 2]),
                           preferences=r_svc.choice([[], [], ["node.labels.az"], ["node.labels.az", "node.labels.rack"]]))
            ports = [host_port(r_svc.choice([80, 443]))] if r_svc.random() < 0.3 else None
            for _ in range(r2.randint(2, 7)):
                new.append(task(f"t{next_id:04d}", service_id=f"svc{j}", spec_version=1, reservations=res, mounts=mounts, placement=pl, ports=ports))
                next_id += 1
        for c in (cm, co):
            for t in new:
                c.create_task(t)
        dm, do = cm.run(), co.run()
        assert comparable(dm, cm) == comparable(do, co), f"seed {seed} step {step}: decisions differ"
        assert cm.s.apply({"op": "volume_usage"})["volumes"] == co.s.apply({"op": "volume_usage"})["volumes"]
        sm, so = cm.s.apply({"op": "device_check"}), co.snapshot()
        assert sm["mismatch"] == [] and sm["nodes"] == so["nodes"] and sm["unassigned"] == so["unassigned"]
        for c in (cm, co):
            r3 = random.Random(seed * 77 + step)
            running = [t for t in c.tasks.values() if t["status"]["state"] == "ASSIGNED"]
            for t in r3.sample(running, min(len(running), r3.randint(0, 4))):
                c.update_task(dict(t, status=dict(t["status"], state="SHUTDOWN")))
                c.delete_task(t["id"])
            if r3.random() < 0.3:
                c.update_volume(_volume(r3, r3.randrange(n_vol + 2)))
            c.advance(5)


def test_three_writers_one_single_writer_volume():
    """k = 3 tasks writing to a ONE_WRITER volume: the first takes it, the re-check of every other node fails on
    VolumesFilter, two tasks stay pending with the reference's explanation -- never two writers."""
    nodes = [node(f"n{i}", description=description(resources=resources(8 * 10**9, 2**34), csi_info=[("plugA", f"a{i}", None)])) for i in range(3)]
    vols = [csi_volume("vol0", "data0", driver="plugA", scope="MULTI_NODE", sharing="ONE_WRITER", volume_id="csi0")]
    out = []
    for mk in (make_shim, make_oracle):
        c = Cluster(mk(), nodes=nodes, volumes=vols, services=[("svc", 1)])
        for i in range(3):
            c.create_task(task(f"t{i}", service_id="svc", spec_version=1, mounts=[cluster_mount("data0", "/a")]))
        d = c.run()
        out.append((comparable(d, c), c.s.apply({"op": "volume_usage"})["volumes"]))
    assert out[0] == out[1]
    dec, usage = out[0]
    assert sorted(k for k, v in dec.items() if v["state"] == "ASSIGNED") == ["t0"] and list(usage["vol0"]) == ["t0"]
    assert dec["t1"]["err"] == "no suitable node (cannot fulfill requested CSI volume mounts on 3 nodes)"   # n1, n2 and, after the wrap, n0 itself


def test_volume_node_set_is_marked_incrementally():
    """Fifty one-off tasks with the same read-only mount on 200 nodes: the node set VolumesFilter allows is written into
    the node table once (its 100 rows); every later task finds it marked and only the rows the tick itself changed move."""
    nodes = [node(f"n{i:03d}", description=description(resources=resources(8 * 10**9, 2**34),
                                                        csi_info=[("plugA", f"a{i}", {"zone": "z1" if i % 2 else "z2"})])) for i in range(200)]
    vols = [csi_volume("vol0", "data0", driver="plugA", scope="MULTI_NODE", sharing="ALL", volume_id="csi0", accessible_topology=[{"zone": "z1"}])]
    c = Cluster(make_shim(), nodes=nodes, volumes=vols, services=["svc"])
    c.run()
    s0 = c.s.apply({"op": "device_check"})
    for i in range(50):
        c.create_task(task(f"t{i:02d}", service_id="svc", mounts=[cluster_mount("data0", "/a", read_only=True)]))
    d = c.run()
    assert all(v["state"] == "ASSIGNED" and int(v["node_id"][1:]) % 2 == 1 for v in d.values())
    s1 = c.s.apply({"op": "device_check"})
    assert s1["mismatch"] == []
    # one full upload when the mark column is allocated (200 rows) and nothing else: placements move the device rows themselves
    assert s1["full_uploads"] - s0["full_uploads"] == 1 and s1["rows_uploaded"] - s0["rows_uploaded"] == 200, (s0, s1)
    for i in range(50, 100):
        c.create_task(task(f"t{i:02d}", service_id="svc", mounts=[cluster_mount("data0", "/a", read_only=True)]))
    c.run()
    s2 = c.s.apply({"op": "device_check"})
    assert s2["mismatch"] == [] and s2["full_uploads"] == s1["full_uploads"] and s2["rows_uploaded"] == s1["rows_uploaded"], (s1, s2)
