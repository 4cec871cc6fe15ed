"""The reference's scheduler-level CSI volume scenarios (scheduler_ginkgo_test.go:80-373, :375-583) against the object-level
oracle -- SURVEY 8(a) a13.  The product refuses groups with cluster mounts; these pin what a host-side
VolumesFilter + volume bookkeeping has to do when it is built."""
from tests.oracle_lib import build_sched
from tests.sched_harness import Cluster, JsonScheduler, cluster_mount, csi_volume, description, node, task


def make_oracle():
    return JsonScheduler(build_sched(), "so")


def _node1():
    return node("nodeID1", description=description(csi_info=[("somePlug", "nodeCSI1")]))


def _volume1(volume_id="csi1"):
    return csi_volume("volumeID1", "volume1", driver="somePlug", scope="SINGLE_NODE", sharing="NONE", volume_id=volume_id)


def _task1(node_id=""):
    return task("task1", service_id="service1", node_id=node_id, mounts=[cluster_mount("volume1", "/var/")])


ATTACHMENT = [{"id": "volumeID1", "source": "volume1", "target": "/var/"}]


def test_should_choose_volumes_for_tasks():   # scheduler_ginkgo_test.go:273-353
    c = Cluster(make_oracle(), nodes=[_node1()], tasks=[_task1()], services=["service1"])
    assert c.run()["task1"]["state"] == "PENDING"                       # the volume does not exist yet
    c.update_volume(_volume1(volume_id=""))                              # created in the store, not yet with its plugin: ignored (:211)
    assert c.run()["task1"]["state"] == "PENDING"
    c.update_volume(_volume1())
    d = c.run()["task1"]
    assert d["state"] == "ASSIGNED" and d["node_id"] == "nodeID1" and d["volumes"] == ATTACHMENT
    assert c.s.apply({"op": "volume_usage"})["volumes"]["volumeID1"] == {"task1": {"node": "nodeID1", "read_only": False}}


def test_should_not_commit_a_task_without_a_volume():   # :355-372
    c = Cluster(make_oracle(), nodes=[_node1()], tasks=[_task1()], services=["service1"])
    for _ in range(3):
        d = c.run()["task1"]
        assert d["state"] == "PENDING" and d["node_id"] == ""
        assert d["err"] == "no suitable node (cannot fulfill requested CSI volume mounts on 1 node)"


def test_global_mode_task_still_gets_its_volume():   # :192-251 (preassigned: taskFitNode chooses the attachment)
    c = Cluster(make_oracle(), nodes=[_node1()], tasks=[_task1(node_id="nodeID1")], services=["service1"])
    c.update_volume(_volume1())
    d = c.run()["task1"]
    assert d["state"] == "ASSIGNED" and d["node_id"] == "nodeID1" and d["volumes"] == ATTACHMENT
    assert d["message"] == "scheduler confirmed task can run on preassigned node"


def test_global_mode_task_does_not_progress_without_a_volume():   # :253-270
    c = Cluster(make_oracle(), nodes=[_node1()], tasks=[_task1(node_id="nodeID1")], services=["service1"])
    d = c.run()["task1"]
    assert d["state"] == "PENDING" and d["err"] == "cannot fulfill requested CSI volume mounts on 1 node"


def test_initialization_tracks_in_use_volumes():   # :375-583
    nodes = [node(f"nodeID{i}", description=description(hostname=f"nodeHost{i}", csi_info=[("somePlug", f"nodeCSI{i}")])) for i in range(3)]
    vols = [csi_volume("volumeID1", "volume1", group="group1", driver="somePlug", scope="MULTI_NODE", sharing="ALL", volume_id="csi1"),
            csi_volume("volumeID2", "volume2", group="group2", driver="somePlug", volume_id="csi2"),
            csi_volume("volumeID3", "volume3", group="group2", driver="somePlug", volume_id="csi3")]
    running = task("runningTask", node_id="nodeID0", state="RUNNING", mounts=[cluster_mount("volume1", "/var/"), cluster_mount("group:group2", "/home/")])
    running["volumes"] = [{"source": "volume1", "target": "/var/", "id": "volumeID1"}, {"source": "group:group2", "target": "/home/", "id": "volumeID3"}]
    shutdown = task("shutdownTask", node_id="nodeID1", state="SHUTDOWN", desired_state="SHUTDOWN", mounts=[cluster_mount("volume1", "/foo/")])
    shutdown["volumes"] = [{"source": "volume1", "target": "/foo/", "id": "volumeID1"}]
    pending = task("pendingID", node_id="nodeID2", state="PENDING", mounts=[cluster_mount("group:group2", "/foo/")])
    c = Cluster(make_oracle(), nodes=nodes, tasks=[running, shutdown, pending], volumes=vols)
    use = c.s.apply({"op": "volume_usage"})["volumes"]
    assert set(use) == {"volumeID1", "volumeID2", "volumeID3"}
    assert use["volumeID1"] == {"runningTask": {"node": "nodeID0", "read_only": False}}
    assert use["volumeID2"] == {}
    assert use["volumeID3"] == {"runningTask": {"node": "nodeID0", "read_only": False}}


def test_single_writer_volume_spreads_no_further_than_it_can():
    """Beyond the reference's scenarios: a one-off task per tick on a SINGLE_NODE / sharing ALL volume -- once the volume is in
    use on one node, the VolumesFilter holds every later task to that node (volumes.go:283-291)."""
    nodes = [node(f"n{i}", description=description(csi_info=[("plug", f"csi{i}")])) for i in range(3)]
    vol = csi_volume("v1", "data", driver="plug", scope="SINGLE_NODE", sharing="ALL", volume_id="x")
    c = Cluster(make_oracle(), nodes=nodes, volumes=[vol], services=["svc"])
    placed = []
    for i in range(4):
        c.create_task(task(f"t{i}", service_id="svc", mounts=[cluster_mount("data", "/d")]))
        d = c.run()[f"t{i}"]
        assert d["state"] == "ASSIGNED"
        placed.append(d["node_id"])
    assert placed == ["n0"] * 4
