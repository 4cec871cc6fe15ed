"""The oracle's restatement of the CSI cluster-volume bookkeeping and of VolumesFilter (the product
refuses such groups) against the reference's own known answers:
topology_test.go:9-177, volumes_test.go:47-160, :164-343, :344-471, :473-528, filter.go:399-440."""
import pytest

from tests.oracle_lib import VolumesOracle


def volume(vid, name, group="", driver="driver", scope="SINGLE_NODE", sharing="NONE", accessible=None, availability="ACTIVE", info=True):
    return {"id": vid, "name": name, "group": group, "driver": driver, "scope": scope, "sharing": sharing, "availability": availability,
            "accessible_topology": accessible, "volume_info": {} if info else None}


def canned(i):   # volumes_test.go:16-36
    return volume(f"volumeID{i}", f"volume{i}", group="group", driver="driver", scope="MULTI_NODE", sharing="ALL")


def seg(**kw):
    return dict(kw)


TOPOLOGY = [   # topology_test.go:11-173
    (seg(region="R1", zone="Z1"), [seg(region="R1", zone="Z1")], True),
    (seg(region="R1", zone="Z2"), [seg(region="R1", zone="Z1"), seg(region="R1", zone="Z2")], True),
    (seg(region="R1", zone="Z3"), [seg(region="R1")], True),
    (seg(region="R1", zone="Z1"), [seg(region="R2", zone="Z1")], False),
    (seg(region="R1", zone="Z1", shelf="S1"), [seg(region="R1", zone="Z1"), seg(region="R1", zone="Z2")], True),
    (seg(region="R1", zone="Z1", shelf="S1"), [seg(region="R1", zone="Z1", shelf="S2"), seg(region="R1", zone="Z2", shelf="S1")], False),
    (seg(region="R1", zone="Z1", shelf="S1"), [seg(region="R1", zone="Z1", shelf="S2"), seg(region="R1", zone="Z2", shelf="S1"),
                                               seg(region="R1", zone="Z1", shelf="S1")], True),
]


@pytest.mark.parametrize("top,accessible,expected", TOPOLOGY)
def test_is_in_topology(top, accessible, expected):
    assert VolumesOracle()(op="is_in_topology", top=top, accessible=accessible)["result"] is expected


def test_is_in_topology_missing_parts_fit():   # topology.go:23-26
    vo = VolumesOracle()
    assert vo(op="is_in_topology", top=None, accessible=[seg(zone="z")])["result"] is True
    assert vo(op="is_in_topology", top=seg(zone="z"), accessible=None)["result"] is True
    assert vo(op="is_in_topology", top=seg(zone="z"), accessible=[])["result"] is True


def test_add_remove_track_reserve():   # volumes_test.go:47-160
    vo = VolumesOracle()
    v1, v2 = canned(1), canned(2)
    vo(op="add_volume", volume=v1); vo(op="add_volume", volume=v2)
    d = vo(op="dump")
    assert set(d["volumes"]) == {"volumeID1", "volumeID2"} and all(v == {"tasks": {}, "nodes": {}} for v in d["volumes"].values())
    assert d["by_group"] == {"group": ["volumeID1", "volumeID2"]}
    assert d["by_name"] == {"volume1": "volumeID1", "volume2": "volumeID2"}
    vo(op="remove_volume", id="volumeID1")                      # :85-102
    d = vo(op="dump")
    assert set(d["volumes"]) == {"volumeID2"} and d["by_name"] == {"volume2": "volumeID2"} and d["by_group"] == {"group": ["volumeID2"]}
    vo(op="add_volume", volume=v1)
    vo(op="reserve", volume="volumeID1", task="task1", node="node1", read_only=True)   # :104-110
    assert vo(op="dump")["volumes"]["volumeID1"]["tasks"] == {"task1": {"node": "node1", "read_only": True}}
    vo(op="release", volume="volumeID1", task="task1")
    # reserveTaskVolumes, :112-160: attachments matched to mounts by (source, target); bind mounts are ignored
    mounts = [{"type": "CLUSTER", "source": "group", "target": "/var/spool/mail"},
              {"type": "BIND", "source": "/var/run/docker.sock", "target": "/var/run/docker.sock"},
              {"type": "CLUSTER", "source": "volume2", "target": "/srv/www", "read_only": True}]
    atts = [{"id": "volumeID1", "source": "group", "target": "/var/spool/mail"}, {"id": "volumeID2", "source": "volume2", "target": "/srv/www"}]
    vo(op="reserve_task", task="task1", node="node1", attachments=atts, mounts=mounts)
    d = vo(op="dump")["volumes"]
    assert d["volumeID1"]["tasks"] == {"task1": {"node": "node1", "read_only": False}}
    assert d["volumeID2"]["tasks"] == {"task1": {"node": "node1", "read_only": True}}
    assert d["volumeID1"]["nodes"] == {"node1": 1}
    vo(op="release", volume="volumeID1", task="task1")          # node reference counts, :632-650
    assert vo(op="dump")["volumes"]["volumeID1"] == {"tasks": {}, "nodes": {"node1": 0}}


MN_RO, MN_OW, MN_ALL, SN_NONE = ("MULTI_NODE", "READ_ONLY"), ("MULTI_NODE", "ONE_WRITER"), ("MULTI_NODE", "ALL"), ("SINGLE_NODE", "NONE")
CHECK_TABLE = [   # volumes_test.go:277-343: (access mode or None = single node / all, in use, in topology, read only, expected)
    ("volume outside of node topology", None, "unused", False, False, False),
    ("volume in use on a different node", None, "wrong_node", True, False, False),
    ("volume is read only, mount is not", MN_RO, "unused", True, False, False),
    ("volume is OneWriter, but already has a writer", MN_OW, "writer", True, False, False),
    ("volume is OneWriter, and has no writer", MN_OW, "only_readers", True, False, True),
    ("volume not in use and is in topology", None, "unused", True, False, True),
    ("in use on a different node, but the scope is multinode", MN_ALL, "wrong_node", True, False, True),
    ("the volume is in use and cannot be shared", SN_NONE, "only_readers", True, True, False),
    ("the volume is not in use and cannot be shared", SN_NONE, "unused", True, True, True),
]


@pytest.mark.parametrize("name,mode,in_use,in_topology,read_only,expected", CHECK_TABLE, ids=[c[0] for c in CHECK_TABLE])
def test_check_volume_table(name, mode, in_use, in_topology, read_only, expected):
    vo = VolumesOracle()
    scope, sharing = mode or ("SINGLE_NODE", "ALL")
    vo(op="add_volume", volume=volume("someVolume", "", driver="somePlugin", scope=scope, sharing=sharing, accessible=[seg(zone="z1")]))
    node = {"id": "someNode", "csi": [{"plugin": "somePlugin", "topology": seg(zone="z1" if in_topology else "z2")}]}
    if in_use == "wrong_node":
        vo(op="reserve", volume="someVolume", task="someTask", node="someOtherNode", read_only=False)
    elif in_use == "only_readers":
        vo(op="reserve", volume="someVolume", task="someTask", node="someNode", read_only=True)
    elif in_use == "writer":
        vo(op="reserve", volume="someVolume", task="someTask", node="someNode", read_only=True)
        vo(op="reserve", volume="someVolume", task="someWriter", node="someNode", read_only=False)
    assert vo(op="check", volume="someVolume", node=node, read_only=read_only)["result"] is expected


def _group_setup():   # volumes_test.go:350-441
    vo = VolumesOracle()
    node = {"id": "someNode", "csi": [{"plugin": "newPlugin", "topology": None}]}
    vo(op="add_volume", volume=volume("volume1", "volumeName1", driver="newPlugin", sharing="ALL"))
    vo(op="add_volume", volume=volume("volume2", "volumeName2", driver="newPlugin", sharing="ALL", info=False))
    vo(op="add_volume", volume=volume("volume3", "volumeName3", group="someVolumeGroup", driver="newPlugin", sharing="ALL"))
    vo(op="add_volume", volume=volume("volume4", "volumeName4", group="someVolumeGroup", driver="newPlugin", sharing="ALL"))
    return vo, node


def test_volume_or_group_availability():   # volumes_test.go:445-470
    vo, node = _group_setup()
    assert vo(op="available", mounts=[{"type": "CLUSTER", "source": "volumeName1", "target": ""}], node=node)["result"] == "volume1"
    assert vo(op="available", mounts=[{"type": "CLUSTER", "source": "volumeNameNotReal", "target": ""}], node=node)["result"] == ""
    got = vo(op="available", mounts=[{"type": "CLUSTER", "source": "group:someVolumeGroup", "target": ""}], node=node)["result"]
    assert got in ("volume3", "volume4") and got == "volume3"        # canonical: ascending volume ID inside a group
    assert vo(op="available", mounts=[{"type": "CLUSTER", "source": "group:noSuchGroup", "target": ""}], node=node)["result"] == ""


def test_choose_task_volumes():   # volumes_test.go:473-528
    vo = VolumesOracle()
    v1 = canned(1); v1["group"] = "volumeGroup"
    for v in (v1, canned(2), canned(3)):
        vo(op="add_volume", volume=v)
    mounts = [{"type": "CLUSTER", "source": "group:volumeGroup", "target": "/somedir", "read_only": True},
              {"type": "CLUSTER", "source": "volume2", "target": "/someOtherDir"},
              {"type": "BIND", "source": "/some/subdir", "target": "/some/container/dir"},
              {"type": "CLUSTER", "source": "volume3", "target": "/some/third/dir"}]
    r = vo(op="choose", task="taskID1", mounts=mounts, node={"id": "node1", "csi": []})
    assert r["ok"] and r["attachments"] == [{"id": "volumeID1", "source": "group:volumeGroup", "target": "/somedir"},
                                            {"id": "volumeID2", "source": "volume2", "target": "/someOtherDir"},
                                            {"id": "volumeID3", "source": "volume3", "target": "/some/third/dir"}]
    # choosing leaves no reservation behind (volumes.go:101-108: the deferred release)
    assert all(v["tasks"] == {} for v in vo(op="dump")["volumes"].values())
    bad = vo(op="choose", task="taskID2", mounts=[{"type": "CLUSTER", "source": "nope", "target": "/x"}], node={"id": "node1", "csi": []})
    assert not bad["ok"] and bad["error"] == "cannot find volume to satisfy mount with source nope"


def test_volumes_filter():   # filter.go:399-440
    vo, node = _group_setup()
    other = {"id": "otherNode", "csi": [{"plugin": "newPlugin", "topology": None}]}
    # SetTask: enabled only by CLUSTER mounts of a container spec
    assert vo(op="filter", mounts=[{"type": "BIND", "source": "/a", "target": "/b"}], nodes=[node])["enabled"] is False
    assert vo(op="filter", has_container=False, mounts=[{"type": "CLUSTER", "source": "volumeName1", "target": "/b"}], nodes=[node])["enabled"] is False
    csi = [{"type": "CLUSTER", "source": "volumeName1", "target": "/b"}]
    r = vo(op="filter", mounts=csi, nodes=[node, other])
    assert r["enabled"] is True and r["pass"] == [True, True]
    # a single-node volume in use elsewhere fails on every other node, passes on its own
    vo(op="reserve", volume="volume1", task="t0", node="someNode", read_only=False)
    assert vo(op="filter", mounts=csi, nodes=[node, other])["pass"] == [True, False]
    # Check is an OR over the requested mounts (filter.go:433-438): one satisfiable mount is enough
    two = csi + [{"type": "CLUSTER", "source": "volumeName2", "target": "/c"}]
    assert vo(op="filter", mounts=two, nodes=[other])["pass"] == [True]
    assert vo(op="filter", mounts=[{"type": "CLUSTER", "source": "missing", "target": "/c"}], nodes=[node])["pass"] == [False]
