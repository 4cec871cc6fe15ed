"""Pins the oracle's building blocks against the reference's unit tests:
manager/constraint/constraint_test.go, manager/scheduler/{constraint,nodeinfo,nodeset}_test.go,
api/genericresource/{helpers,resource_management}_test.go."""
import pytest

from tests.oracle_lib import build_sched
from tests.sched_harness import (JsonScheduler, description, discrete, engine, named, node, placement, resources, task)


@pytest.fixture()
def so():
    return JsonScheduler(build_sched(), "so")


def nset(kind, *vals):
    return [named(kind, v) for v in vals]


# ---- manager/constraint/constraint_test.go:9-72 TestParse
@pytest.mark.parametrize("expr,ok,key,exp", [
    ("", False, None, None), (" ", False, None, None), ("nodeabc", False, None, None), ("node ~ abc", False, None, None),
    ("1node==a2", False, None, None), (" node == node1", True, "node", "node1"), ("no de== node1", False, None, None),
    ("no*de==node1", False, None, None), ("==node1", False, None, None), ("node==", False, None, None),
    ("node== ", False, None, None), ("no$de==node1", False, None, None), ("NoDe==node1", True, "NoDe", "node1"),
    ("no.de==node1", True, "no.de", "node1"), ("_node==_node1", True, "_node", "_node1"),
    ("node==[a-b]+c*(n|b)/", True, "node", "[a-b]+c*(n|b)/"), ("node==node 1", True, "node", "node 1"),
])
def test_parse(so, expr, ok, key, exp):
    r = so.apply({"op": "parse_constraints", "constraints": [expr]})
    assert r["ok"] is ok
    if ok:
        assert r["constraints"][0]["key"] == key and r["constraints"][0]["exp"] == exp


# ---- manager/constraint/constraint_test.go:74-117 TestMatch
@pytest.mark.parametrize("expr,whats,expect", [
    ("node.name==foo", ["foo"], True), ("node.name==foo", ["fo"], False), ("node.name==foo", ["fooE"], False),
    ("node.name!=foo", ["foo"], False), ("node.name!=foo", ["bar"], True), ("node.name!=foo", ["fo"], True),
    ("node.name!=foo", ["fooExtra"], True), ("node.name==f*o", ["fo"], False), ("node.name==f*o", ["f*o"], True),
    ("node.name==f*o", ["F*o"], True), ("node.name==f*o", ["foo", "fo", "bar"], False),
    ("node.name==f*o", ["foo", "f*o", "bar"], True), ("node.name==f*o", ["foo"], False),
    ("node.name==f.-$o", ["fa-$o"], False), ("node.name==f.-$o", ["f.-$o"], True),
])
def test_match(so, expr, whats, expect):
    c = so.apply({"op": "parse_constraints", "constraints": [expr]})["constraints"][0]
    r = so.apply({"op": "match", "key": c["key"], "operator": c["op"], "exp": c["exp"], "whats": whats})
    assert r["match"] is expect


def test_equal_fold_kelvin_and_long_s(so):
    # strings.EqualFold is Unicode simple folding: U+212A ~ k, U+017F ~ s (SURVEY hard part C)
    c = so.apply({"op": "parse_constraints", "constraints": ["node.labels.x==kiss"]})["constraints"][0]
    for what, expect in (("Kiss", True), ("KIſs", True), ("kis", False), ("é", False)):
        assert so.apply({"op": "match", "key": c["key"], "operator": 0, "exp": c["exp"], "whats": [what]})["match"] is expect


# ---- manager/scheduler/constraint_test.go (ConstraintFilter against one NodeInfo)
def env_node(**kw):
    base = dict(labels={}, addr="186.17.9.41", description=description(engine=engine(labels={})))
    base.update(kw)
    return node("nodeid-1", **base)


def check(so, constraints, n):
    t = task("id1", state="ASSIGNED", placement=placement(constraints=constraints) if constraints is not None else None)
    r = so.apply({"op": "pipeline_check", "task": t, "node": n})
    return r["enabled"][3], r["pass"]     # (ConstraintFilter.SetTask, Check)


def test_constraint_set_task(so):                                       # constraint_test.go:62-76
    assert check(so, None, env_node())[0] is False
    assert check(so, ["node.hostname == node-2", "node.labels.security != low"], env_node())[0] is True
    assert check(so, ["node.id == nodeid-2", "engine.labels.operatingsystem != ubuntu"], env_node())[0] is True


def test_wrong_syntax(so):                                              # :78-93 unknown key rejects even with !=
    assert check(so, ["node.abc.bcd == high"], env_node()) == (True, False)
    assert check(so, ["node.abc.bcd != high"], env_node()) == (True, False)


def test_node_hostname(so):                                             # :95-118
    c = ["node.hostname != node-1"]
    assert check(so, c, env_node())[1] is True
    for host, expect in (("node-2", True), ("node-1", False), ("NODe-1", False)):
        assert check(so, c, env_node(description=description(hostname=host, engine=engine(labels={}))))[1] is expect


@pytest.mark.parametrize("addr,expr,enabled,verdict", [
    ("186.17.9.41", "node.ip == 186.17.9.41", True, True), ("186.17.9.41", "node.ip != 186.17.9.41", True, False),
    ("186.17.9.41", "node.ip == 186.17.9.42", True, False), ("186.17.9.41", "node.ip == 186.17.9.4/24", True, True),
    ("186.17.9.41", "node.ip == 186.17.8.41/24", True, False), ("186.17.9.41", "node.ip == 186.17.9.41/34", True, False),
    ("186.17.9.41", "node.ip != 266.17.9.41", True, False), ("186.17.9.41", "node.ip != 0.0.0.0", True, True),
    ("186.17.9.41", "node.ip == ", False, None), ("186.17.9.41", "node.ip == not_ip_addr", True, False),
    ("2001:db8::2", "node.ip == 2001:db8::2", True, True), ("2001:db8::2", "node.ip == 2001:db8:0::2", True, True),
    ("2001:db8::2", "node.ip != 2001:db8::2/128", True, False), ("2001:db8::2", "node.ip == 2001:db8::/64", True, True),
    ("2001:db8::2", "node.ip == 2001:db9::/64", True, False), ("2001:db8::2", "node.ip != 2001:db9::/64", True, True),
    ("", "node.ip == 0.0.0.0", True, False), ("", "node.ip != 0.0.0.0", True, True),
])
def test_node_ip(so, addr, expr, enabled, verdict):                     # :120-190
    en, ok = check(so, [expr], env_node(addr=addr))
    assert en is enabled
    if enabled:
        assert ok is verdict


def test_node_id(so):                                                   # :192-216
    assert check(so, ["node.id == nodeid-1"], env_node())[1] is True
    assert check(so, ["node.id == nodeid-1-extra"], env_node())[1] is False
    assert check(so, ["node.id == nodeid-"], env_node())[1] is False


def test_node_role(so):                                                 # :218-239
    assert check(so, ["node.role == worker"], env_node())[1] is True
    assert check(so, ["node.role == manager"], env_node())[1] is False
    assert check(so, ["node.role == worker-manager"], env_node())[1] is False


def test_node_platform(so):                                             # :241-274  (arch is NOT normalised here)
    assert check(so, ["node.platform.os == linux"], env_node())[1] is False
    lin = env_node(description=description(platform={"os": "linux", "arch": "x86_64"}, engine=engine(labels={})))
    win = env_node(description=description(platform={"os": "windows", "arch": "x86_64"}, engine=engine(labels={})))
    assert check(so, ["node.platform.os == linux"], lin)[1] is True
    assert check(so, ["node.platform.os == linux"], win)[1] is False
    assert check(so, ["node.platform.arch == amd64"], win)[1] is False
    assert check(so, ["node.platform.arch != amd64"], win)[1] is True


def test_node_and_engine_labels(so):                                    # :276-312
    c = ["node.labels.security == high"]
    assert check(so, c, env_node())[1] is False
    assert check(so, c, env_node(description=description(engine=engine(labels={"security": "high"}))))[1] is False
    assert check(so, c, env_node(labels={"security": "high"}))[1] is True
    c = ["engine.labels.disk != ssd"]
    assert check(so, c, env_node())[1] is True
    assert check(so, c, env_node(labels={"disk": "ssd"}))[1] is True
    assert check(so, c, env_node(description=description(engine=engine(labels={"disk": "ssd"}))))[1] is False
    assert check(so, c, env_node(description=description(engine=engine(labels={"disk": "ssd", "memory": "large"}))))[1] is False


def test_multiple_constraints(so):                                      # :314-350
    c = ["node.hostname == node-1", "engine.labels.operatingsystem != Ubuntu 14.04"]
    mk = lambda host, oslabel=None, labels=None: env_node(                                         # noqa: E731
        labels=labels or {}, description=description(hostname=host, engine=engine(labels={"operatingsystem": oslabel} if oslabel else {})))
    assert check(so, c, mk(""))[1] is False
    assert check(so, c, mk("node-1"))[1] is True
    assert check(so, c, mk("node-1", "Ubuntu 14.04"))[1] is False
    assert check(so, c, mk("node-1", "ubuntu 14.04"))[1] is False
    assert check(so, c, mk("node-1", "ubuntu 15.04"))[1] is True
    c3 = c + ["node.labels.security == high"]
    assert check(so, c3, mk("node-1", "ubuntu 15.04"))[1] is False
    assert check(so, c3, mk("node-1", "ubuntu 15.04", {"security": "low"}))[1] is False
    assert check(so, c3, mk("node-1", "ubuntu 15.04", {"security": "high"}))[1] is True


# ---- manager/scheduler/nodeinfo_test.go
def test_remove_task(so):                                               # :11-100
    spec = resources(100000, 1000000, nset("orange", "blue", "red", "green") + [discrete("apple", 6)])
    n = node("n", description=description(resources=spec))
    avail = resources(100000, 1000000, nset("orange", "blue", "red") + [discrete("apple", 5)])
    res = resources(5000, 5000, [discrete("apple", 1), discrete("orange", 1)])
    t1 = task("task1", reservations=res)
    t1["assigned_generic"] = nset("orange", "green") + [discrete("apple", 1)]
    r = so.apply({"op": "nodeinfo", "node": n, "tasks": [], "available": avail, "ops": [{"op": "remove", "task": t1}]})
    assert r["results"] == [False]
    r = so.apply({"op": "nodeinfo", "node": n, "tasks": [task("task1"), task("task2")], "available": avail,
                  "ops": [{"op": "remove", "task": t1}, {"op": "remove", "task": task("task3")}]})
    assert r["results"] == [True, False]
    a = r["available"]
    assert a["nano_cpus"] == 105000 and a["memory_bytes"] == 1005000 and len(a["generic"]) == 4
    assert sorted(g["named"] for g in a["generic"] if g["kind"] == "orange") == ["blue", "green", "red"]
    assert [g["value"] for g in a["generic"] if g["kind"] == "apple"] == [6]


def test_add_task(so):                                                  # :102-172
    avail = resources(100000, 1000000, nset("orange", "blue", "red") + [discrete("apple", 5)])
    res = resources(5000, 5000, [discrete("apple", 2), discrete("orange", 1)])
    t3 = task("task3", reservations=res)
    r = so.apply({"op": "nodeinfo", "node": node("n"), "tasks": [task("task1"), task("task2")], "available": avail,
                  "ops": [{"op": "add", "task": task("task1")}, {"op": "add", "task": t3}, {"op": "add", "task": t3}]})
    assert r["results"] == [False, True, False]
    a = r["available"]
    assert a["nano_cpus"] == 95000 and a["memory_bytes"] == 995000
    oranges = [g for g in a["generic"] if g["kind"] == "orange"]
    assert len(oranges) == 1 and oranges[0]["named"] in ("blue", "red")
    assert [g["value"] for g in a["generic"] if g["kind"] == "apple"] == [3]


# ---- manager/scheduler/nodeset_test.go:9-163 TestTreeTaskCountConsistency
def test_tree_task_count_consistency(so):
    def n(i, labels, cnt):
        d = node(f"node{i}", labels=labels)
        d["by_service"] = {"service1": cnt}
        return d
    nodes = [n(1, {"datacenter": "dc1", "rack": "r1"}, 3), n(2, {"datacenter": "dc1", "rack": "r2"}, 2),
             n(3, {"datacenter": "dc2", "rack": "r2"}, 4), n(4, {}, 2), n(5, {}, 1)]
    tr = so.apply({"op": "tree", "nodes": nodes, "service_id": "service1", "max_assignments": 10,
                   "preferences": ["node.labels.datacenter", "node.labels.rack"]})["tree"]

    def verify(d):
        if not d["next"]:
            return d["tasks"]
        assert d["tasks"] == sum(verify(c) for c in d["next"].values())
        return d["tasks"]
    verify(tr)
    assert tr["tasks"] == 12
    assert tr["next"]["dc1"]["tasks"] == 5 and tr["next"]["dc1"]["next"]["r1"]["tasks"] == 3
    assert tr["next"]["dc1"]["next"]["r2"]["tasks"] == 2
    assert tr["next"]["dc2"]["tasks"] == 4 and tr["next"]["dc2"]["next"]["r2"]["tasks"] == 4
    assert tr["next"][""]["tasks"] == 3 and tr["next"][""]["next"][""]["tasks"] == 3


# ---- api/genericresource/helpers_test.go
def gen(so, fn, available, **kw):
    ev = {"op": "generic", "fn": fn, "available": resources(0, 0, available)}
    for k, v in kw.items():
        ev[k] = resources(0, 0, v)
    return so.apply(ev)


def test_consume_resources_single(so):                                  # helpers_test.go:10-27
    a = gen(so, "consume", nset("apple", "red", "orange", "blue"), res=nset("apple", "red"))["available"]
    assert len(a) == 2
    a = gen(so, "consume", a + [discrete("apple", 1)], res=[discrete("apple", 1)])["available"]
    assert len(a) == 2
    a = gen(so, "consume", a + [discrete("apple", 4)], res=[discrete("apple", 1)])["available"]
    assert len(a) == 3 and a[2]["value"] == 3


def test_consume_resources_multiple(so):                                # helpers_test.go:29-62
    avail = (nset("apple", "red", "orange", "blue", "green", "yellow") + [discrete("orange", 5), discrete("banana", 3)]
             + nset("grape", "red", "orange", "blue", "green", "yellow") + [discrete("cakes", 3)])
    res = (nset("apple", "red") + [discrete("banana", 2)] + nset("apple", "green", "blue", "red")
           + nset("grape", "red", "blue", "red") + [discrete("cakes", 3)])
    a = gen(so, "consume", avail, res=res)["available"]
    assert len(a) == 7
    by = lambda k: [g for g in a if g["kind"] == k]                                                # noqa: E731
    assert sorted(g["named"] for g in by("apple")) == ["orange", "yellow"]
    assert sorted(g["named"] for g in by("grape")) == ["green", "orange", "yellow"]
    assert by("orange")[0]["value"] == 5 and by("banana")[0]["value"] == 1


# ---- api/genericresource/resource_management_test.go
def test_claim_single_discrete(so):                                     # :10-23
    r = gen(so, "claim", [discrete("apple", 3)], reservations=[discrete("apple", 2)])
    assert r["ok"] and r["available"] == [discrete("apple", 1)] and r["assigned"] == [discrete("apple", 2)]


def test_claim_multiple_discrete(so):                                   # :25-50
    r = gen(so, "claim", [discrete("apple", 3), discrete("orange", 4), discrete("banana", 2), discrete("cake", 1)],
            reservations=[discrete("orange", 4), discrete("apple", 2)])
    assert r["ok"] and len(r["available"]) == 3 and len(r["assigned"]) == 2
    assert {g["kind"]: g["value"] for g in r["assigned"]} == {"apple": 2, "orange": 4}


def test_claim_single_str(so):                                          # :52-67
    r = gen(so, "claim", nset("apple", "red", "orange", "blue", "green"), reservations=[discrete("apple", 2)])
    assert r["ok"] and len(r["available"]) == 2 and sorted(g["named"] for g in r["assigned"]) == ["orange", "red"]


def test_claim_multiple_str(so):                                        # :69-95
    avail = nset("apple", "red", "orange", "blue", "green") + nset("oranges", "red", "orange", "blue", "green") + \
        nset("bananas", "red", "orange", "blue", "green")
    r = gen(so, "claim", avail, reservations=[discrete("oranges", 4), discrete("apple", 2)])
    assert r["ok"] and len(r["available"]) == 6 and len(r["assigned"]) == 6
    assert sorted(g["named"] for g in r["assigned"] if g["kind"] == "apple") == ["orange", "red"]
    assert sorted(g["named"] for g in r["assigned"] if g["kind"] == "oranges") == ["blue", "green", "orange", "red"]


def test_reclaim_single_discrete(so):                                   # :97-113
    a = gen(so, "reclaim_resources", [], assigned=[discrete("apple", 2)])["available"]
    assert a == [discrete("apple", 2)]
    a = gen(so, "reclaim_resources", a, assigned=[discrete("apple", 2)])["available"]
    assert a == [discrete("apple", 4)]


def test_reclaim_multiple_discrete(so):                                 # :115-139
    a = gen(so, "reclaim_resources", [discrete("apple", 3), discrete("banana", 2)],
            assigned=[discrete("orange", 4), discrete("apple", 2)])["available"]
    assert {g["kind"]: g["value"] for g in a} == {"apple": 5, "orange": 4, "banana": 2}


def test_reclaim_str(so):                                               # :141-182
    a = gen(so, "reclaim_resources", [], assigned=nset("apple", "red", "orange"))["available"]
    assert len(a) == 2
    a = gen(so, "reclaim_resources", a, assigned=nset("apple", "blue", "red"))["available"]
    assert len(a) == 4 and sorted(g["named"] for g in a) == ["blue", "orange", "red", "red"]
    a = gen(so, "reclaim_resources", nset("orange", "green"),
            assigned=nset("apple", "red", "orange") + nset("orange", "red", "orange"))["available"]
    assert len(a) == 5 and sorted(g["named"] for g in a if g["kind"] == "orange") == ["green", "orange", "red"]


def test_named_want_zero_claims_every_member(so):
    # selectNodeResources never hits its `len(nrs) == tr.Value` exit for Value 0 (resource_management.go:58-63)
    r = gen(so, "claim", nset("gpu", "a", "b", "c"), reservations=[discrete("gpu", 0)])
    assert r["ok"] and r["available"] == [] and len(r["assigned"]) == 3


def test_has_resource_discrete(so):                                     # validate_test.go:10-29 (the three HasResource cases)
    # The path asks the same question through HasEnough (validate.go:24-51, called by ResourceFilter.Check filter.go:96-103);
    # for a discrete want on a discrete node resource the two functions agree: want <= available.
    for want, expect in ((1, True), (5, True), (6, False)):
        assert gen(so, "has_enough", [discrete("apple", 5)], reservations=[discrete("apple", want)])["enough"] is expect
    # HasEnough's own branches: a kind the node does not have, and a named set counted by its members (:36-47)
    assert gen(so, "has_enough", [discrete("apple", 5)], reservations=[discrete("pear", 1)])["enough"] is False
    assert gen(so, "has_enough", nset("apple", "red", "orange", "blue"), reservations=[discrete("apple", 3)])["enough"] is True
    assert gen(so, "has_enough", nset("apple", "red", "orange", "blue"), reservations=[discrete("apple", 4)])["enough"] is False


def _kinds(a):
    out = {}
    for g in a:
        out.setdefault(g["kind"], []).append(g)
    return out


def test_reclaim_resources_mixed(so):                                   # resource_management_test.go:187-230 TestReclaimResources
    node_res = nset("orange", "green", "blue") + [discrete("apple", 3)] + nset("banana", "red", "orange", "green") + [discrete("cake", 2)]
    assigned = nset("orange", "red", "orange") + nset("grape", "red", "orange") + [discrete("apple", 3), discrete("coffe", 2)]
    a = gen(so, "reclaim_resources", node_res, assigned=assigned)["available"]
    k = _kinds(a)
    assert len(a) == 12 and {n: len(v) for n, v in k.items()} == {"apple": 1, "orange": 4, "banana": 3, "cake": 1, "grape": 2, "coffe": 1}
    assert k["apple"][0]["value"] == 6 and k["cake"][0]["value"] == 2 and k["coffe"][0]["value"] == 2
    assert sorted(g["named"] for g in k["orange"]) == ["blue", "green", "orange", "red"]
    assert sorted(g["named"] for g in k["banana"]) == ["green", "orange", "red"]
    assert sorted(g["named"] for g in k["grape"]) == ["orange", "red"]


def _sanitize(so, node_res, available):
    """sanitize(nodeRes, &available) (resource_management.go:135-205) as the path reaches it: Reclaim with nothing to give
    back (NodeInfo.removeTask -> genericresource.Reclaim, nodeinfo.go:84-102, resource_management.go:74-91)."""
    return so.apply({"op": "generic", "fn": "reclaim", "available": resources(0, 0, available), "assigned": resources(0, 0, []),
                     "node": resources(0, 0, node_res)})["available"]


def test_sanitize_discrete(so):                                         # :232-271 TestSanitizeDiscrete
    assert _sanitize(so, [], [discrete("orange", 4)]) == []
    assert _sanitize(so, [discrete("orange", 6)], [discrete("orange", 4)]) == [discrete("orange", 4)]
    assert _sanitize(so, [discrete("orange", 4)], [discrete("orange", 4)]) == [discrete("orange", 4)]
    assert _sanitize(so, [discrete("orange", 2)], [discrete("orange", 4)]) == [discrete("orange", 2)]
    a = _sanitize(so, [discrete("orange", 2), discrete("banana", 6), discrete("cake", 6)],
                  [discrete("orange", 2), discrete("cake", 2), discrete("apple", 4), discrete("banana", 8)])
    assert [g["value"] for g in a] == [2, 2, 6] and [g["kind"] for g in a] == ["orange", "cake", "banana"]


def test_sanitize_str(so):                                              # :273-291 TestSanitizeStr
    avail = nset("apple", "red", "orange", "blue")
    assert _sanitize(so, [], avail) == []
    assert len(_sanitize(so, nset("apple", "red", "orange", "blue", "green"), avail)) == 3
    assert len(_sanitize(so, nset("apple", "red", "orange", "blue"), avail)) == 3
    assert len(_sanitize(so, nset("apple", "red", "orange"), avail)) == 2


def test_sanitize_change_discrete_to_set(so):                           # :293-339 TestSanitizeChangeDiscreteToSet
    assert _sanitize(so, nset("apple", "red"), [discrete("apple", 5)]) == nset("apple", "red")
    a = _sanitize(so, nset("apple", "red", "orange", "green"), [discrete("apple", 5)])
    assert sorted(g["named"] for g in a) == ["green", "orange", "red"]
    node_res = nset("apple", "red", "orange", "green") + nset("orange", "red", "orange", "green") + nset("cake", "red", "orange", "green")
    a = _sanitize(so, node_res, nset("apple", "green") + [discrete("cake", 3)] + nset("orange", "orange", "blue"))
    k = _kinds(a)
    assert len(a) == 5 and [g["named"] for g in k["apple"]] == ["green"] and [g["named"] for g in k["orange"]] == ["orange"]
    assert sorted(g["named"] for g in k["cake"]) == ["green", "orange", "red"]


def test_sanitize_change_set_to_discrete(so):                           # :341-373 TestSanitizeChangeSetToDiscrete
    assert _sanitize(so, [discrete("apple", 5)], nset("apple", "red")) == [discrete("apple", 5)]
    assert _sanitize(so, [discrete("apple", 5)], nset("apple", "red", "orange", "green")) == [discrete("apple", 5)]
    a = _sanitize(so, [discrete("apple", 5), discrete("orange", 3), discrete("cake", 1)],
                  [discrete("apple", 5), discrete("cake", 2)] + nset("orange", "orange", "blue"))
    assert {g["kind"]: g["value"] for g in a} == {"apple": 5, "orange": 3, "cake": 1} and len(a) == 3
