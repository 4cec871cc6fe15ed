"""Random flat-ABI workloads that exercise every filter, the rank key, ties,
rotated tie-break, non-counting tasks, removed nodes and mixed group sizes.
Used by the parity tests (GPU engine vs CPU oracle)."""
from __future__ import annotations

import numpy as np

from swarmkit_b200 import abi
from swarmkit_b200.abi import NodeTable, Tick


def random_nodes(rng: np.random.Generator, n: int, n_labels=3, n_svc=4, n_gen=3, n_slots=40, tight=False):
    rows = np.zeros(n, abi.node_row_dt)
    rows["node_idx"] = np.arange(n)
    valid = rng.random(n) < 0.95
    ready = rng.random(n) < 0.9
    flags = np.where(valid, abi.PE_NODE_VALID, 0) | np.where(ready, abi.PE_NODE_READY, 0)
    flags |= np.where(rng.random(n) < 0.9, abi.PE_NODE_HAS_PLATFORM, 0)
    flags |= np.where(rng.random(n) < 0.8, abi.PE_NODE_HAS_ENGINE, 0)
    flags |= np.where(rng.random(n) < 0.5, abi.PE_NODE_HAS_LOGPLUGIN, 0)
    ipkind = rng.integers(0, 10, n)  # 0: invalid, 1: v6, else v4
    flags |= np.where(ipkind > 0, abi.PE_NODE_IP_VALID, 0) | np.where(ipkind > 1, abi.PE_NODE_IP_V4, 0)
    rows["flags"] = np.where(valid, flags, 0)
    rows["os_id"] = rng.integers(1, 4, n)
    rows["arch_id"] = rng.integers(1, 4, n)
    scale = 4 if tight else 40
    rows["cpu_avail"] = rng.integers(-2, scale, n) * 1_000_000_000
    rows["mem_avail"] = rng.integers(-1, scale, n) * (1 << 30)
    ip = np.zeros((n, 4), np.uint32)
    v4 = ipkind > 1
    ip[:, 2] = np.where(v4, 0xFFFF, rng.integers(0, 4, n))
    ip[:, 3] = np.where(v4, (10 << 24) | (rng.integers(0, 4, n) << 16) | (rng.integers(0, 4, n) << 8) | rng.integers(0, 8, n),
                        rng.integers(0, 16, n))
    ip[:, 0] = np.where(v4, 0, 0x20010DB8)
    rows["ip"] = ip
    # attributes
    n_cols = abi.PE_ATTR_FIRST_LABEL + n_labels
    keys = np.tile(np.arange(n_cols, dtype=np.uint32), (n, 1))
    vals = rng.integers(0, 4, (n, n_cols)).astype(np.uint32)
    attrs = np.zeros(n * n_cols, abi.kv32_dt)
    attrs["key"], attrs["value"] = keys.reshape(-1), vals.reshape(-1)
    rows["attr_off"], rows["attr_cnt"] = np.arange(n) * n_cols, n_cols
    # generic resources
    gens = np.zeros(n * n_gen, abi.kv64_dt)
    typ = rng.integers(0, 3, (n, n_gen))
    cnt = rng.integers(0, 5, (n, n_gen))
    cnt = np.where(typ == abi.PE_GEN_NAMED, np.maximum(cnt, 1), cnt)
    cell = np.where(typ == abi.PE_GEN_ABSENT, 0, (cnt << 2) | typ)
    gens["key"] = np.tile(np.arange(n_gen, dtype=np.uint32), (n, 1)).reshape(-1)
    gens["value"] = cell.reshape(-1)
    rows["gen_off"], rows["gen_cnt"] = np.arange(n) * n_gen, n_gen
    # per-service counts + totals
    sc = (rng.random((n, n_svc)) < 0.3) * rng.integers(1, 4, (n, n_svc))
    svcs = np.zeros(n * n_svc, abi.kv32_dt)
    svcs["key"] = np.tile(np.arange(n_svc, dtype=np.uint32), (n, 1)).reshape(-1)
    svcs["value"] = sc.reshape(-1)
    rows["svc_off"], rows["svc_cnt"] = np.arange(n) * n_svc, n_svc
    rows["total_tasks"] = sc.sum(1) + (rng.random(n) < 0.3) * rng.integers(0, 3, n)
    # ports / plugins
    ports, plugs, po, pc, qo, qc = [], [], [], [], [], []
    for i in range(n):
        p = np.flatnonzero(rng.random(n_slots) < 0.15)
        q = np.flatnonzero(rng.random(n_slots) < 0.6)
        po.append(len(ports)); pc.append(p.size); ports.extend(p.tolist())
        qo.append(len(plugs)); qc.append(q.size); plugs.extend(q.tolist())
    rows["port_off"], rows["port_cnt"], rows["plug_off"], rows["plug_cnt"] = po, pc, qo, qc
    return NodeTable(rows, attrs, gens, svcs, np.array(ports, np.uint32), np.array(plugs, np.uint32))


def random_tick(rng: np.random.Generator, n_nodes: int, n_groups: int, n_labels=3, n_svc=4, n_gen=3, n_slots=40,
                p_oneoff=0.6, kmax=40, feature_p=0.35, rotate=True, p_leaf=0.0):
    g = np.zeros(n_groups, abi.group_dt)
    g["log_plugin"] = abi.PE_NONE
    gens, cons, ips, plats, ports, plugs, fails = [], [], [], [], [], [], []
    task_off = 0
    flags = []
    n_cols = abi.PE_ATTR_FIRST_LABEL + n_labels
    for i in range(n_groups):
        k = 1 if rng.random() < p_oneoff else int(rng.integers(2, kmax + 1))
        fm = 1 << abi.PE_F_READY
        g[i]["svc_id"] = rng.integers(0, n_svc)
        g[i]["n_tasks"], g[i]["task_off"] = k, task_off
        task_off += k
        flags.extend((rng.random(k) < 0.93).astype(np.uint8).tolist())
        if rng.random() < feature_p:  # resources
            fm |= 1 << abi.PE_F_RESOURCE
            g[i]["cpu_res"] = rng.integers(0, 4) * 500_000_000
            g[i]["mem_res"] = rng.integers(0, 4) * (1 << 29)
            ng = int(rng.integers(0, 3))
            g[i]["gen_off"], g[i]["gen_cnt"] = len(gens), ng
            for _ in range(ng):
                gens.append((rng.integers(0, n_gen), 0, rng.integers(0, 3)))
        if rng.random() < feature_p:  # plugins
            fm |= 1 << abi.PE_F_PLUGIN
            nq = int(rng.integers(0, 3))
            g[i]["plug_off"], g[i]["plug_cnt"] = len(plugs), nq
            plugs.extend(rng.integers(0, n_slots, nq).tolist())
            if rng.random() < 0.5:
                g[i]["flags"] |= abi.PE_G_LOG_DRIVER
                g[i]["log_plugin"] = rng.integers(0, n_slots)
        if rng.random() < feature_p + 0.2:  # constraints
            fm |= 1 << abi.PE_F_CONSTRAINT
            nc = int(rng.integers(0, 4))
            g[i]["con_off"], g[i]["con_cnt"] = len(cons), nc
            for _ in range(nc):
                cons.append((rng.integers(0, n_cols), rng.integers(0, 4), rng.integers(0, 2) if rng.random() < 0.7 else 1))
            if rng.random() < 0.25:
                g[i]["ip_off"], g[i]["ip_cnt"] = len(ips), 1
                if rng.random() < 0.5:  # CIDR 10.a.0.0/16
                    net = (0, 0, 0xFFFF, (10 << 24) | (int(rng.integers(0, 4)) << 16))
                    mask = (0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFF0000)
                    ips.append((net, mask, rng.integers(0, 2), 1, 1))
                else:  # single address
                    net = (0, 0, 0xFFFF, (10 << 24) | (int(rng.integers(0, 4)) << 16) | (int(rng.integers(0, 4)) << 8) | int(rng.integers(0, 8)))
                    ips.append((net, (0xFFFFFFFF,) * 4, rng.integers(0, 2), 0, 0))
            if rng.random() < 0.05:
                g[i]["flags"] |= abi.PE_G_CONSTRAINT_NEVER
        if p_leaf and rng.random() < p_leaf:  # one leaf visit of a placement-preference tree: the terms follow the constraints
            if g[i]["con_cnt"] == 0:
                g[i]["con_off"] = len(cons)
            nl = int(rng.integers(1, 3))
            g[i]["leaf_cnt"] = nl
            for _ in range(nl):
                cons.append((rng.integers(abi.PE_ATTR_FIRST_LABEL, n_cols), rng.integers(0, 4), 0))
        if rng.random() < feature_p:  # platforms
            fm |= 1 << abi.PE_F_PLATFORM
            npf = int(rng.integers(1, 6))
            g[i]["plat_off"], g[i]["plat_cnt"] = len(plats), npf
            for _ in range(npf):
                plats.append((rng.integers(0, 4), rng.integers(0, 4)))
        if rng.random() < feature_p:  # host ports
            fm |= 1 << abi.PE_F_HOSTPORT
            npp = int(rng.integers(1, 3))
            g[i]["port_off"], g[i]["port_cnt"] = len(ports), npp
            ports.extend(rng.integers(0, n_slots, npp).tolist())
        if rng.random() < feature_p:  # max replicas
            fm |= 1 << abi.PE_F_MAXREPLICAS
            g[i]["max_replicas"] = rng.integers(1, 5)
        if rng.random() < 0.2:  # recent failures
            nf = int(rng.integers(1, max(2, n_nodes // 4)))
            idx = np.sort(rng.choice(n_nodes, size=min(nf, n_nodes), replace=False))
            g[i]["fail_off"], g[i]["fail_cnt"] = len(fails), idx.size
            fails.extend(zip(idx.tolist(), rng.integers(1, 9, idx.size).tolist()))
        if rotate and rng.random() < 0.5:
            g[i]["tie_start"] = rng.integers(0, n_nodes)
        g[i]["filter_mask"] = fm
    return Tick(g, np.array(flags, np.uint8),
                gens=np.array(gens, abi.generic_want_dt) if gens else None,
                cons=np.array(cons, abi.constraint_dt) if cons else None,
                ips=np.array(ips, abi.ip_constraint_dt) if ips else None,
                plats=np.array(plats, abi.platform_dt) if plats else None,
                ports=np.array(ports, np.uint32), plugs=np.array(plugs, np.uint32),
                fails=np.array(fails, abi.node_fail_dt) if fails else None)


def compare_results(tick: Tick, a, b, what=""):
    """a, b = (out_node, out_fail).  Failure counters are compared for groups with unplaced tasks."""
    an, af = a
    bn, bf = b
    bad = np.flatnonzero(an != bn)
    assert bad.size == 0, f"{what}: {bad.size} placements differ, first task {bad[:5]}: {an[bad[:5]]} vs {bn[bad[:5]]}"
    g = tick.groups
    for i in range(g.size):
        t0, k = int(g[i]["task_off"]), int(g[i]["n_tasks"])
        if k and (an[t0:t0 + k] == abi.PE_NONE).any():
            assert (af[i] == bf[i]).all(), f"{what}: group {i} failure counters {af[i]} vs {bf[i]}"


def compare_state(e1, e2, n_nodes: int, n_svc: int, n_gen: int, n_slots: int, what=""):
    s1, s2 = e1.snapshot(0, n_nodes), e2.snapshot(0, n_nodes)
    for f in ("flags", "total_tasks", "cpu_avail", "mem_avail"):
        assert (s1[f] == s2[f]).all(), f"{what}: node column {f} differs"
    for s in range(n_svc):
        assert (e1.snapshot_service(s, 0, n_nodes) == e2.snapshot_service(s, 0, n_nodes)).all(), f"{what}: service {s} counts differ"
    for k in range(n_gen):
        assert (e1.snapshot_generic(k, 0, n_nodes) == e2.snapshot_generic(k, 0, n_nodes)).all(), f"{what}: generic kind {k} differs"
    for p in range(n_slots):
        assert (e1.snapshot_ports(p, 0, n_nodes) == e2.snapshot_ports(p, 0, n_nodes)).all(), f"{what}: port slot {p} differs"
