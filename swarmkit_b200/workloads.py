"""Seeded synthetic workloads for BASELINE.json's configs, in the flat ABI.

SURVEY.md 8(d): PRNG = SplitMix64, seed 0x5EED0000 + config number.  Node rows
are generated directly in canonical (sorted node-ID) order: attributes are iid,
so drawing random 25-char IDs (identity.NewID, identity/randomid.go) and sorting
them would only permute an iid sample.  Task IDs are likewise random, so the
canonical task order of one-offs (ascending task ID) is a uniform shuffle of the
service membership.

Shapes follow the reference's own benchmark, benchScheduler
(manager/scheduler/scheduler_test.go:3378-3468): one-off tasks have no
SpecVersion, so every task is its own k=1 group.
"""
from __future__ import annotations

import numpy as np

from . import abi
from .abi import NodeTable, Tick

_GOLD = np.uint64(0x9E3779B97F4A7C15)
GIB = 1 << 30
MIB = 1 << 20


class SplitMix64:
    """Vectorised SplitMix64 stream (counter mode: element i depends only on seed and i)."""

    def __init__(self, seed: int):
        self.seed = np.uint64(seed)
        self.pos = 0

    def next(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            i = np.arange(self.pos + 1, self.pos + n + 1, dtype=np.uint64)
            z = self.seed + i * _GOLD
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        self.pos += n
        return z

    def below(self, n: int, bound: int) -> np.ndarray:
        return (self.next(n) % np.uint64(bound)).astype(np.int64)

    def uniform(self, n: int) -> np.ndarray:
        return (self.next(n) >> np.uint64(11)).astype(np.float64) / float(1 << 53)

    def choice(self, n: int, probs) -> np.ndarray:
        cdf = np.cumsum(np.asarray(probs, dtype=np.float64))
        return np.searchsorted(cdf / cdf[-1], self.uniform(n), side="right").astype(np.int64)

    def permutation(self, n: int) -> np.ndarray:
        return np.argsort(self.next(n), kind="stable")


def _rows(n: int) -> np.ndarray:
    rows = np.zeros(n, abi.node_row_dt)
    rows["node_idx"] = np.arange(n, dtype=np.uint32)
    rows["flags"] = abi.PE_NODE_VALID | abi.PE_NODE_READY
    return rows


def _groups(n: int) -> np.ndarray:
    g = np.zeros(n, abi.group_dt)
    g["filter_mask"] = 1 << abi.PE_F_READY
    g["log_plugin"] = abi.PE_NONE
    return g


class Workload:
    def __init__(self, name: str, nodes: NodeTable, tick: Tick, meta: dict):
        self.name, self.nodes, self.tick, self.meta = name, nodes, tick, meta

    @property
    def n_nodes(self) -> int:
        return int(self.nodes.rows.size)


# --------------------------------------------------------------------------- cfg1
def cfg1(n_nodes: int = 100, n_tasks: int = 1000) -> Workload:
    """cmd/swarm-bench shape: one replicated service, identical replicas, no constraints."""
    rows = _rows(n_nodes)
    g = _groups(1)
    g["svc_id"], g["n_tasks"], g["task_off"] = 0, n_tasks, 0
    flags = np.full(n_tasks, abi.PE_T_COUNTS, np.uint8)
    return Workload("cfg1", NodeTable(rows), Tick(g, flags), {"mode": "grouped", "services": 1})


# --------------------------------------------------------------------------- cfg2
def cfg2(mode: str = "oneoff", n_nodes: int = 10_000, n_tasks: int = 100_000, n_services: int = 100,
         seed: int = 0x5EED0002) -> Workload:
    """CPU + memory ResourceFilter only, spread strategy."""
    rng = SplitMix64(seed)
    rows = _rows(n_nodes)
    rows["cpu_avail"] = np.array([4, 8, 16, 32, 64], np.int64)[rng.below(n_nodes, 5)] * 1_000_000_000
    rows["mem_avail"] = np.array([8, 16, 32, 64, 128, 256], np.int64)[rng.below(n_nodes, 6)] * GIB
    fm = (1 << abi.PE_F_READY) | (1 << abi.PE_F_RESOURCE)
    if mode == "grouped":
        per = n_tasks // n_services
        g = _groups(n_services)
        g["svc_id"] = np.arange(n_services)
        g["n_tasks"] = per
        g["task_off"] = np.arange(n_services) * per
        g["filter_mask"] = fm
        g["cpu_res"] = (1 + rng.below(n_services, 20)) * 100_000_000
        g["mem_res"] = (1 + rng.below(n_services, 64)) * 64 * MIB
        flags = np.full(per * n_services, abi.PE_T_COUNTS, np.uint8)
    else:
        g = _groups(n_tasks)
        g["svc_id"] = 0  # benchScheduler tasks carry no ServiceID (scheduler_test.go:3428-3439)
        g["n_tasks"] = 1
        g["task_off"] = np.arange(n_tasks)
        g["filter_mask"] = fm
        g["cpu_res"] = (1 + rng.below(n_tasks, 20)) * 100_000_000
        g["mem_res"] = (1 + rng.below(n_tasks, 64)) * 64 * MIB
        flags = np.full(n_tasks, abi.PE_T_COUNTS, np.uint8)
    return Workload(f"cfg2-{mode}", NodeTable(rows), Tick(g, flags), {"mode": mode, "services": n_services})


# --------------------------------------------------------------------------- cfg3
# folded value ids (shared dictionary; 0 = "")
_ROLE = {"worker": 1, "manager": 2}
_OS = {"linux": 3, "windows": 4}
_ARCH = {"amd64": 5, "arm64": 6, "x86_64": 7}
_LABELS = [("zone", 16), ("rack", 64), ("disk", 2), ("gpu", 4), ("tier", 4), ("env", 3)]
_LABEL_BASE = 100  # value id of label k, value v = _LABEL_BASE + 100*k + v
# exact-string ids for PlatformFilter (arch already normalised, filter.go:291-306)
_PLAT_OS = {"linux": 1, "windows": 2}
_PLAT_ARCH = {"amd64": 1, "arm64": 2}
_P_PRESENT = 0.9


def _cfg3_nodes(rng: SplitMix64, n: int):
    rows = _rows(n)
    rows["flags"] |= abi.PE_NODE_HAS_PLATFORM | abi.PE_NODE_HAS_ENGINE
    role = np.where(rng.uniform(n) < 0.03, _ROLE["manager"], _ROLE["worker"])
    plat = rng.choice(n, [0.70, 0.20, 0.10])  # linux/amd64, linux/arm64, windows/amd64
    as_x86 = rng.uniform(n) < 0.10
    os_fold = np.where(plat == 2, _OS["windows"], _OS["linux"])
    arch_fold = np.where(plat == 1, _ARCH["arm64"], np.where(as_x86, _ARCH["x86_64"], _ARCH["amd64"]))
    rows["os_id"] = np.where(plat == 2, _PLAT_OS["windows"], _PLAT_OS["linux"])
    rows["arch_id"] = np.where(plat == 1, _PLAT_ARCH["arm64"], _PLAT_ARCH["amd64"])
    n_cols = 3 + len(_LABELS)
    keys = np.zeros((n, n_cols), np.uint32)
    vals = np.zeros((n, n_cols), np.uint32)
    keys[:, 0], vals[:, 0] = abi.PE_ATTR_ROLE, role
    keys[:, 1], vals[:, 1] = abi.PE_ATTR_OS, os_fold
    keys[:, 2], vals[:, 2] = abi.PE_ATTR_ARCH, arch_fold
    for k, (_, card) in enumerate(_LABELS):
        present = rng.uniform(n) < _P_PRESENT
        v = rng.below(n, card)
        keys[:, 3 + k] = abi.PE_ATTR_FIRST_LABEL + k
        vals[:, 3 + k] = np.where(present, _LABEL_BASE + 100 * k + v, 0)
    attrs = np.zeros(n * n_cols, abi.kv32_dt)
    attrs["key"] = keys.reshape(-1)
    attrs["value"] = vals.reshape(-1)
    rows["attr_off"] = np.arange(n, dtype=np.uint32) * n_cols
    rows["attr_cnt"] = n_cols
    return rows, attrs


def _cfg3_constraints(rng: SplitMix64, n_services: int, with_platform: np.ndarray):
    """8 expressions per service, rejection-sampled to 5-20 % node selectivity.

    Selectivity is exact under the generator's own distribution: per key keep the
    set of still-admissible values (id 0 = label absent, which fails `==` and
    passes `!=`, constraint.go:173-199) and multiply the per-key probabilities;
    (os, arch) is one composite key because the node generator draws them jointly,
    and a service with Placement.Platforms=[{amd64,linux}] additionally keeps only
    linux x {amd64, x86_64} (PlatformFilter normalises x86_64, filter.go:291-296).
    """
    L, Wn = _OS["linux"], _OS["windows"]
    A, X, R = _ARCH["amd64"], _ARCH["x86_64"], _ARCH["arm64"]
    plat_dom = {(L, A): .63, (L, X): .07, (L, R): .20, (Wn, A): .09, (Wn, X): .01}
    keys = [(abi.PE_ATTR_ROLE, {_ROLE["worker"]: .97, _ROLE["manager"]: .03}),
            (abi.PE_ATTR_OS, {L: .9, Wn: .1}),
            (abi.PE_ATTR_ARCH, {A: .72, R: .20, X: .08})]
    for k, (_, card) in enumerate(_LABELS):
        dom = {_LABEL_BASE + 100 * k + v: _P_PRESENT / card for v in range(card)}
        dom[0] = 1.0 - _P_PRESENT
        keys.append((abi.PE_ATTR_FIRST_LABEL + k, dom))

    def selectivity(exprs, platform):
        sel = 1.0
        pd = dict(plat_dom)
        if platform:
            pd = {k: v for k, v in pd.items() if k[0] == L and k[1] in (A, X)}
        for (c, val, neq) in exprs:
            if c == abi.PE_ATTR_OS:
                pd = {k: v for k, v in pd.items() if (k[0] == val) != bool(neq)}
            elif c == abi.PE_ATTR_ARCH:
                pd = {k: v for k, v in pd.items() if (k[1] == val) != bool(neq)}
        sel *= sum(pd.values())
        for (col, dom) in keys:
            if col in (abi.PE_ATTR_OS, abi.PE_ATTR_ARCH):
                continue
            allowed = dict(dom)
            for (c, val, neq) in exprs:
                if c != col:
                    continue
                if neq:
                    allowed.pop(val, None)
                else:
                    allowed = {val: allowed[val]} if val in allowed else {}
            sel *= sum(allowed.values())
        return sel

    cons = np.zeros(n_services * 8, abi.constraint_dt)
    est = np.zeros(n_services)
    for s in range(n_services):
        best = None
        plat = bool(with_platform[s])
        for _attempt in range(64):
            u = rng.uniform(24)
            exprs = []
            for e in range(8):
                col, dom = keys[int(u[3 * e] * len(keys))]
                vals = [v for v in dom if v != 0]
                val = vals[int(u[3 * e + 1] * len(vals))]
                neq = 0 if u[3 * e + 2] < 0.5 else 1
                if selectivity(exprs + [(col, val, neq)], plat) < 0.05:  # over-constrained: complementary operator
                    neq ^= 1
                exprs.append((col, val, neq))
            sel = selectivity(exprs, plat)
            score = 0.0 if 0.05 <= sel <= 0.20 else min(abs(sel - 0.05), abs(sel - 0.20))
            if best is None or score < best[0]:
                best = (score, sel, exprs)
            if score == 0.0:
                break
        est[s] = best[1]
        for e, (col, val, neq) in enumerate(best[2]):
            cons[s * 8 + e] = (col, val, neq)
    return cons, est


def cfg3(mode: str = "oneoff", n_nodes: int = 100_000, n_tasks: int = 1_000_000, n_services: int = 1000,
         seed: int = 0x5EED0003) -> Workload:
    """8 constraint expressions (node.labels / role / platform) + PlatformFilter on half the services + spread."""
    rng = SplitMix64(seed)
    rows, attrs = _cfg3_nodes(rng, n_nodes)
    with_platform = rng.uniform(n_services) < 0.5
    cons, est = _cfg3_constraints(rng, n_services, with_platform)
    plats = np.zeros(1, abi.platform_dt)
    plats[0] = (_PLAT_OS["linux"], _PLAT_ARCH["amd64"])  # Placement.Platforms = [{amd64, linux}]
    fm_base = (1 << abi.PE_F_READY) | (1 << abi.PE_F_CONSTRAINT)
    per = n_tasks // n_services
    if mode == "grouped":
        g = _groups(n_services)
        svc = np.arange(n_services)
        g["n_tasks"] = per
        g["task_off"] = svc * per
    else:
        g = _groups(per * n_services)
        svc = np.repeat(np.arange(n_services), per)[rng.permutation(per * n_services)]
        g["n_tasks"] = 1
        g["task_off"] = np.arange(per * n_services)
    g["svc_id"] = svc
    g["con_off"] = svc * 8
    g["con_cnt"] = 8
    g["filter_mask"] = fm_base | (with_platform[svc].astype(np.uint32) << abi.PE_F_PLATFORM)
    g["plat_off"] = 0
    g["plat_cnt"] = with_platform[svc].astype(np.uint32)
    flags = np.full(per * n_services, abi.PE_T_COUNTS, np.uint8)
    return Workload(f"cfg3-{mode}", NodeTable(rows, attrs=attrs), Tick(g, flags, cons=cons, plats=plats),
                    {"mode": mode, "services": n_services, "est_selectivity_mean": float(est.mean())})


# --------------------------------------------------------------------------- cfg4
def cfg4(mode: str = "oneoff", n_nodes: int = 1_000_000, n_tasks: int = 1_000_000, n_services: int = 1000,
         seed: int = 0x5EED0004) -> Workload:
    """cfg3 + cfg2 resources + 3 generic kinds (2 discrete, 1 named) + host ports + MaxReplicas."""
    w = cfg3(mode, n_nodes, n_tasks, n_services, seed)
    rng = SplitMix64(seed ^ 0xABCDEF)
    rows, g = w.nodes.rows, w.tick.groups
    rows["cpu_avail"] = np.array([4, 8, 16, 32, 64], np.int64)[rng.below(n_nodes, 5)] * 1_000_000_000
    rows["mem_avail"] = np.array([8, 16, 32, 64, 128, 256], np.int64)[rng.below(n_nodes, 6)] * GIB
    # generic kinds: 0,1 discrete (present w.p. 0.5, value 1..8), 2 named (present w.p. 0.3, 1..4 members)
    gk = np.zeros((n_nodes, 3), abi.kv64_dt)
    cnt = np.zeros(n_nodes, np.uint32)
    for kind in range(3):
        present = rng.uniform(n_nodes) < (0.5 if kind < 2 else 0.3)
        value = 1 + rng.below(n_nodes, 8 if kind < 2 else 4)
        typ = abi.PE_GEN_DISCRETE if kind < 2 else abi.PE_GEN_NAMED
        gk["key"][:, kind] = kind
        gk["value"][:, kind] = np.where(present, (value << 2) | typ, 0)
    rows["gen_off"] = np.arange(n_nodes, dtype=np.uint32) * 3
    rows["gen_cnt"] = 3
    del cnt
    per_service = n_services
    s_cpu = (1 + rng.below(per_service, 20)) * 100_000_000
    s_mem = (1 + rng.below(per_service, 64)) * 64 * MIB
    s_gen_kind = rng.below(per_service, 3)
    s_gen_on = rng.uniform(per_service) < 0.3
    s_gen_val = 1 + rng.below(per_service, 2)
    s_port_on = rng.uniform(per_service) < 0.2
    s_nports = 1 + rng.below(per_service, 2)
    s_port0 = rng.below(per_service, 64)
    s_port1 = rng.below(per_service, 64)
    s_max_on = rng.uniform(per_service) < 0.3
    s_max = np.array([1, 2, 4], np.uint64)[rng.below(per_service, 3)]
    gens = np.zeros(per_service, abi.generic_want_dt)
    gens["kind"], gens["value"] = s_gen_kind, s_gen_val
    ports = np.zeros(per_service * 2, np.uint32)
    ports[0::2], ports[1::2] = s_port0, np.where(s_port1 == s_port0, (s_port0 + 1) % 64, s_port1)
    svc = g["svc_id"].astype(np.int64)
    g["cpu_res"], g["mem_res"] = s_cpu[svc], s_mem[svc]
    g["gen_off"], g["gen_cnt"] = svc, s_gen_on[svc].astype(np.uint32)
    g["port_off"], g["port_cnt"] = svc * 2, np.where(s_port_on[svc], s_nports[svc], 0)
    g["max_replicas"] = np.where(s_max_on[svc], s_max[svc], 0)
    g["filter_mask"] |= (1 << abi.PE_F_RESOURCE)
    g["filter_mask"] |= (s_port_on[svc].astype(np.uint32) << abi.PE_F_HOSTPORT)
    g["filter_mask"] |= (s_max_on[svc].astype(np.uint32) << abi.PE_F_MAXREPLICAS)
    nodes = NodeTable(rows, attrs=w.nodes.attrs, gens=gk.reshape(-1))
    tick = Tick(g, w.tick.task_flags, gens=gens, cons=w.tick.cons, plats=w.tick.plats, ports=ports)
    return Workload(f"cfg4-{mode}", nodes, tick, dict(w.meta))


# --------------------------------------------------------------------------- cfg5
def cfg5(mode: str = "grouped", n_nodes: int = 100_000, n_tasks: int = 500_000, n_services: int = 5000,
         tasks_per_node: int = 50, drain_fraction: float = 0.1, seed: int = 0x5EED0005) -> Workload:
    """Streaming reschedule storm (BASELINE.json configs[4]): a steady state of
    n_nodes x tasks_per_node running tasks across n_services services with cpu/memory reservations; a tenth of the
    nodes is drained and their tasks re-enter as PENDING in ONE tick.  The reference flow is
    replicated.handleNodeChange -> restartTasksByNodeID (manager/orchestrator/replicated/tasks.go:86-117) ->
    restart.Supervisor.Restart -> orchestrator.NewTask: replacements keep ServiceID + SpecVersion, so the tick is
    GROUPED (one group per service that lost tasks); the drained nodes fail ReadyFilter; the old tasks were flipped to
    DesiredState SHUTDOWN, so they left the spread counters (nodeinfo.go:110-122) but still hold their resources.
    n_tasks is the expected number of re-entering tasks = drain_fraction x n_nodes x tasks_per_node (it fixes
    tasks_per_node when given); mode is accepted for symmetry (the storm is always grouped)."""
    rng = SplitMix64(seed)
    if n_tasks and drain_fraction > 0:
        tasks_per_node = max(1, int(round(n_tasks / (drain_fraction * n_nodes))))
    rows = _rows(n_nodes)
    # capacity: every node could hold about twice its steady-state share, so the survivors absorb the storm
    s_cpu = (1 + rng.below(n_services, 10)).astype(np.int64) * 100_000_000          # 0.1 .. 1.0 CPU
    s_mem = (1 + rng.below(n_services, 16)).astype(np.int64) * 64 * MIB             # 64 MiB .. 1 GiB
    # steady state: every node runs tasks_per_node tasks of distinct random services
    svc_of = rng.below(n_nodes * tasks_per_node, n_services).reshape(n_nodes, tasks_per_node).astype(np.int64)
    svc_of.sort(axis=1)
    used_cpu, used_mem = s_cpu[svc_of].sum(1), s_mem[svc_of].sum(1)
    rows["cpu_avail"] = used_cpu + (1 + rng.below(n_nodes, 4)).astype(np.int64) * int(0.55 * tasks_per_node * 100_000_000)
    rows["mem_avail"] = used_mem + (1 + rng.below(n_nodes, 4)).astype(np.int64) * int(8.5 * tasks_per_node * 64 * MIB)
    drained = rng.uniform(n_nodes) < drain_fraction
    live = ~drained
    rows["flags"] = np.where(drained, abi.PE_NODE_VALID, abi.PE_NODE_VALID | abi.PE_NODE_READY)
    rows["cpu_avail"] -= used_cpu      # running tasks (and, on the drained nodes, the ones still shutting down) hold their share
    rows["mem_avail"] -= used_mem
    rows["total_tasks"] = np.where(live, tasks_per_node, 0)
    # per-node service counters of the live nodes (duplicates of a service on one node add up)
    li = np.flatnonzero(live)
    keys = (li[:, None] * np.int64(n_services) + svc_of[li]).reshape(-1)
    uniq, cnt = np.unique(keys, return_counts=True)
    svcs = np.zeros(uniq.size, abi.kv32_dt)
    svcs["key"], svcs["value"] = (uniq % n_services).astype(np.uint32), cnt.astype(np.uint32)
    node_of = uniq // n_services
    per_node = np.bincount(node_of, minlength=n_nodes)
    rows["svc_cnt"] = per_node
    rows["svc_off"] = np.concatenate(([0], np.cumsum(per_node)[:-1]))
    # the storm: the drained nodes' tasks, one group per service
    lost = np.bincount(svc_of[drained].reshape(-1), minlength=n_services)
    sv = np.flatnonzero(lost)
    g = _groups(sv.size)
    g["svc_id"] = sv
    g["n_tasks"] = lost[sv]
    g["task_off"] = np.concatenate(([0], np.cumsum(lost[sv])[:-1]))
    g["filter_mask"] = (1 << abi.PE_F_READY) | (1 << abi.PE_F_RESOURCE)
    g["cpu_res"], g["mem_res"] = s_cpu[sv], s_mem[sv]
    flags = np.full(int(lost.sum()), abi.PE_T_COUNTS, np.uint8)
    return Workload("cfg5-grouped", NodeTable(rows, svcs=svcs), Tick(g, flags),
                    {"mode": "grouped", "services": n_services, "drained_nodes": int(drained.sum()), "running_tasks": int(live.sum()) * tasks_per_node})


def by_name(name: str, **kw) -> Workload:
    base, _, mode = name.partition("-")
    fn = {"cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}[base]
    if base == "cfg1":
        return fn(**kw)
    return fn(mode or ("grouped" if base == "cfg5" else "oneoff"), **kw)
