"""ctypes/numpy mirror of include/placement_engine.h (the C ABI of the engine).

`FlatABI` binds any shared library that exports the ABI under a symbol prefix.
The product (`PlacementEngine`, engine.py) binds libplacement.so with `pe_`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

PE_ABI_VERSION = 2
PE_NONE = 0xFFFFFFFF
PE_NUM_FILTERS = 8
(PE_F_READY, PE_F_RESOURCE, PE_F_PLUGIN, PE_F_CONSTRAINT, PE_F_PLATFORM, PE_F_HOSTPORT, PE_F_MAXREPLICAS,
 PE_F_VOLUMES) = range(8)

PE_NODE_VALID, PE_NODE_READY, PE_NODE_HAS_PLATFORM, PE_NODE_HAS_ENGINE = 0x01, 0x02, 0x04, 0x08
PE_NODE_HAS_LOGPLUGIN, PE_NODE_IP_VALID, PE_NODE_IP_V4 = 0x10, 0x20, 0x40
PE_GEN_ABSENT, PE_GEN_DISCRETE, PE_GEN_NAMED = 0, 1, 2
PE_ATTR_NODE_ID, PE_ATTR_HOSTNAME, PE_ATTR_ROLE, PE_ATTR_OS, PE_ATTR_ARCH, PE_ATTR_FIRST_LABEL = 0, 1, 2, 3, 4, 5
PE_G_CONSTRAINT_NEVER, PE_G_LOG_DRIVER = 0x1, 0x2
PE_T_COUNTS = 0x1
PE_CFG_NO_SPECULATION = 0x1
PE_CFG_ORDERED_ONLY = 0x2

PE_OK, PE_ERR_INVALID, PE_ERR_CUDA, PE_ERR_NOMEM, PE_ERR_UNSUPPORTED, PE_ERR_NO_DEVICE, PE_ERR_OVERFLOW = range(7)


def gen_encode(count: int, typ: int) -> int:
    return (int(count) << 2) | typ


kv32_dt = np.dtype([("key", "<u4"), ("value", "<u4")], align=True)
kv64_dt = np.dtype([("key", "<u4"), ("pad", "<u4"), ("value", "<i8")], align=True)
node_row_dt = np.dtype([
    ("node_idx", "<u4"), ("flags", "<u4"), ("os_id", "<u4"), ("arch_id", "<u4"),
    ("cpu_avail", "<i8"), ("mem_avail", "<i8"), ("total_tasks", "<u4"),
    ("attr_off", "<u4"), ("attr_cnt", "<u4"), ("gen_off", "<u4"), ("gen_cnt", "<u4"),
    ("svc_off", "<u4"), ("svc_cnt", "<u4"), ("port_off", "<u4"), ("port_cnt", "<u4"),
    ("plug_off", "<u4"), ("plug_cnt", "<u4"), ("ip", "<u4", (4,)),
], align=True)
constraint_dt = np.dtype([("col", "<u4"), ("value", "<u4"), ("neq", "<u4")], align=True)
ip_constraint_dt = np.dtype([("net", "<u4", (4,)), ("mask", "<u4", (4,)), ("neq", "<u4"), ("is_cidr", "<u4"),
                             ("is_v4", "<u4")], align=True)
platform_dt = np.dtype([("os_id", "<u4"), ("arch_id", "<u4")], align=True)
generic_want_dt = np.dtype([("kind", "<u4"), ("pad", "<u4"), ("value", "<i8")], align=True)
node_fail_dt = np.dtype([("node_idx", "<u4"), ("count", "<u4")], align=True)
group_dt = np.dtype([
    ("svc_id", "<u4"), ("n_tasks", "<u4"), ("task_off", "<u4"), ("filter_mask", "<u4"),
    ("cpu_res", "<i8"), ("mem_res", "<i8"), ("max_replicas", "<u8"),
    ("gen_off", "<u4"), ("gen_cnt", "<u4"), ("con_off", "<u4"), ("con_cnt", "<u4"),
    ("ip_off", "<u4"), ("ip_cnt", "<u4"), ("plat_off", "<u4"), ("plat_cnt", "<u4"),
    ("port_off", "<u4"), ("port_cnt", "<u4"), ("plug_off", "<u4"), ("plug_cnt", "<u4"),
    ("log_plugin", "<u4"), ("fail_off", "<u4"), ("fail_cnt", "<u4"), ("tie_start", "<u4"), ("flags", "<u4"),
    ("leaf_cnt", "<u4"),
], align=True)
task_delta_dt = np.dtype([
    ("node_idx", "<u4"), ("svc_id", "<u4"), ("sign", "<i4"), ("counts", "<u4"), ("cpu", "<i8"), ("mem", "<i8"),
    ("gen_off", "<u4"), ("gen_cnt", "<u4"), ("port_off", "<u4"), ("port_cnt", "<u4"),
], align=True)
node_state_dt = np.dtype([("flags", "<u4"), ("total_tasks", "<u4"), ("cpu_avail", "<i8"), ("mem_avail", "<i8")],
                         align=True)

assert node_row_dt.itemsize == 96 and group_dt.itemsize == 112 and task_delta_dt.itemsize == 48


class pe_config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("node_capacity", C.c_uint32),
                ("flags", C.c_uint32), ("max_batch", C.c_uint32), ("rank", C.c_int32), ("world_size", C.c_int32),
                ("nccl_unique_id", C.c_void_p)]


class pe_tick(C.Structure):
    _fields_ = [
        ("groups", C.c_void_p), ("n_groups", C.c_uint32),
        ("task_flags", C.c_void_p), ("n_tasks", C.c_uint32),
        ("gens", C.c_void_p), ("n_gens", C.c_uint32),
        ("cons", C.c_void_p), ("n_cons", C.c_uint32),
        ("ips", C.c_void_p), ("n_ips", C.c_uint32),
        ("plats", C.c_void_p), ("n_plats", C.c_uint32),
        ("ports", C.c_void_p), ("n_ports", C.c_uint32),
        ("plugs", C.c_void_p), ("n_plugs", C.c_uint32),
        ("fails", C.c_void_p), ("n_fails", C.c_uint32),
    ]


class pe_stats(C.Structure):
    _fields_ = [("evals", C.c_uint64), ("evals_generic", C.c_uint64), ("scan_bytes", C.c_uint64),
                ("placements", C.c_uint64), ("fast_path", C.c_uint64), ("medium_path", C.c_uint64), ("slow_path", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("scan_launches", C.c_uint64), ("scan_ms", C.c_double),
                ("sequencer_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double), ("run_ms", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("seq_cycles_fast", C.c_uint64),
                ("seq_cycles_medium", C.c_uint64), ("seq_cycles_generic", C.c_uint64),
                ("pairs", C.c_uint64), ("scan_rows", C.c_uint64), ("static_evals", C.c_uint64), ("prep_ms", C.c_double),
                ("seq_stops", C.c_uint64 * 5), ("seq_cons_wait", C.c_uint64), ("seq_cons_work", C.c_uint64),
                ("seq_rewalks", C.c_uint64), ("seq_prof", C.c_uint64 * 16),
                ("place_ms", C.c_double), ("place_tasks", C.c_uint64), ("place_cuts", C.c_uint64), ("place_amb", C.c_uint64),
                ("place_tails", C.c_uint64), ("place_chunks", C.c_uint64), ("place_cyc", C.c_uint64 * 3)]

    def as_dict(self) -> dict:
        return {n: (list(getattr(self, n)) if hasattr(getattr(self, n), "__len__") else getattr(self, n)) for n, _ in self._fields_}


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"placement engine error {code}: {msg}")
        self.code = code


ABI_SYMBOLS = [
    "abi_version", "create", "destroy", "last_error", "node_upsert", "node_remove", "set_node_count",
    "node_task_delta", "schedule", "tick_upload", "tick_run", "tick_download", "fit", "snapshot",
    "snapshot_service", "snapshot_generic", "snapshot_ports", "get_stats", "stats_reset", "fold_value",
    "nccl_unique_id", "pref_leaves", "match_matrix",
]


def _ptr(a):
    return None if a is None or a.size == 0 else a.ctypes.data


def _arr(a, dt):
    if a is None:
        return np.zeros(0, dt)
    a = np.ascontiguousarray(a, dtype=dt)
    return a


class Tick:
    """One scheduling pass in the flat ABI: groups + side arrays (numpy)."""

    def __init__(self, groups, task_flags, gens=None, cons=None, ips=None, plats=None, ports=None, plugs=None,
                 fails=None):
        self.groups = _arr(groups, group_dt)
        self.task_flags = _arr(task_flags, np.uint8)
        self.gens = _arr(gens, generic_want_dt)
        self.cons = _arr(cons, constraint_dt)
        self.ips = _arr(ips, ip_constraint_dt)
        self.plats = _arr(plats, platform_dt)
        self.ports = _arr(ports, np.uint32)
        self.plugs = _arr(plugs, np.uint32)
        self.fails = _arr(fails, node_fail_dt)

    @property
    def n_groups(self):
        return int(self.groups.size)

    @property
    def n_tasks(self):
        return int(self.task_flags.size)

    def c_struct(self) -> pe_tick:
        t = pe_tick()
        for name in ("groups", "task_flags", "gens", "cons", "ips", "plats", "ports", "plugs", "fails"):
            a = getattr(self, name)
            setattr(t, name, _ptr(a))
        t.n_groups, t.n_tasks = self.groups.size, self.task_flags.size
        t.n_gens, t.n_cons, t.n_ips = self.gens.size, self.cons.size, self.ips.size
        t.n_plats, t.n_ports, t.n_plugs, t.n_fails = self.plats.size, self.ports.size, self.plugs.size, self.fails.size
        return t

    def slice_groups(self, lo: int, hi: int) -> "Tick":
        """Sub-tick over groups [lo, hi); task offsets are rebased."""
        g = self.groups[lo:hi].copy()
        if g.size == 0:
            return Tick(g, np.zeros(0, np.uint8), self.gens, self.cons, self.ips, self.plats, self.ports, self.plugs,
                        self.fails)
        t0 = int(g["task_off"].min())
        t1 = int((g["task_off"] + g["n_tasks"]).max())
        g["task_off"] -= t0
        return Tick(g, self.task_flags[t0:t1], self.gens, self.cons, self.ips, self.plats, self.ports, self.plugs,
                    self.fails)


class NodeTable:
    """Node rows + side arrays for pe_node_upsert (numpy)."""

    def __init__(self, rows, attrs=None, gens=None, svcs=None, ports=None, plugs=None):
        self.rows = _arr(rows, node_row_dt)
        self.attrs = _arr(attrs, kv32_dt)
        self.gens = _arr(gens, kv64_dt)
        self.svcs = _arr(svcs, kv32_dt)
        self.ports = _arr(ports, np.uint32)
        self.plugs = _arr(plugs, np.uint32)


class FlatABI:
    """Binds one implementation of the placement ABI (symbol prefix + library)."""

    def __init__(self, lib_path: str, prefix: str, node_capacity: int = 0, device: int = -1, flags: int = 0,
                 max_batch: int = 0, rank: int = 0, world_size: int = 1, nccl_unique_id: bytes | None = None):
        self.lib = C.CDLL(lib_path)
        self.prefix = prefix
        self.f = {}
        for s in ABI_SYMBOLS:
            self.f[s] = getattr(self.lib, prefix + s)  # AttributeError if a symbol is missing
        self.f["abi_version"].restype = C.c_uint32
        self.f["last_error"].restype = C.c_char_p
        self.f["last_error"].argtypes = [C.c_void_p]
        self.f["destroy"].restype = None
        self.f["destroy"].argtypes = [C.c_void_p]
        for s in ABI_SYMBOLS:
            if s not in ("abi_version", "last_error", "destroy"):
                self.f[s].restype = C.c_int32
        self.f["create"].argtypes = [C.POINTER(pe_config), C.POINTER(C.c_void_p)]
        self.f["node_upsert"].argtypes = [C.c_void_p, C.c_void_p, C.c_uint32] + [C.c_void_p] * 5
        self.f["node_remove"].argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        self.f["set_node_count"].argtypes = [C.c_void_p, C.c_uint32]
        self.f["node_task_delta"].argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        self.f["schedule"].argtypes = [C.c_void_p, C.POINTER(pe_tick), C.c_void_p, C.c_void_p]
        self.f["tick_upload"].argtypes = [C.c_void_p, C.POINTER(pe_tick)]
        self.f["tick_run"].argtypes = [C.c_void_p]
        self.f["tick_download"].argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.f["fit"].argtypes = [C.c_void_p, C.POINTER(pe_tick), C.c_void_p, C.c_void_p, C.c_void_p]
        self.f["snapshot"].argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        for s in ("snapshot_service", "snapshot_generic", "snapshot_ports"):
            self.f[s].argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        self.f["get_stats"].argtypes = [C.c_void_p, C.POINTER(pe_stats)]
        self.f["stats_reset"].argtypes = [C.c_void_p]
        self.f["fold_value"].argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32]
        self.f["nccl_unique_id"].argtypes = [C.c_void_p]
        self.f["match_matrix"].argtypes = [C.c_void_p, C.POINTER(pe_tick), C.c_void_p]
        self.f["pref_leaves"].argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        if self.f["abi_version"]() != PE_ABI_VERSION:
            raise RuntimeError("ABI version mismatch")
        self._nccl_id = None
        if world_size > 1:
            if nccl_unique_id is None or len(nccl_unique_id) != 128:
                raise ValueError("world_size > 1 needs the 128-byte id from nccl_unique_id() of rank 0")
            self._nccl_id = C.create_string_buffer(bytes(nccl_unique_id), 128)
        cfg = pe_config(PE_ABI_VERSION, device, node_capacity, flags, max_batch, rank, world_size,
                        C.cast(self._nccl_id, C.c_void_p) if self._nccl_id is not None else None)
        self.h = C.c_void_p()
        rc = self.f["create"](C.byref(cfg), C.byref(self.h))
        if rc != PE_OK:
            msg = self.f["last_error"](None)
            self.h = None
            raise EngineError(rc, msg.decode() if msg else "create failed")

    # -- helpers
    def _check(self, rc: int):
        if rc != PE_OK:
            msg = self.f["last_error"](self.h)
            raise EngineError(rc, msg.decode() if msg else "?")

    def close(self):
        if getattr(self, "h", None):
            self.f["destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- node mirror
    def node_upsert(self, nt: NodeTable):
        self._check(self.f["node_upsert"](self.h, _ptr(nt.rows), nt.rows.size, _ptr(nt.attrs), _ptr(nt.gens),
                                          _ptr(nt.svcs), _ptr(nt.ports), _ptr(nt.plugs)))

    def node_remove(self, idx):
        idx = np.ascontiguousarray(idx, np.uint32)
        self._check(self.f["node_remove"](self.h, _ptr(idx), idx.size))

    def set_node_count(self, n: int):
        self._check(self.f["set_node_count"](self.h, n))

    def node_task_delta(self, deltas, gens=None, ports=None):
        d = _arr(deltas, task_delta_dt)
        g = _arr(gens, kv64_dt)
        p = _arr(ports, np.uint32)
        self._check(self.f["node_task_delta"](self.h, _ptr(d), d.size, _ptr(g), _ptr(p)))

    # -- hot path
    def schedule(self, tick: Tick, out_node=None, out_fail=None):
        """pe_schedule.  out_node / out_fail: optional caller-owned result buffers (uint32, at least n_tasks and
        n_groups * 8 long) -- e.g. page-locked arrays, so the copy back does not go through a staging buffer."""
        if out_node is None:
            out_node = np.full(max(tick.n_tasks, 1), PE_NONE, np.uint32)
        if out_fail is None:
            out_fail = np.zeros(max(tick.n_groups, 1) * PE_NUM_FILTERS, np.uint32)
        assert out_node.dtype == np.uint32 and out_fail.dtype == np.uint32
        assert out_node.size >= tick.n_tasks and out_fail.size >= tick.n_groups * PE_NUM_FILTERS
        ts = tick.c_struct()
        self._check(self.f["schedule"](self.h, C.byref(ts), _ptr(out_node), _ptr(out_fail)))
        return out_node[:tick.n_tasks], out_fail[:tick.n_groups * PE_NUM_FILTERS].reshape(-1, PE_NUM_FILTERS)

    def tick_upload(self, tick: Tick):
        ts = tick.c_struct()
        self._tick_shape = (tick.n_tasks, tick.n_groups)
        self._check(self.f["tick_upload"](self.h, C.byref(ts)))

    def tick_run(self):
        self._check(self.f["tick_run"](self.h))

    def tick_download(self):
        nt, ng = self._tick_shape
        out_node = np.full(max(nt, 1), PE_NONE, np.uint32)
        out_fail = np.zeros(max(ng, 1) * PE_NUM_FILTERS, np.uint32)
        self._check(self.f["tick_download"](self.h, _ptr(out_node), _ptr(out_fail)))
        return out_node[:nt], out_fail[:ng * PE_NUM_FILTERS].reshape(-1, PE_NUM_FILTERS)

    def fit(self, tick: Tick, node_idx):
        node_idx = np.ascontiguousarray(node_idx, np.uint32)
        ok = np.zeros(max(tick.n_groups, 1), np.uint8)
        out_fail = np.zeros(max(tick.n_groups, 1) * PE_NUM_FILTERS, np.uint32)
        ts = tick.c_struct()
        self._check(self.f["fit"](self.h, C.byref(ts), _ptr(node_idx), _ptr(ok), _ptr(out_fail)))
        return ok[:tick.n_groups], out_fail[:tick.n_groups * PE_NUM_FILTERS].reshape(-1, PE_NUM_FILTERS)

    def match_matrix(self, tick: Tick, n_nodes: int):
        """[n_groups, n_nodes] bool: the node-attribute filters each group enables, for every (group, node) pair."""
        words = (n_nodes + 31) // 32
        out = np.zeros(max(tick.n_groups * words, 1), np.uint32)
        ts = tick.c_struct()
        self._check(self.f["match_matrix"](self.h, C.byref(ts), _ptr(out)))
        bits = np.unpackbits(out[:tick.n_groups * words].view(np.uint8), bitorder="little").reshape(tick.n_groups, words * 32)
        return bits[:, :n_nodes].astype(bool)

    def pref_leaves(self, svc: int, cols, cap: int = 4096):
        """Leaves of the placement-preference tree of one service: (value tuples [n, levels], task sums [n])."""
        cols = np.ascontiguousarray(cols, np.uint32)
        vals = np.zeros(max(cap * max(cols.size, 1), 1), np.uint32)
        tasks = np.zeros(max(cap, 1), np.uint32)
        n = C.c_uint32(0)
        self._check(self.f["pref_leaves"](self.h, svc, _ptr(cols), cols.size, cap, _ptr(vals), _ptr(tasks), C.byref(n)))
        return vals[:n.value * cols.size].reshape(n.value, cols.size), tasks[:n.value]

    # -- introspection
    def snapshot(self, first: int, n: int):
        out = np.zeros(max(n, 1), node_state_dt)
        self._check(self.f["snapshot"](self.h, first, n, _ptr(out)))
        return out[:n]

    def snapshot_service(self, svc: int, first: int, n: int):
        out = np.zeros(max(n, 1), np.uint32)
        self._check(self.f["snapshot_service"](self.h, svc, first, n, _ptr(out)))
        return out[:n]

    def snapshot_generic(self, kind: int, first: int, n: int):
        out = np.zeros(max(n, 1), np.int64)
        self._check(self.f["snapshot_generic"](self.h, kind, first, n, _ptr(out)))
        return out[:n]

    def snapshot_ports(self, slot: int, first: int, n: int):
        out = np.zeros(max(n, 1), np.uint8)
        self._check(self.f["snapshot_ports"](self.h, slot, first, n, _ptr(out)))
        return out[:n]

    def stats(self) -> dict:
        s = pe_stats()
        self._check(self.f["get_stats"](self.h, C.byref(s)))
        return s.as_dict()

    def stats_reset(self):
        self._check(self.f["stats_reset"](self.h))

    def fold_value(self, s: bytes) -> bytes:
        buf = C.create_string_buffer(len(s) + 1)
        n = self.f["fold_value"](s, len(s), buf, len(s) + 1)
        if n < 0:
            raise ValueError("fold_value failed")
        return buf.raw[:n]
