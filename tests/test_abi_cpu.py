"""CPU-side checks of the C-ABI boundary: the built CUDA library loads, exports
every symbol include/placement_engine.h declares, refuses to run without a
device (no CPU fallback), and the oracle exports the same ABI under ope_."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from swarmkit_b200 import _build, abi
from tests.oracle_lib import OracleEngine, build_flat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "placement_engine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pe_[a-z_]+)\s*\(", text)))


def test_header_declares_the_bound_symbols():
    assert declared_symbols() == sorted("pe_" + s for s in abi.ABI_SYMBOLS)


def test_cuda_library_exports_every_declared_symbol(engine_lib):
    lib = C.CDLL(engine_lib)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} missing from libplacement.so"
    lib.pe_abi_version.restype = C.c_uint32
    assert lib.pe_abi_version() == abi.PE_ABI_VERSION


def test_oracle_exports_the_same_abi(oracle_flat_lib):
    lib = C.CDLL(oracle_flat_lib)
    for s in declared_symbols():
        assert hasattr(lib, "o" + s), f"o{s} missing from the oracle"


def test_no_cpu_fallback(engine_lib):
    """Without a CUDA device pe_create must fail loudly (PE_ERR_NO_DEVICE)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from swarmkit_b200 import PlacementEngine
    with pytest.raises(abi.EngineError) as e:
        PlacementEngine(node_capacity=16)
    assert e.value.code == abi.PE_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_struct_layouts_match_the_header():
    # sizes the C compiler produces for the ABI structs (checked against the numpy mirrors)
    import subprocess, tempfile, textwrap
    src = textwrap.dedent("""
        #include <stdio.h>
        #include "placement_engine.h"
        int main(void) {
            printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pe_node_row), sizeof(pe_group), sizeof(pe_task_delta),
                   sizeof(pe_constraint), sizeof(pe_ip_constraint), sizeof(pe_platform), sizeof(pe_generic_want), sizeof(pe_node_fail),
                   sizeof(pe_tick), sizeof(pe_config), sizeof(pe_stats));
            return 0;
        }""")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "t")]).split()))
    expect = [abi.node_row_dt.itemsize, abi.group_dt.itemsize, abi.task_delta_dt.itemsize, abi.constraint_dt.itemsize,
              abi.ip_constraint_dt.itemsize, abi.platform_dt.itemsize, abi.generic_want_dt.itemsize, abi.node_fail_dt.itemsize,
              C.sizeof(abi.pe_tick), C.sizeof(abi.pe_config), C.sizeof(abi.pe_stats)]
    assert sizes == expect


@pytest.mark.parametrize("s,folded", [(b"Hello-World_1", b"hello-world_1"), ("K".encode(), b"k"), ("ſ".encode(), b"s"),
                                      ("Straße".encode(), "straße".encode()), (b"", b"")])
def test_fold_value(s, folded):
    assert OracleEngine().fold_value(s) == folded


def test_flat_oracle_cfg1_known_answer():
    # swarm-bench shape: 1000 identical replicas over 100 empty nodes -> 10 each
    from swarmkit_b200 import workloads as W
    w = W.cfg1()
    o = OracleEngine()
    o.node_upsert(w.nodes)
    o.set_node_count(w.n_nodes)
    out, _ = o.schedule(w.tick)
    assert (np.bincount(out, minlength=100) == 10).all()


def test_flat_oracle_fill_worked_example():
    """sorted leaf [a:0, b:0, c:5], k=7 -> a=1, b=6, c=5 (not a balanced water-fill)."""
    rows = np.zeros(3, abi.node_row_dt)
    rows["node_idx"] = [0, 1, 2]
    rows["flags"] = abi.PE_NODE_VALID | abi.PE_NODE_READY
    rows["total_tasks"] = [0, 0, 5]
    svcs = np.zeros(1, abi.kv32_dt)
    svcs[0] = (0, 5)
    rows["svc_off"], rows["svc_cnt"] = [0, 0, 0], [0, 0, 1]
    g = np.zeros(1, abi.group_dt)
    g["n_tasks"], g["filter_mask"], g["log_plugin"] = 7, 1, abi.PE_NONE
    o = OracleEngine()
    o.node_upsert(abi.NodeTable(rows, svcs=svcs))
    o.set_node_count(3)
    out, _ = o.schedule(abi.Tick(g, np.ones(7, np.uint8)))
    assert np.bincount(out, minlength=3).tolist() == [1, 6, 0]
    g["n_tasks"] = 4
    o2 = OracleEngine()
    rows["total_tasks"] = 0
    rows["svc_cnt"] = 0
    o2.node_upsert(abi.NodeTable(rows))
    o2.set_node_count(3)
    out, _ = o2.schedule(abi.Tick(g, np.ones(4, np.uint8)))
    assert np.bincount(out, minlength=3).tolist() == [2, 1, 1]
