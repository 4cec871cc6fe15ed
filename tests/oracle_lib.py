"""Test-side access to the CPU oracle (oracle/).  Nothing in swarmkit_b200/ imports this."""
from __future__ import annotations

import os
import subprocess

from swarmkit_b200.abi import FlatABI

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")


def _make(target: str) -> str:
    out = os.path.join(ORACLE, "_build", target)
    res = subprocess.run(["make", "-C", ORACLE, os.path.join("_build", target)], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return out


def build_flat() -> str:
    return _make("liboracle_flat.so")


def build_sched() -> str:
    return _make("liboracle_sched.so")


def build_shim_on_oracle() -> str:
    """The product's host shim compiled against the oracle ABI (test-only artifact)."""
    return _make("libshim_on_oracle.so")


class OracleEngine(FlatABI):
    """The flat CPU oracle behind the same ABI as the CUDA engine (prefix ope_)."""

    def __init__(self, node_capacity: int = 0):
        super().__init__(build_flat(), "ope_", node_capacity=node_capacity)


def build_volumes() -> str:
    return _make("liboracle_volumes.so")


class VolumesOracle:
    """JSON-op driver of oracle/volumes_oracle.cpp (volumeSet / IsInTopology / VolumesFilter restatement)."""

    def __init__(self):
        import ctypes as C
        import json
        self._json = json
        self.lib = C.CDLL(build_volumes())
        self.lib.vo_create.restype = C.c_void_p
        self.lib.vo_apply.restype = C.c_void_p
        self.lib.vo_apply.argtypes = [C.c_void_p, C.c_char_p]
        self.lib.vo_free.argtypes = [C.c_void_p]
        self.lib.vo_destroy.argtypes = [C.c_void_p]
        self.h = self.lib.vo_create()
        self._C = C

    def __call__(self, **ev):
        p = self.lib.vo_apply(self.h, self._json.dumps(ev).encode())
        try:
            out = self._json.loads(self._C.string_at(p).decode())
        finally:
            self.lib.vo_free(p)
        if "error" in out and out["error"] and ev.get("op") != "choose":
            raise RuntimeError(out["error"])
        return out

    def __del__(self):
        try:
            self.lib.vo_destroy(self.h)
        except Exception:
            pass
