"""Build helpers: compile the CUDA engine (sm_100a) and the host shim in-tree.

The shared objects are git-ignored but travel to the GPU box with the
snapshot, so nothing is JIT-compiled there.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "swarmkit_b200")
CSRC = os.path.join(PKG, "csrc")
LIB_ENGINE = os.path.join(PKG, "libplacement.so")
LIB_SCHED = os.path.join(PKG, "libswarmsched.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _sources(directory: str, exts=(".cu", ".cuh", ".h", ".hpp", ".cpp")) -> list[str]:
    out = []
    for base, _, files in os.walk(directory):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return out


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_engine(force: bool = False, verbose: bool = False) -> str:
    """nvcc -> swarmkit_b200/libplacement.so (CUDA kernels + C ABI)."""
    deps = _sources(CSRC) + [os.path.join(ROOT, "include", "placement_engine.h")]
    if not force and _newer(LIB_ENGINE, deps):
        return LIB_ENGINE
    cmd = [nvcc_path(), *NVCC_FLAGS, os.path.join(CSRC, "engine.cu"), "-o", LIB_ENGINE]
    if os.environ.get("PE_SEQ_PROFILE") == "1":   # diagnostic build: the sequencer's ordered warp tallies wait / work cycles
        cmd.insert(1, "-DPE_SEQ_PROFILE=1")
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed for engine.cu")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_ENGINE


def build_scheduler_shim(force: bool = False) -> str | None:
    """g++ -> swarmkit_b200/libswarmsched.so (host-side mirror of the Go scheduler)."""
    src = os.path.join(CSRC, "scheduler_host.cpp")
    if not os.path.exists(src):
        return None
    deps = _sources(CSRC) + [os.path.join(ROOT, "include", "placement_engine.h")]
    if not force and _newer(LIB_SCHED, deps):
        return LIB_SCHED
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", LIB_SCHED,
           "-L" + PKG, "-lplacement", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed for scheduler_host.cpp")
    return LIB_SCHED


def build_all(force: bool = False) -> None:
    build_engine(force=force)
    build_scheduler_shim(force=force)
