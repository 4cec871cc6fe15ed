"""PlacementEngine: the CUDA engine (libplacement.so) behind the flat C ABI.

This is the product binding.  It loads only swarmkit_b200/libplacement.so and
raises if that library is missing or if no CUDA device is present -- there is no
CPU fallback (pe_create returns PE_ERR_NO_DEVICE)."""
from __future__ import annotations

import os

from .abi import FlatABI

_HERE = os.path.dirname(os.path.abspath(__file__))


def engine_library_path() -> str:
    return os.path.join(_HERE, "libplacement.so")


class PlacementEngine(FlatABI):
    def __init__(self, node_capacity: int = 0, device: int = -1, flags: int = 0, max_batch: int = 0):
        path = engine_library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). The placement engine has no CPU fallback.")
        super().__init__(path, "pe_", node_capacity=node_capacity, device=device, flags=flags, max_batch=max_batch)
