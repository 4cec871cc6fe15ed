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


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId for a node-sharded engine group (call on rank 0, ship to the other ranks)."""
    import ctypes as C
    lib = C.CDLL(engine_library_path())
    buf = C.create_string_buffer(128)
    lib.pe_nccl_unique_id.restype = C.c_int32
    rc = lib.pe_nccl_unique_id(buf)
    if rc != 0:
        lib.pe_last_error.restype = C.c_char_p
        raise RuntimeError(f"pe_nccl_unique_id failed ({rc}): {lib.pe_last_error(None).decode()}")
    return buf.raw


class PlacementEngine(FlatABI):
    """rank / world_size / nccl_id: node-sharded group -- every rank mirrors all nodes and submits the
    same ticks; each scans its slice of the node axis; placements come out identical on every rank."""

    def __init__(self, node_capacity: int = 0, device: int = -1, flags: int = 0, max_batch: int = 0,
                 rank: int = 0, world_size: int = 1, nccl_id: bytes | None = None):
        path = engine_library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). The placement engine has no CPU fallback.")
        super().__init__(path, "pe_", node_capacity=node_capacity, device=device, flags=flags, max_batch=max_batch,
                         rank=rank, world_size=world_size, nccl_unique_id=nccl_id)
