"""swarmkit_b200 -- B200-native batched task->node placement engine for the
SwarmKit manager/scheduler hot path (filter pipeline + spread ranking +
reservation).  See DESIGN.md.  The compute path is hand-written sm_100a CUDA
behind the C ABI in include/placement_engine.h; there is no CPU fallback."""
from .abi import (EngineError, FlatABI, NodeTable, Tick, PE_NONE, PE_NUM_FILTERS)  # noqa: F401
from .engine import PlacementEngine, engine_library_path  # noqa: F401

__all__ = ["PlacementEngine", "engine_library_path", "EngineError", "NodeTable", "Tick", "PE_NONE", "PE_NUM_FILTERS"]
