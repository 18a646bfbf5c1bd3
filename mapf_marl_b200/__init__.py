"""mapf_marl_b200 -- B200-native batched MAPF grid-world step/observation engine.

Only the hot path of the reference's mapf_gridworld.py / mapf_primal.py lives here:
  csrc/            hand-written sm_100a kernels + the C ABI (include/mapf_b200.h)
  engine.py        MapfEngine: device-resident batch of environments (vector API)
  mapf_gridworld.py, mapf_primal.py, marl_partial.py, multiagentenv.py, registry.py
                   the reference's own env interfaces on top of the engine (drop-in)
  maps.py          MovingAI ingestion and synthetic worlds (reset-time, host side)
  sharding.py      environments sharded by index over the GPUs of a box
"""
from ._lib import build, load  # noqa: F401

__all__ = ["build", "load", "MapfEngine", "MAPF_GRID", "MAPFEnv", "MARL_PARTIAL_ENV", "REGISTRY"]


def __getattr__(name):
    # torch is imported lazily so that `build()` works in a minimal environment
    if name == "MapfEngine":
        from .engine import MapfEngine
        return MapfEngine
    if name == "MAPF_GRID":
        from .mapf_gridworld import MAPF_GRID
        return MAPF_GRID
    if name == "MAPFEnv":
        from .mapf_primal import MAPFEnv
        return MAPFEnv
    if name == "MARL_PARTIAL_ENV":
        from .marl_partial import MARL_PARTIAL_ENV
        return MARL_PARTIAL_ENV
    if name == "REGISTRY":
        from .registry import REGISTRY
        return REGISTRY
    raise AttributeError(name)
