"""Environment registry with the reference's convention (MARL-curve-main/src/envs/__init__.py:33-63):
REGISTRY[name] = partial(env_fn, env=Cls), constructed as REGISTRY[args.env](**args.env_args)."""
from functools import partial

from .mapf_gridworld import MAPF_GRID
from .mapf_primal import MAPFEnv
from .marl_partial import MARL_PARTIAL_ENV
from .multiagentenv import MultiAgentEnv


def env_fn(env, **kwargs) -> MultiAgentEnv:
    return env(**kwargs)


REGISTRY = {}
REGISTRY["mapf_gridworld"] = partial(env_fn, env=MAPF_GRID)
REGISTRY["mapf_primal"] = partial(env_fn, env=MAPFEnv)
REGISTRY["marl_partial"] = partial(env_fn, env=MARL_PARTIAL_ENV)   # the one the reference registers
