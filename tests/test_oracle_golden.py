"""Pins the CPU oracle (oracle/mapf_oracle.c) to traces recorded from the live reference.

The fixtures under tests/golden/ were produced by tests/golden/gen_golden.py from the
unmodified reference classes (mapf_gridworld.MAPF_GRID, mapf_primal.MAPFEnv,
marl_partial.MARL_PARTIAL_ENV).  Everything is compared bit-for-bit.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import Oracle
from oracle.oracle import MODE_GRID, MODE_PARTIAL, MODE_PRIMAL


@pytest.mark.parametrize("name", golden_names("GRID"))
def test_grid_oracle_matches_reference_trace(name):
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    o = Oracle(1, N, H, W, MODE_GRID, episode_limit=int(g["episode_limit"]),
               step_reward=float(g["step_reward"]), collide_reward=float(g["collide_reward"]),
               sum_mode=int(g["py_sum_mode"]), step_is_int=int(g["step_is_int"]),
               collide_is_int=int(g["collide_is_int"]))
    o.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    assert np.array_equal(o.grid_state()[0], g["obs0"])
    assert np.array_equal(o.grid_avail()[0], g["avail0"])
    for t in range(g["actions"].shape[0]):
        out = o.grid_step(g["actions"][t][None])
        assert out["bad_actions"] == 0
        assert np.array_equal(o.positions()[0], g["pos"][t]), t
        assert np.array_equal(out["node"][0], g["node"][t]), t
        assert np.array_equal(out["edge"][0], g["edge"][t]), t
        assert np.array_equal(out["dones"][0], g["dones"][t]), t
        # bit-exact f64: compare the raw bit patterns
        assert out["reward"].view(np.uint64)[0] == g["reward"][t:t + 1].view(np.uint64)[0], (t, out["reward"], g["reward"][t])
        assert np.array_equal(o.grid_state()[0], g["state"][t]), t
        assert np.array_equal(out["avail"][0], g["avail"][t]), t
        assert o.step_count()[0] == g["step_count"][t]
        assert out["terminated"][0] == int(g["dones"][t].all())


@pytest.mark.parametrize("name", golden_names("PRIMAL") + golden_names("PRIMALB") + golden_names("PRIMALD"))
def test_primal_oracle_matches_reference_trace(name):
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    F = int(g["fov"])
    o = Oracle(1, N, H, W, MODE_PRIMAL, fov=F)
    o.set_blocking(bool(g["blocking_enabled"]))
    o.set_diagonal(bool(g["diagonal"]))
    o.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    obs, vec = o.primal_observe()
    assert np.array_equal(obs[0], g["obs0"])
    assert np.array_equal(vec[0].view(np.uint64), g["vec0"].view(np.uint64))
    assert np.array_equal(o.primal_avail()[0], g["avail0"])
    nc = g["costs0"].shape[0]
    costs = o.goal_dist(primal_costs=True)[0]
    assert np.array_equal(costs[:nc], g["costs0"])
    for t in range(g["actions"].shape[0]):
        out = o.primal_sweep(g["actions"][t][None])
        assert np.array_equal(out["status"][0], g["status"][t]), t
        assert np.array_equal(out["agent_reward"][0].view(np.uint64), g["reward"][t].view(np.uint64)), t
        assert np.array_equal(out["done_mid"][0], g["done_mid"][t]), t
        assert np.array_equal(out["next_mid"][0], g["next_mid"][t]), t
        assert np.array_equal(out["dones"][0], g["on_goal"][t]), t
        assert np.array_equal(out["valid"][0], g["valid"][t]), t
        if "blocking" in g:
            assert np.array_equal(out["blocking"][0], g["blocking"][t]), t
        assert np.array_equal(o.positions()[0], g["pos"][t]), t
        assert np.array_equal(out["avail"][0], g["avail"][t]), t
        assert out["terminated"][0] == g["done"][t]
        obs, vec = o.primal_observe()
        assert np.array_equal(obs[0], g["obs"][t]), t
        assert np.array_equal(vec[0].view(np.uint64), g["vec"][t].view(np.uint64)), t
    costs = o.goal_dist(primal_costs=True)[0]
    assert np.array_equal(costs[:nc], g["costsT"])


def test_primal_single_agent_steps_equal_sweep():
    """_step((id, a)) one agent at a time == one sweep (PRIMAL:549)."""
    g = load_golden("primal_crowd")
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    o = Oracle(1, N, H, W, MODE_PRIMAL, fov=int(g["fov"]))
    o.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    for t in range(10):
        for i in range(N):
            out = o.primal_sweep(g["actions"][t][None], lo=i, hi=i + 1)
            assert out["status"][0, i] == g["status"][t, i]
            assert np.array_equal(out["next_mid"][0, i], g["next_mid"][t, i])
        assert np.array_equal(o.positions()[0], g["pos"][t])


@pytest.mark.parametrize("name", golden_names("PDIST"))
def test_goal_dist_oracle_matches_partial_reference(name):
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["goals"].shape[0]
    o = Oracle(1, N, H, W, MODE_PRIMAL)
    o.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    d = o.goal_dist()[0].astype(np.int32)
    ref = g["dist"]
    assert np.array_equal(d == -1, g["obst"][None].repeat(N, 0).astype(bool))
    free = ~g["obst"].astype(bool)
    assert np.array_equal(d[:, free], ref[:, free])


def test_oracle_batch_is_independent_per_env():
    g = load_golden("primal_c2")
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    E = 5
    o = Oracle(E, N, H, W, MODE_PRIMAL, fov=int(g["fov"]), threads=3)
    o.reset(np.repeat(g["obst"][None], E, 0), np.repeat(g["starts"][None], E, 0), np.repeat(g["goals"][None], E, 0))
    for t in range(8):
        a = np.repeat(g["actions"][t][None], E, 0)
        o.primal_sweep(a)
        obs, vec = o.primal_observe()
        for e in range(E):
            assert np.array_equal(obs[e], g["obs"][t])
            assert np.array_equal(o.positions()[e], g["pos"][t])


def partial_kwargs(g):
    keys = ("obs_window", "obs_knn_agents", "move_reward", "stay_reward", "stay_goal_reward", "node_collide_reward",
            "edge_collide_reward", "env_collide_reward", "complete_reward", "complete_fac", "gamma")
    return {k: (int(g["cfg_" + k]) if g["cfgint_" + k] else float(g["cfg_" + k])) for k in keys}


@pytest.mark.parametrize("name", golden_names("PARTIAL"))
def test_partial_oracle_matches_reference_trace(name):
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    o = Oracle(1, N, H, W, MODE_PARTIAL, episode_limit=int(g["cfg_episode_limit"]), sum_mode=int(g["py_sum_mode"]))
    o.partial_config(**partial_kwargs(g))
    o.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    assert np.array_equal(o.partial_observe()[0].view(np.uint64), g["obs0"].view(np.uint64))
    assert np.array_equal(o.grid_avail()[0], g["avail0"])
    assert np.array_equal(o.partial_state()[0], g["state0"])
    for t in range(g["actions"].shape[0]):
        out = o.partial_step(g["actions"][t][None])
        assert np.array_equal(o.positions()[0], g["pos"][t]), t
        assert np.array_equal(out["node"][0], g["node"][t]), t
        assert np.array_equal(out["edge"][0], g["edge"][t]), t
        assert np.array_equal(out["at_goal"][0], g["at_goal"][t]), t
        assert np.array_equal(out["dones"][0], g["dones"][t]), t
        assert np.array_equal(out["goal_cost"][0], g["goal_cost"][t]), t
        assert np.array_equal(out["agent_steps"][0], g["agent_steps"][t]), t
        assert out["reward"].view(np.uint64)[0] == g["reward"][t:t + 1].view(np.uint64)[0], (t, out["reward"], g["reward"][t])
        assert out["terminated"][0] == g["terminated"][t]
        assert np.array_equal(out["avail"][0], g["avail"][t]), t
        assert np.array_equal(o.partial_state()[0], g["state"][t]), t
        assert np.array_equal(o.partial_observe()[0].view(np.uint64), g["obs"][t].view(np.uint64)), t
