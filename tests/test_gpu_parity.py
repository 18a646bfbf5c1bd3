"""GPU parity tests: the CUDA path (through the C ABI) against
  (a) traces recorded from the live reference (tests/golden/*.npz), and
  (b) the CPU oracle on seeded random batches, and
  (c) size-independent invariants at BASELINE.json's full sizes.
Integer / byte / index outputs and float64 rewards / goal vectors are all compared bit-for-bit.
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


def _engine(*a, **k):
    from mapf_marl_b200.engine import MapfEngine
    return MapfEngine(*a, **k)


def _oracle(*a, **k):
    from oracle import Oracle
    return Oracle(*a, **k)


def _bits(x):
    x = np.ascontiguousarray(x)
    return x.view(np.uint32 if x.dtype == np.float32 else np.uint64)


def _np(t):
    return t.cpu().numpy()


GRID_WANT = ("reward", "terminated", "agent_reward", "dones", "status", "node", "edge", "avail")
PRIMAL_WANT = ("reward", "terminated", "agent_reward", "dones", "status", "valid", "done_mid", "next_mid", "avail")


# ------------------------------------------------------------------------------------------ golden traces
@pytest.mark.parametrize("name", golden_names("GRID"))
def test_grid_engine_matches_reference_trace(name):
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    sr = int(g["step_reward"]) if g["step_is_int"] else float(g["step_reward"])
    cr = int(g["collide_reward"]) if g["collide_is_int"] else float(g["collide_reward"])
    eng = _engine(1, N, H, W, mode="grid", episode_limit=int(g["episode_limit"]), step_reward=sr,
                  collide_reward=cr, reward_sum_mode=int(g["py_sum_mode"]))
    eng.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    assert np.array_equal(_np(eng.observe()[0])[0], g["obs0"])
    assert np.array_equal(_np(eng.avail())[0], g["avail0"])
    for t in range(g["actions"].shape[0]):
        out = eng.step(torch.as_tensor(g["actions"][t][None]), want=GRID_WANT)
        assert np.array_equal(_np(eng.positions())[0], g["pos"][t]), t
        assert np.array_equal(_np(out["node"])[0], g["node"][t]), t
        assert np.array_equal(_np(out["edge"])[0], g["edge"][t]), t
        assert np.array_equal(_np(out["dones"])[0], g["dones"][t]), t
        assert _bits(_np(out["reward"]))[0] == _bits(g["reward"][t:t + 1])[0], (t, _np(out["reward"]), g["reward"][t])
        assert np.array_equal(_np(eng.observe()[0])[0], g["state"][t]), t
        assert np.array_equal(_np(out["avail"])[0], g["avail"][t]), t
        assert _np(eng.step_count())[0] == g["step_count"][t]
        assert _np(out["terminated"])[0] == int(g["dones"][t].all())
    # grid_real32 goes through the reference's own .scen parser, whose x/y transposition (GRID:443-445) puts
    # some agents on obstacle cells; the trace is reproduced all the same and the engine reports the fact.
    from mapf_marl_b200 import _lib
    # (grid_onwall*: agents deliberately start on walls, where `_full_obs` reads -1 + count, GRID:278, 299)
    on_wall = bool(g["obst"][g["starts"][:, 0], g["starts"][:, 1]].any())
    assert on_wall == (name in ("grid_real32", "grid_onwall4", "grid_onwall7"))
    assert eng.error_flags() == (_lib.FLAG_START_ON_WALL if on_wall else 0)


@pytest.mark.parametrize("name", golden_names("PRIMAL"))
def test_primal_engine_matches_reference_trace(name):
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    F = int(g["fov"])
    eng = _engine(1, N, H, W, mode="primal", fov=F)
    eng.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    obs, vec = eng.observe()
    assert np.array_equal(_np(obs)[0], g["obs0"])
    assert np.array_equal(_bits(_np(vec)[0]), _bits(g["vec0"]))
    assert np.array_equal(_np(eng.avail())[0], g["avail0"])
    nc = g["costs0"].shape[0]
    assert np.array_equal(_np(eng.goal_dist(primal_costs=True))[0, :nc], g["costs0"])
    for t in range(g["actions"].shape[0]):
        fused = (t % 2 == 0)
        a = torch.as_tensor(g["actions"][t][None])
        if fused:
            out = eng.step_observe(a, want=PRIMAL_WANT)
            obs, vec = out["obs"], out["vec"]
        else:
            out = eng.step(a, want=PRIMAL_WANT)
            obs, vec = eng.observe()
        assert np.array_equal(_np(out["status"])[0], g["status"][t]), t
        assert np.array_equal(_bits(_np(out["agent_reward"])[0]), _bits(g["reward"][t])), t
        assert np.array_equal(_np(out["done_mid"])[0], g["done_mid"][t]), t
        assert np.array_equal(_np(out["next_mid"])[0], g["next_mid"][t]), t
        assert np.array_equal(_np(out["dones"])[0], g["on_goal"][t]), t
        assert np.array_equal(_np(out["valid"])[0], g["valid"][t]), t
        assert np.array_equal(_np(eng.positions())[0], g["pos"][t]), t
        assert np.array_equal(_np(out["avail"])[0], g["avail"][t]), t
        assert _np(out["terminated"])[0] == g["done"][t]
        assert np.array_equal(_np(obs)[0], g["obs"][t]), t
        assert np.array_equal(_bits(_np(vec)[0]), _bits(g["vec"][t])), t
    assert np.array_equal(_np(eng.goal_dist(primal_costs=True))[0, :nc], g["costsT"])
    assert eng.error_flags() == 0


@pytest.mark.parametrize("name", golden_names("PDIST"))
def test_goal_dist_matches_partial_reference(name):
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["goals"].shape[0]
    eng = _engine(1, N, H, W, mode="primal", fov=5)
    eng.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    d = _np(eng.goal_dist())[0].astype(np.int32)
    free = ~g["obst"].astype(bool)
    assert np.array_equal(d[:, free], g["dist"][:, free])
    assert (d[:, ~free] == -1).all()


# ------------------------------------------------------------------------------------------ drop-in classes
def _write_movingai(tmp_path, obst, n_lines=40):
    H, W = obst.shape
    mp = tmp_path / "m.map"
    with open(mp, "w") as f:
        f.write("type octile\nheight %d\nwidth %d\nmap\n" % (H, W))
        for r in range(H):
            f.write("".join("@" if v else "." for v in obst[r]) + "\n")
    free = [(i, j) for i in range(H) for j in range(W) if not obst[i, j]]
    for k in range(1, 26):
        with open(tmp_path / ("s-%d.scen" % k), "w") as f:
            f.write("version 1\n")
            for n in range(n_lines):
                a, b = free[(n * 7 + k) % len(free)], free[(n * 13 + 3 * k + 1) % len(free)]
                f.write("0\tm.map\t%d\t%d\t%d\t%d\t%d\t%d\t1.0\n" % (W, H, a[0], a[1], b[0], b[1]))
    return str(mp), str(tmp_path / "s-")


@pytest.mark.parametrize("name", ["grid_kat_b1", "grid_c1", "grid_crowd6", "grid_intint"])
def test_mapf_grid_dropin_class(name, tmp_path):
    """The MAPF_GRID class with the reference's constructor, return types and attribute names."""
    from mapf_marl_b200.mapf_gridworld import MAPF_GRID
    g = load_golden(name)
    mp, sp = _write_movingai(tmp_path, g["obst"])
    N = g["starts"].shape[0]
    sr = int(g["step_reward"]) if g["step_is_int"] else float(g["step_reward"])
    cr = int(g["collide_reward"]) if g["collide_is_int"] else float(g["collide_reward"])
    env = MAPF_GRID(mp, sp, n_agents=N, episode_limit=int(g["episode_limit"]), seed=1, render="none",
                    step_reward=sr, collide_reward=cr)
    env.set_starts_goals(g["starts"], g["goals"])
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (N, g["obst"].size) and obs.dtype == np.int64
    assert np.array_equal(obs[0], g["obs0"]) and np.array_equal(obs[-1], g["obs0"])
    info = env.get_env_info()
    assert info == {"state_shape": g["obst"].size, "obs_shape": g["obst"].size, "n_actions": 5, "n_agents": N,
                    "episode_limit": int(g["episode_limit"])}
    for t in range(min(40, g["actions"].shape[0])):
        acts = torch.as_tensor(g["actions"][t].astype(np.int64)) if t % 2 else g["actions"][t]
        reward, dones, info = env.step(acts)
        assert isinstance(dones, list) and dones is env._agent_dones
        assert float(reward) == g["reward"][t] and isinstance(reward, int) == bool(g["reward_is_int"][t])
        assert [int(x) for x in dones] == g["dones"][t].tolist()
        assert info == {"_step_count": int(g["step_count"][t])}
        assert env.agent_positions == [tuple(p) for p in g["pos"][t].tolist()]
        assert env._node_collision_agents == g["node"][t].tolist()
        assert env._edge_collision_agents == g["edge"][t].tolist()
        assert env.get_avail_actions() == g["avail"][t].tolist()
        assert np.array_equal(env.get_state(), g["state"][t])
        assert np.array_equal(env.get_obs()[N - 1], g["state"][t])
        assert env.episode_done() == bool(g["dones"][t].all())
    with pytest.raises(AssertionError):
        env.step([0] * (N - 1) + [7])


@pytest.mark.parametrize("name", ["primal_kat_b2", "primal_crowd", "primal_f10"])
def test_mapfenv_dropin_class_single_agent_steps(name):
    """MAPFEnv._step((id, action)) one agent at a time, exactly as PRIMAL drives it."""
    from mapf_marl_b200.mapf_primal import MAPFEnv
    g = load_golden(name)
    N = g["starts"].shape[0]
    F = int(g["fov"])
    world0 = -g["obst"].astype(int)
    goals0 = np.zeros_like(world0)
    for k in range(N):
        world0[tuple(g["starts"][k])] = k + 1
        goals0[tuple(g["goals"][k])] = k + 1
    env = MAPFEnv(num_agents=N, observation_size=F, world0=world0, goals0=goals0)
    assert env.getPositions() == [tuple(p) for p in g["starts"].tolist()]
    assert env.getGoals() == [tuple(p) for p in g["goals"].tolist()]
    assert np.array_equal(env.getObstacleMap(), g["obst"].astype(int))
    for i in range(1, N + 1):
        assert env._listNextValidActions(i) == [a for a in range(5) if g["avail0"][i - 1, a]]
    for t in range(min(12, g["actions"].shape[0])):
        for i in range(1, N + 1):
            a = int(g["actions"][t, i - 1])
            state, reward, done, nxt, on_goal, blocking, valid = env._step((i, a))
            assert reward == g["reward"][t, i - 1]
            assert done == bool(g["done_mid"][t, i - 1])
            assert nxt == [k for k in range(5) if g["next_mid"][t, i - 1, k]]
            assert on_goal == bool(g["on_goal"][t, i - 1]) and valid == bool(g["valid"][t, i - 1])
            assert blocking is False
        assert env.getPositions() == [tuple(p) for p in g["pos"][t].tolist()]
        for i in range(1, N + 1):
            maps4, vec = env._observe(i)
            for c in range(4):
                assert maps4[c].dtype == np.float64 and np.array_equal(maps4[c], g["obs"][t, i - 1, c])
            assert np.array_equal(_bits(np.array(vec, dtype=np.float64)), _bits(g["vec"][t, i - 1]))
            assert env._listNextValidActions(i, int(g["actions"][t, i - 1])) == \
                [k for k in range(5) if g["avail"][t, i - 1, k]]
        assert env.world.done() == bool(g["done"][t])
    nc = g["costs0"].shape[0]
    if g["actions"].shape[0] <= 12:
        for k in range(nc):
            costs = env.getAstarCosts(env.world.getPos(k + 1), env.world.getGoal(k + 1))
            assert np.array_equal(costs, g["costsT"][k])


# ------------------------------------------------------------------------------------------ differential vs oracle
PRIMAL_CASES = [
    # E, N, H, W, F, density, shared, T
    (512, 8, 20, 20, 11, 0.2, False, 12),    # c2 shape
    (256, 32, 32, 32, 11, 0.3, False, 8),    # c3 shape
    (37, 7, 40, 40, 11, 0.25, False, 8),     # ragged: N % 4 != 0, E not a multiple of the tile, W > 32
    (64, 5, 9, 13, 3, 0.1, False, 10),       # tiny FOV, rectangular
    (64, 12, 8, 8, 5, 0.05, False, 20),      # crowded
    (48, 6, 12, 12, 10, 0.15, False, 8),     # even FOV
    (32, 128, 64, 64, 11, 0.0, True, 4),     # c4 shape: 128 agents, one shared warehouse map
    (16, 9, 16, 16, 6, 0.2, False, 6),       # FOV without a specialised kernel -> generic observation kernel
    (3, 1, 5, 5, 7, 0.0, False, 6),          # single agent
]


@pytest.mark.parametrize("case", PRIMAL_CASES, ids=lambda c: "E%d_N%d_%dx%d_F%d" % c[:5])
def test_primal_batch_matches_oracle(case):
    from mapf_marl_b200 import maps
    from oracle.oracle import MODE_PRIMAL
    E, N, H, W, F, dens, shared, T = case
    if shared:
        obst = maps.warehouse_layout(H, W)
        _, starts, goals = maps.synthetic_batch(7, E, H, W, 0.0, N, shared_map=True)
        rs = np.random.RandomState(5)
        free = np.argwhere(obst == 0)
        for e in range(E):
            idx = rs.permutation(len(free))
            starts[e] = free[idx[:N]]
            goals[e] = free[rs.permutation(len(free))[:N]]
    else:
        obst, starts, goals = maps.synthetic_batch(100 + E, E, H, W, dens, N, distinct=0)
    eng = _engine(E, N, H, W, mode="primal", fov=F, shared_map=shared)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=F, shared_map=shared)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    robs, rvec = orc.primal_observe()
    obs, vec = eng.observe()
    assert np.array_equal(_np(obs), robs)
    assert np.array_equal(_bits(_np(vec)), _bits(rvec))
    f32, _ = eng.observe(dtype=torch.float32)
    assert np.array_equal(_np(f32), robs.astype(np.float32))
    assert np.array_equal(_np(eng.avail()), orc.primal_avail())
    rs = np.random.RandomState(E + N)
    for t in range(T):
        a = rs.randint(0, 5, (E, N)).astype(np.uint8)
        ad = torch.as_tensor(a.astype(np.int64) if t % 3 == 1 else a, device="cuda")
        if t % 2 == 0:
            out = eng.step_observe(ad, want=PRIMAL_WANT)
            obs, vec = out["obs"], out["vec"]
        else:
            out = eng.step(ad, want=PRIMAL_WANT)
            obs, vec = eng.observe()
        ref = orc.primal_sweep(a)
        robs, rvec = orc.primal_observe()
        for k in ("status", "dones", "valid", "done_mid", "next_mid", "avail", "terminated"):
            assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        assert np.array_equal(_bits(_np(out["agent_reward"])), _bits(ref["agent_reward"])), t
        assert np.array_equal(_bits(_np(out["reward"])), _bits(ref["reward"])), t
        assert np.array_equal(_np(eng.positions()), orc.positions()), t
        assert np.array_equal(_np(obs), robs), t
        assert np.array_equal(_bits(_np(vec)), _bits(rvec)), t
    assert np.array_equal(_np(eng.goal_dist()), orc.goal_dist())
    assert np.array_equal(_np(eng.goal_dist(primal_costs=True)), orc.goal_dist(primal_costs=True))
    assert eng.error_flags() == 0
    st = eng.stats()
    assert st["env_steps"] == E * T and st["agent_steps"] == E * N * T


GRID_CASES = [
    (300, 4, 10, 10, 0.0, -0.01, -10, 25, 10000),
    (64, 12, 16, 16, 0.2, -0.013, -0.7, 30, 10000),
    (33, 10, 6, 6, 0.05, -0.01, -10, 40, 25),       # crowded, episode limit reached
    (16, 32, 32, 32, 0.2, -1, -10, 12, 10000),       # all-int rewards
    (8, 130, 40, 40, 0.1, -0.01, -0.5, 6, 10000),    # more than 128 agents
    # starts anywhere, walls included (negative density = |density| with on-wall starts): a wall cell holding an agent
    # reads -1 + count in `_full_obs`, is no obstacle (GRID:278) and may be entered (pinned by grid_onwall4/7)
    (96, 12, 7, 7, -0.35, -0.5, -3, 40, 10000),
    (20, 140, 24, 24, -0.3, -0.01, -10, 10, 10000),
]


@pytest.mark.parametrize("case", GRID_CASES, ids=lambda c: "E%d_N%d_%dx%d" % c[:4])
@pytest.mark.parametrize("sum_mode", [0, 1])
def test_grid_batch_matches_oracle(case, sum_mode):
    from oracle.oracle import MODE_GRID
    E, N, H, W, dens, sr, cr, T, limit = case
    rs = np.random.RandomState(E * 7 + N)
    on_wall = dens < 0
    obst = (rs.rand(E, H, W) < abs(dens)).astype(np.uint8)
    starts = np.zeros((E, N, 2), np.int16)
    goals = np.zeros((E, N, 2), np.int16)
    for e in range(E):
        free = np.argwhere(obst[e] >= 0) if on_wall else np.argwhere(obst[e] == 0)
        starts[e] = free[rs.randint(0, len(free), N)]      # overlapping starts are legal in GRID
        goals[e] = free[rs.randint(0, len(free), N)]
    eng = _engine(E, N, H, W, mode="grid", episode_limit=limit, step_reward=sr, collide_reward=cr,
                  reward_sum_mode=sum_mode)
    orc = _oracle(E, N, H, W, MODE_GRID, episode_limit=limit, step_reward=sr, collide_reward=cr, sum_mode=sum_mode)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    assert np.array_equal(_np(eng.observe()[0]), orc.grid_state())
    assert np.array_equal(_np(eng.avail()), orc.grid_avail())
    for t in range(T):
        a = rs.randint(0, 5, (E, N)).astype(np.uint8)
        if t % 2 == 0:
            out = eng.step_observe(torch.as_tensor(a, device="cuda"), want=GRID_WANT)
            state = out["obs"]
        else:
            out = eng.step(torch.as_tensor(a, device="cuda"), want=GRID_WANT)
            state = eng.observe()[0]
        ref = orc.grid_step(a)
        for k in ("terminated", "dones", "status", "node", "edge", "avail"):
            assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        assert np.array_equal(_bits(_np(out["agent_reward"])), _bits(ref["agent_reward"])), t
        assert np.array_equal(_bits(_np(out["reward"])), _bits(ref["reward"])), t
        assert np.array_equal(_np(eng.positions()), orc.positions()), t
        assert np.array_equal(_np(state), orc.grid_state()), t
        assert np.array_equal(_np(eng.step_count()), orc.step_count()), t
    from mapf_marl_b200 import _lib
    assert eng.error_flags() == (_lib.FLAG_START_ON_WALL if on_wall else 0)


# ------------------------------------------------------------------------------------------ API behaviour
def test_partial_sweeps_equal_full_sweep_and_masked_reset():
    from mapf_marl_b200 import maps
    E, N, H, W, F = 40, 8, 12, 12, 5
    obst, starts, goals = maps.synthetic_batch(3, E, H, W, 0.1, N, distinct=0)
    a = _engine(E, N, H, W, mode="primal", fov=F)
    b = _engine(E, N, H, W, mode="primal", fov=F)
    a.reset(obst, starts, goals)
    b.reset(obst, starts, goals)
    rs = np.random.RandomState(0)
    for t in range(6):
        act = torch.as_tensor(rs.randint(0, 5, (E, N)).astype(np.uint8), device="cuda")
        full = {k: v.clone() for k, v in a.step(act, want=PRIMAL_WANT).items()}
        for i in range(N):
            part = b.step(act, want=PRIMAL_WANT, agent_range=(i, i + 1))
            for k in ("status", "valid", "done_mid", "next_mid", "dones"):
                assert torch.equal(part[k][:, i], full[k][:, i]), (k, t, i)
        assert torch.equal(a.positions(), b.positions())
    # masked reset: only the selected environments go back to their start cells
    mask = np.zeros(E, np.uint8)
    mask[::3] = 1
    before = a.positions().clone()
    a.reset(env_mask=mask)
    after = _np(a.positions())
    assert np.array_equal(after[mask == 1], starts[mask == 1])
    assert np.array_equal(after[mask == 0], _np(before)[mask == 0])
    assert (_np(a.step_count())[mask == 1] == 0).all() and (_np(a.step_count())[mask == 0] == 6).all()


def test_bad_action_sets_device_flag_and_bad_args_are_rejected():
    from mapf_marl_b200 import _lib
    from mapf_marl_b200.engine import MapfError
    eng = _engine(4, 3, 6, 6, mode="primal", fov=5)
    eng.reset(np.zeros((4, 6, 6), np.uint8), np.tile(np.array([[0, 0], [1, 1], [2, 2]], np.int16), (4, 1, 1)),
              np.tile(np.array([[5, 5], [4, 4], [3, 3]], np.int16), (4, 1, 1)))
    assert eng.error_flags() == 0
    bad = torch.zeros((4, 3), dtype=torch.uint8, device="cuda")
    bad[2, 1] = 9
    eng.step(bad)
    assert eng.error_flags() == _lib.FLAG_BAD_ACTION
    assert eng.error_flags() == 0          # reading clears
    with pytest.raises(ValueError):
        eng.step(torch.zeros((4, 2), dtype=torch.uint8, device="cuda"))
    with pytest.raises(MapfError):
        eng.step(torch.zeros((4, 3), dtype=torch.uint8, device="cuda"), agent_range=(2, 1))
    with pytest.raises(MapfError):
        _engine(4, 3, 6, 6, mode="grid", obs_mode="fov")   # FOV needs one agent per cell
    with pytest.raises(MapfError):
        _engine(4, 300, 6, 6)
    # start on a wall / out of bounds / overlapping starts are flagged
    obst = np.zeros((4, 6, 6), np.uint8)
    obst[:, 0, 0] = 1
    eng.reset(obst, np.tile(np.array([[0, 0], [1, 1], [1, 1]], np.int16), (4, 1, 1)),
              np.tile(np.array([[5, 5], [4, 4], [9, 3]], np.int16), (4, 1, 1)))
    f = eng.error_flags()
    assert f & _lib.FLAG_START_ON_WALL and f & _lib.FLAG_BAD_POSITION and f & _lib.FLAG_START_OVERLAP


HOST_CASES = [
    # E, N, H, W, F, packed transport requested
    (96, 8, 20, 20, 11, True),
    (96, 8, 20, 20, 11, False),
    (37, 32, 32, 32, 11, True),     # last tile partial (37 envs, 4 per tile)
    (50, 7, 18, 18, 9, True),       # N not a multiple of the group size: tiles of 8 envs, last one short
    (33, 128, 64, 64, 11, True),    # one environment per tile
    (16, 9, 16, 16, 6, True),       # FOV without a specialised kernel: bits unsupported, dense copy
    (20, 5, 9, 13, 3, True),        # tiny FOV (G = 8, N = 5)
]


@pytest.mark.parametrize("case", HOST_CASES, ids=lambda c: "E%d_N%d_%dx%d_F%d_%s" % (c[:5] + ("packed" if c[5] else "dense",)))
def test_host_buffer_entry_point_matches_device_path(case):
    """mapf_step_observe_host: dense copies and the bit-packed PCIe transport (bits expanded by the library's host
    threads) deliver byte-identical host buffers, equal to the device path."""
    from mapf_marl_b200 import maps
    E, N, H, W, F, packed = case
    obst, starts, goals = maps.synthetic_batch(11, E, H, W, 0.2 if H < 64 else 0.1, N, distinct=0)
    a = _engine(E, N, H, W, mode="primal", fov=F)
    b = _engine(E, N, H, W, mode="primal", fov=F)
    a.reset(obst, starts, goals)
    b.reset(obst, starts, goals)
    mode = b.host_transport(packed)
    assert mode == int(packed and b.bits_supported())
    io, bufs, h2d, d2h = b.make_host_io()
    obs_bytes = b.packed_obs_bytes() if mode else E * N * 4 * F * F
    assert h2d == E * N and d2h == E * 8 + E + E * N + E * N * 5 + obs_bytes + E * N * 24
    bufs["obs"].fill_(7)
    rs = np.random.RandomState(1)
    for t in range(4):
        act = rs.randint(0, 5, (E, N)).astype(np.uint8)
        out = a.step_observe(torch.as_tensor(act, device="cuda"))
        bufs["actions"].copy_(torch.as_tensor(act))
        b.step_observe_host(io)
        for k in ("reward", "terminated", "dones", "avail", "obs", "vec"):
            assert torch.equal(out[k].cpu(), bufs[k]), (k, t)
    # float32 observations through the same transport
    io32, bufs32, _, d2h32 = b.make_host_io(obs_dtype=torch.float32)
    assert d2h32 == d2h - obs_bytes + (b.packed_obs_bytes() if mode else E * N * 4 * F * F * 4)
    act = rs.randint(0, 5, (E, N)).astype(np.uint8)
    out = a.step_observe(torch.as_tensor(act, device="cuda"), dtype=torch.float32)
    bufs32["actions"].copy_(torch.as_tensor(act))
    b.step_observe_host(io32)
    assert torch.equal(out["obs"].cpu(), bufs32["obs"])
    # an unaligned, unpinned destination buffer works as well
    raw = np.empty(E * N * 4 * F * F + 3, np.uint8)
    io.obs_host = raw[3:].ctypes.data
    act = rs.randint(0, 5, (E, N)).astype(np.uint8)
    out = a.step_observe(torch.as_tensor(act, device="cuda"))
    bufs["actions"].copy_(torch.as_tensor(act))
    b.step_observe_host(io)
    assert np.array_equal(raw[3:].reshape(E, N, 4, F, F), _np(out["obs"]))
    # the bit stream itself as the host output (no expansion): bit i == cell i
    if b.bits_supported():
        iob, bufsb, _, d2hb = b.make_host_io(obs_dtype="bits")
        assert d2hb == d2h - obs_bytes + b.packed_obs_bytes()
        bufsb["obs"].fill_(-1)
        act = rs.randint(0, 5, (E, N)).astype(np.uint8)
        out = a.step_observe(torch.as_tensor(act, device="cuda"))
        bufsb["actions"].copy_(torch.as_tensor(act))
        b.step_observe_host(iob)
        got = np.unpackbits(bufsb["obs"].numpy().view(np.uint8), bitorder="little")[:E * N * 4 * F * F]
        assert np.array_equal(got, _np(out["obs"]).reshape(-1))
        for k in ("reward", "terminated", "dones", "avail", "vec"):
            assert torch.equal(out[k].cpu(), bufsb[k]), k
    else:
        from mapf_marl_b200.engine import MapfError
        iob, bufsb, _, _ = b.make_host_io(obs_dtype="bits")
        with pytest.raises(MapfError):
            b.step_observe_host(iob)


@pytest.mark.parametrize("mode", ["grid", "partial"])
def test_host_buffer_entry_point_other_modes(mode):
    """mapf_step_observe_host for the GRID (int8 full map) and PARTIAL (float64 / float32 window + K-nearest block)
    observations: dense copies, equal to the device path."""
    from mapf_marl_b200 import maps
    E, N, H, W = 40, 6, 14, 14
    obst, starts, goals = maps.synthetic_batch(21, E, H, W, 0.15, N, distinct=0)
    kw = dict(episode_limit=50) if mode == "grid" else dict(episode_limit=50, obs_window=5, obs_knn_agents=4)
    a = _engine(E, N, H, W, mode=mode, **kw)
    b = _engine(E, N, H, W, mode=mode, **kw)
    a.reset(obst, starts, goals)
    b.reset(obst, starts, goals)
    rs = np.random.RandomState(3)
    for dt in ((torch.float64, torch.float32) if mode == "partial" else (torch.int8,)):
        io, bufs, h2d, d2h = b.make_host_io(obs_dtype=dt, want=("reward", "terminated", "dones", "avail", "obs"))
        assert bufs["obs"].dtype == dt
        for t in range(3):
            act = rs.randint(0, 5, (E, N)).astype(np.uint8)
            out = a.step_observe(torch.as_tensor(act, device="cuda"), dtype=dt)
            bufs["actions"].copy_(torch.as_tensor(act))
            b.step_observe_host(io)
            for k in ("reward", "terminated", "dones", "avail", "obs"):
                assert torch.equal(out[k].cpu(), bufs[k]), (k, t, dt)


@pytest.mark.parametrize("case", [(64, 8, 20, 20, 11), (37, 32, 32, 32, 11), (9, 128, 64, 64, 11), (40, 6, 12, 12, 10),
                                  (24, 16, 16, 16, 5)], ids=lambda c: "E%d_N%d_%dx%d_F%d" % c)
def test_bit_packed_observation_equals_the_uint8_tensor(case):
    """obs dtype MAPF_BITS: bit i of the stream == byte i of the uint8 observation."""
    from mapf_marl_b200 import maps
    E, N, H, W, F = case
    obst, starts, goals = maps.synthetic_batch(5, E, H, W, 0.15, N, distinct=0)
    eng = _engine(E, N, H, W, mode="primal", fov=F)
    eng.reset(obst, starts, goals)
    assert eng.bits_supported()
    rs = np.random.RandomState(2)
    n = E * N * 4 * F * F
    for t in range(3):
        dense, vec = eng.observe()
        dense = _np(dense).copy()
        bits, vec2 = eng.observe(dtype="bits")
        assert np.array_equal(np.unpackbits(_np(bits), bitorder="little")[:n], dense.reshape(-1)), t
        assert torch.equal(vec, vec2)
        act = torch.as_tensor(rs.randint(0, 5, (E, N)).astype(np.uint8), device="cuda")
        out = eng.step_observe(act, dtype="bits")
        dense2, _ = eng.observe()
        assert np.array_equal(np.unpackbits(_np(out["obs"]), bitorder="little")[:n], _np(dense2).reshape(-1)), t
    gen = _engine(16, 9, 16, 16, mode="primal", fov=6)     # no specialised kernel for F = 6
    assert not gen.bits_supported()
    gen.reset(*maps.synthetic_batch(5, 16, 16, 16, 0.1, 9, distinct=0))
    from mapf_marl_b200.engine import MapfError
    with pytest.raises(MapfError):
        gen.observe(dtype="bits")


@pytest.mark.parametrize("shape", [(32, 32), (64, 64), (40, 28), (20, 20), (30, 30), (64, 31), (31, 64)],
                         ids=lambda c: "%dx%d" % c)
def test_goal_dist_serpentine_maps_deeper_than_255_levels(shape):
    """The register BFS keeps distances in eight bit planes (255 levels); deeper maps (a serpentine corridor through
    the whole grid) overflow to the shared-memory kernel.  Mixed batch: serpentine, open and random maps, goals on
    both ends, with and without a dirty mask; every row width class of the store path (W % 8 == 0, W % 4 == 0, odd)."""
    from mapf_marl_b200 import maps
    from oracle.oracle import MODE_PRIMAL
    H, W = shape
    E, N = 6, 3
    obst = np.zeros((E, H, W), np.int8)
    for e in (0, 1, 2):                                   # wall rows with a gap at alternating ends
        for r in range(1, H, 2):
            obst[e, r, :] = 1
            obst[e, r, (W - 1) if (r // 2) % 2 == 0 else 0] = 0
    rs = np.random.RandomState(H * 100 + W)
    obst[4] = (rs.rand(H, W) < 0.25)
    obst[5] = (rs.rand(H, W) < 0.35)
    starts = np.zeros((E, N, 2), np.int16)
    goals = np.zeros((E, N, 2), np.int16)
    for e in range(E):
        free = np.argwhere(obst[e] == 0)
        idx = rs.permutation(len(free))
        starts[e] = free[idx[:N]]
        goals[e] = free[idx[N:2 * N]]
    last = H - 1 if (H - 1) % 2 == 0 else H - 2          # last corridor row
    goals[0, 0] = (0, 0)                                   # far end of the corridor: depth ~ H*W/2
    goals[1, 1] = (last, 0)
    goals[2, 2] = (last, W - 1)
    eng = _engine(E, N, H, W, mode="primal", fov=5)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=5)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    ref = orc.goal_dist()
    if H * W >= 1024:
        assert ref.max() > 255
    got = _np(eng.goal_dist())
    assert np.array_equal(got, ref)
    assert np.array_equal(_np(eng.goal_dist(primal_costs=True)), orc.goal_dist(primal_costs=True))
    dirty = np.zeros((E, N), np.uint8)
    dirty[0, 0] = dirty[2, 2] = dirty[3, 1] = dirty[5, 0] = 1
    out = torch.full((E, N, H, W), 77, dtype=torch.int16, device="cuda")
    eng.goal_dist(dirty=dirty, out=out)
    got = _np(out)
    assert np.array_equal(got[dirty != 0], ref[dirty != 0])
    assert (got[dirty == 0] == 77).all()


def test_set_goals_and_dirty_bfs():
    from mapf_marl_b200 import maps
    from oracle.oracle import MODE_PRIMAL
    E, N, H, W = 24, 6, 16, 16
    obst, starts, goals = maps.synthetic_batch(5, E, H, W, 0.2, N, distinct=0)
    eng = _engine(E, N, H, W, mode="primal", fov=5, goal_dist=True)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=5)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    dist = torch.full((E, N, H, W), -9, dtype=torch.int16, device="cuda")
    eng.goal_dist(out=dist)
    ref = orc.goal_dist()
    assert np.array_equal(_np(dist), ref)
    rs = np.random.RandomState(2)
    dirty = (rs.rand(E, N) < 0.3).astype(np.uint8)
    new_goals = goals.copy()
    for e in range(E):
        free = np.argwhere(obst[e] == 0)
        new_goals[e] = free[rs.permutation(len(free))[:N]]
    eng.set_goals(new_goals, dirty)
    orc.set_goals(new_goals, dirty)
    eng.goal_dist(dirty=dirty, out=dist)
    orc.goal_dist(dirty=dirty, out=ref)
    assert np.array_equal(_np(dist), ref)
    assert np.array_equal(_np(eng.goals()), np.where(dirty[..., None] != 0, new_goals, goals))
    obs, vec = eng.observe()
    robs, rvec = orc.primal_observe()
    assert np.array_equal(_np(obs), robs) and np.array_equal(_bits(_np(vec)), _bits(rvec))


# ------------------------------------------------------------------------------------------ full-size invariants
FULL_SIZE = [
    # (workload, envs, global env offset): the BASELINE.json configurations at their full per-GPU sizes
    ("c2", 4096, 0),
    ("c3", 16384, 0),
    ("c4", 8192, 0),                      # shared warehouse map, lifelong goal reassignment every step
    ("c5", 131072, 7 * 131072),           # the last GPU's shard of the 1M-env sweep on 8 GPUs
]


@pytest.mark.parametrize("cfg", FULL_SIZE, ids=lambda c: "%s_E%d_off%d" % c)
def test_full_size_every_env_every_step_matches_oracle(cfg):
    """Every environment of every BASELINE configuration, at full size, against the CPU oracle (all host threads)
    after EVERY one of 16 steps: status, per-agent and team rewards, dones, valid, terminated, action masks, positions,
    the 4-channel observation and the goal vectors, bit for bit.  Worlds and actions are functions of the GLOBAL
    environment index (>= 4096 distinct worlds; the c5 case is the shard a rank with env_offset 7 * 131072 owns);
    even steps draw uniform actions, odd steps draw from the action mask (agents keep moving and arrive).  c4 pops
    new goals from per-agent queues whenever agents arrive (LifelongGoals) and re-runs the BFS of those goals."""
    from mapf_marl_b200 import workloads
    from mapf_marl_b200.lifelong import LifelongGoals
    from oracle.oracle import MODE_PRIMAL
    name, E, lo = cfg
    wl = workloads.WORKLOADS[name]
    N, H, W, F = wl["N"], wl["H"], wl["W"], wl["F"]
    shared = wl["warehouse"]
    distinct = 0 if E <= 4096 else 4096
    obst, starts, goals = workloads.make_world(wl, E, lo, distinct=distinct)
    assert len({starts[e].tobytes() for e in range(min(E, 4096))}) >= min(E, 2048)       # genuinely different worlds
    eng = _engine(E, N, H, W, mode="primal", fov=F, shared_map=shared)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=F, shared_map=shared)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    want = ("status", "agent_reward", "reward", "dones", "valid", "terminated", "avail")
    life = None
    if shared:
        Q = 4
        queue = workloads.make_goal_queue(wl, obst, goals, E, lo, depth=Q, distinct=distinct)
        dist = torch.full((E, N, H, W), -9, dtype=torch.int16, device="cuda")
        eng.goal_dist(out=dist)
        life = LifelongGoals(eng, queue, dist_out=dist, overlap=True, fused=True)   # queues popped by the step kernel
        head = np.zeros((E, N), np.int64)
        cur_goals = goals.copy()
        ref_dist = orc.goal_dist()
        assert np.array_equal(_np(dist), ref_dist)
    avail = eng.avail()
    assert np.array_equal(_np(avail), orc.primal_avail())
    arrivals = 0
    for t in range(16):
        act = workloads.hash_actions_torch(99, lo, E, t, N, "cuda", avail=avail if t % 2 else None)
        a_np = _np(act)
        assert np.array_equal(a_np, workloads.hash_actions_np(99, range(lo, lo + E), t, N,
                                                              avail=_np(avail) if t % 2 else None))
        out = eng.step_observe(act, want=want)
        ref = orc.primal_sweep(a_np)
        for k in ("status", "dones", "valid", "terminated", "avail"):
            assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        assert np.array_equal(_bits(_np(out["agent_reward"])), _bits(ref["agent_reward"])), t
        assert np.array_equal(_bits(_np(out["reward"])), _bits(ref["reward"])), t
        assert np.array_equal(_np(eng.positions()), orc.positions()), t
        robs, rvec = orc.primal_observe()
        assert np.array_equal(_bits(_np(out["vec"])), _bits(rvec)), t
        # the observation is compared on the device, in slabs that bound the host-side staging
        slab = max(1, (1 << 28) // (N * 4 * F * F))
        for e0 in range(0, E, slab):
            assert torch.equal(out["obs"][e0:e0 + slab], torch.as_tensor(robs[e0:e0 + slab], device="cuda")), (t, e0)
        avail = out["avail"]
        arrivals += int(ref["dones"].sum())
        if life is not None:                                  # lifelong: arrived agents take the next queued goal
            life.reassign(out["dones"])
            life.sync()
            ref_dirty = ((ref["dones"] != 0) & (head < Q)).astype(np.uint8)
            new_goals = np.take_along_axis(queue, np.minimum(head, Q - 1)[..., None, None].repeat(2, -1), 2)[:, :, 0, :]
            head += ref_dirty
            orc.set_goals(new_goals, ref_dirty)
            orc.goal_dist(dirty=ref_dirty, out=ref_dist)
            if t % 8 == 7:
                assert np.array_equal(_np(dist), ref_dist), t
            cur_goals[ref_dirty != 0] = new_goals[ref_dirty != 0]
            assert np.array_equal(_np(eng.goals()), cur_goals), t
            assert np.array_equal(_np(life.head), head), t
    assert arrivals > 0 and eng.error_flags() == 0
    if life is not None:
        assert int(head.sum()) > 0


def test_sharded_engines_equal_the_unsharded_engine_bitwise():
    """SURVEY section 8e on the device: two engines of E/2 environments built from the global indices [0, E/2) and
    [E/2, E) produce, environment by environment, exactly what one engine of E environments produces."""
    from mapf_marl_b200 import workloads
    wl = workloads.WORKLOADS["c3"]
    E, N, F = 2048, wl["N"], wl["F"]
    mk = lambda n: _engine(n, N, wl["H"], wl["W"], mode="primal", fov=F)   # noqa: E731
    whole, parts = mk(E), [mk(E // 2), mk(E // 2)]
    whole.reset(*workloads.make_world(wl, E, 0, distinct=0))
    for r, p in enumerate(parts):
        p.reset(*workloads.make_world(wl, E // 2, r * E // 2, distinct=0))
    want = ("reward", "terminated", "agent_reward", "dones", "status", "avail")
    av = [None, None, None]
    for t in range(12):
        outs = []
        for k, (eng, lo, n) in enumerate([(whole, 0, E), (parts[0], 0, E // 2), (parts[1], E // 2, E // 2)]):
            act = eng.random_actions(5, t, avail=av[k] if t % 2 else None, env_offset=lo, dtype=torch.uint8)
            o = eng.step_observe(act, want=want, dtype="bits")
            av[k] = o["avail"]
            outs.append({kk: v.clone() for kk, v in o.items()})
            outs[-1]["pos"] = eng.positions().clone()
            outs[-1]["act"] = act.clone()
        for kk in outs[0]:
            joined = torch.cat([outs[1][kk], outs[2][kk]])
            assert torch.equal(outs[0][kk].view(torch.uint8), joined.view(torch.uint8)), (kk, t)
    cs = workloads.state_checksum_torch
    assert cs(outs[0]["pos"][:E // 2], outs[0]["avail"][:E // 2]) == cs(outs[1]["pos"], outs[1]["avail"])
    assert cs(outs[0]["pos"][:E // 2]) == workloads.state_checksum_np(_np(outs[1]["pos"]))


def test_random_actions_kernel_equals_the_numpy_counter_hash():
    from mapf_marl_b200 import workloads
    E, N = 300, 7
    eng = _engine(E, N, 12, 12, mode="primal", fov=5)
    rs = np.random.RandomState(0)
    avail = (rs.rand(E, N, 5) < 0.5).astype(np.uint8)
    avail[..., 0] |= (avail.sum(-1) == 0)
    for seed, step, off in ((0, 0, 0), (1234, 17, 5000), (2 ** 31 + 5, 70000, 7 * 131072)):
        a = eng.random_actions(seed, step, env_offset=off, dtype=torch.uint8)
        assert np.array_equal(_np(a), workloads.hash_actions_np(seed, range(off, off + E), step, N))
        b = eng.random_actions(seed, step, avail=torch.as_tensor(avail, device="cuda"), env_offset=off)
        assert b.dtype == torch.int64
        assert np.array_equal(_np(b), workloads.hash_actions_np(seed, range(off, off + E), step, N, avail=avail))


def test_primal_convoys_cycles_and_long_dependency_chains_match_oracle():
    """The ordered sweep is resolved in parallel rounds on the device (primal_classify): stress the cases where
    an agent's outcome depends on a chain of lower ids -- convoys in ascending / descending / shuffled id order,
    rotating 2x2 cycles, three-way races for one cell -- against the serial oracle."""
    from oracle.oracle import MODE_PRIMAL
    H, W, N = 6, 40, 36
    rs = np.random.RandomState(42)
    envs = []
    # rows 1 and 3 are corridors; agents stand shoulder to shoulder and all push the same way
    for order in ("asc", "desc", "shuffle", "shuffle2"):
        ids = np.arange(N)
        if order == "desc":
            ids = ids[::-1].copy()
        elif order.startswith("shuffle"):
            ids = rs.permutation(N)
        starts = np.zeros((N, 2), np.int16)
        for slot, a in enumerate(ids):       # slot 0 is the head of the convoy (nearest the free space ahead)
            starts[a] = (1 + 2 * (slot // 18), 30 - (slot % 18))
        envs.append(starts)
    # 2x2 rotating blocks and 3-way races, packed anywhere
    for _ in range(12):
        cells = [(r, c) for r in range(H) for c in range(W)]
        idx = rs.permutation(len(cells))[:N]
        envs.append(np.array([cells[i] for i in idx], np.int16))
    E = len(envs)
    starts = np.stack(envs)
    goals = np.zeros((E, N, 2), np.int16)
    for e in range(E):
        cells = [(r, c) for r in range(H) for c in range(W)]
        goals[e] = np.array([cells[i] for i in rs.permutation(len(cells))[:N]], np.int16)
    obst = np.zeros((E, H, W), np.uint8)
    eng = _engine(E, N, H, W, mode="primal", fov=5)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=5)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    for t in range(30):
        a = np.zeros((E, N), np.uint8)
        a[:4] = 1 if t % 6 < 4 else 3                    # convoys push east (PRIMAL action 1 = (0,+1)), then back west
        a[4:] = rs.randint(0, 5, (E - 4, N))
        if t % 5 == 4:
            a[4:] = rs.randint(1, 5)                      # everybody tries the same direction: long chains
        out = eng.step_observe(torch.as_tensor(a, device="cuda"), want=PRIMAL_WANT)
        ref = orc.primal_sweep(a)
        robs, _ = orc.primal_observe()
        for k in ("status", "dones", "valid", "done_mid", "next_mid", "avail", "terminated"):
            assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        assert np.array_equal(_np(eng.positions()), orc.positions()), t
        assert np.array_equal(_np(out["obs"]), robs), t
    # the ascending convoy moved as a whole in its first step, the descending one only at its head
    assert eng.error_flags() == 0


# ------------------------------------------------------------------------------------------ PARTIAL (marl_partial.py)
def _partial_kwargs(g):
    keys = ("obs_window", "obs_knn_agents", "move_reward", "stay_reward", "stay_goal_reward", "node_collide_reward",
            "edge_collide_reward", "env_collide_reward", "complete_reward", "complete_fac", "gamma")
    return {k: (int(g["cfg_" + k]) if g["cfgint_" + k] else float(g["cfg_" + k])) for k in keys}


PARTIAL_WANT = ("reward", "terminated", "agent_reward", "dones", "status", "node", "edge", "avail")


@pytest.mark.parametrize("name", golden_names("PARTIAL"))
def test_partial_engine_matches_reference_trace(name):
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    eng = _engine(1, N, H, W, mode="partial", episode_limit=int(g["cfg_episode_limit"]),
                  reward_sum_mode=int(g["py_sum_mode"]), **_partial_kwargs(g))
    eng.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    free = ~g["obst"].astype(bool)
    assert np.array_equal(_np(eng.goal_dist())[0][:, free].astype(np.int32), g["dist"][:, free])
    assert np.array_equal(_bits(_np(eng.observe()[0])[0]), _bits(g["obs0"]))
    assert np.array_equal(_np(eng.avail())[0], g["avail0"])
    assert np.array_equal(_np(eng.partial_state()["state"])[0], g["state0"])
    for t in range(g["actions"].shape[0]):
        out = eng.step(torch.as_tensor(g["actions"][t][None]), want=PARTIAL_WANT)
        ps = eng.partial_state()
        assert np.array_equal(_np(eng.positions())[0], g["pos"][t]), t
        assert np.array_equal(_np(out["node"])[0], g["node"][t]), t
        assert np.array_equal(_np(out["edge"])[0], g["edge"][t]), t
        assert np.array_equal(_np(ps["at_goal"])[0], g["at_goal"][t]), t
        assert np.array_equal(_np(out["dones"])[0], g["dones"][t]), t
        assert np.array_equal(_np(ps["goal_cost"])[0], g["goal_cost"][t]), t
        assert np.array_equal(_np(ps["agent_steps"])[0], g["agent_steps"][t]), t
        assert _bits(_np(out["reward"]))[0] == _bits(g["reward"][t:t + 1])[0], (t, _np(out["reward"]), g["reward"][t])
        assert _np(out["terminated"])[0] == g["terminated"][t]
        assert np.array_equal(_np(out["avail"])[0], g["avail"][t]), t
        assert np.array_equal(_np(ps["state"])[0], g["state"][t]), t
        assert np.array_equal(_bits(_np(eng.observe()[0])[0]), _bits(g["obs"][t])), t
    assert eng.error_flags() == 0


@pytest.mark.parametrize("case", [(96, 15, 8, 8, 0.0, 5, 5, 30), (40, 32, 32, 32, 0.2, 11, 8, 60),
                                  (25, 7, 12, 12, 0.1, 4, 9, 20), (8, 100, 24, 24, 0.05, 3, 4, 15),
                                  (6, 70, 20, 20, 0.05, 6, 40, 8)],
                         ids=lambda c: "E%d_N%d_%dx%d" % c[:4])
def test_partial_batch_matches_oracle(case):
    from oracle.oracle import MODE_PARTIAL
    E, N, H, W, dens, Wn, K, limit = case
    rs = np.random.RandomState(E + N)
    obst = np.zeros((E, H, W), np.uint8)
    starts = np.zeros((E, N, 2), np.int16)
    goals = np.zeros((E, N, 2), np.int16)
    from mapf_marl_b200 import maps
    for e in range(E):
        while True:
            m = (rs.rand(H, W) < dens).astype(np.uint8)
            lab = maps.label_components(m)
            big = np.argmax(np.bincount(lab[lab >= 0]))
            m[lab != big] = 1                       # one connected region: PARTIAL needs every goal reachable
            free = np.argwhere(m == 0)
            if len(free) >= 2 * N:
                break
        obst[e] = m
        starts[e] = free[rs.randint(0, len(free), N)]
        goals[e] = free[rs.randint(0, len(free), N)]
    kw = dict(obs_window=Wn, obs_knn_agents=K, move_reward=-0.01, stay_reward=-0.02, stay_goal_reward=0.3,
              node_collide_reward=-1, edge_collide_reward=-1.5, env_collide_reward=-1, complete_reward=100,
              complete_fac=1.5, gamma=0.97)
    eng = _engine(E, N, H, W, mode="partial", episode_limit=limit, **kw)
    orc = _oracle(E, N, H, W, MODE_PARTIAL, episode_limit=limit)
    orc.partial_config(**kw)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    assert np.array_equal(_bits(_np(eng.observe()[0])), _bits(orc.partial_observe()))
    for t in range(limit + 3):
        a = rs.randint(0, 5, (E, N)).astype(np.uint8)
        if t % 2 == 0:
            out = eng.step_observe(torch.as_tensor(a, device="cuda"), want=PARTIAL_WANT)
            obs = out["obs"]
        else:
            out = eng.step(torch.as_tensor(a, device="cuda"), want=PARTIAL_WANT)
            obs = eng.observe()[0]
        ref = orc.partial_step(a)
        ps = eng.partial_state()
        for k in ("terminated", "dones", "node", "edge", "avail"):
            assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        for k in ("at_goal", "goal_cost", "agent_steps"):
            assert np.array_equal(_np(ps[k]), ref[k]), (k, t)
        assert np.array_equal(_bits(_np(out["agent_reward"])), _bits(ref["agent_reward"])), t
        assert np.array_equal(_bits(_np(out["reward"])), _bits(ref["reward"])), t
        assert np.array_equal(_np(eng.positions()), orc.positions()), t
        assert np.array_equal(_np(ps["state"]), orc.partial_state()), t
        robs = orc.partial_observe()
        assert np.array_equal(_bits(_np(obs)), _bits(robs)), t
        if t % 5 == 0:       # float32 output == the float64 observation rounded once (what the episode batch stores)
            o32 = _np(eng.observe(dtype=torch.float32)[0])
            assert o32.dtype == np.float32 and np.array_equal(o32.view(np.uint32), robs.astype(np.float32).view(np.uint32)), t
    assert eng.error_flags() == 0


def test_partial_agents_on_wall_cells_follow_the_full_obs_count_rule():
    """marl_partial.py tests `_full_obs == -1` like mapf_gridworld.py does (:521, :339-342): a wall cell that holds an
    agent is no obstacle for moves, masks and the window maps, and shows count - 1 agents.  (The reference's own reset
    cannot produce such a state from a valid .scen; the oracle carries the rule, the engine must agree with it.)"""
    from oracle.oracle import MODE_PARTIAL
    E, N, H, W, Wn, K, limit = 48, 10, 9, 9, 5, 4, 30
    rs = np.random.RandomState(5)
    obst = (rs.rand(E, H, W) < 0.3).astype(np.uint8)
    cells = np.argwhere(np.ones((H, W), bool))
    starts = np.stack([cells[rs.randint(0, len(cells), N)] for _ in range(E)]).astype(np.int16)
    goals = np.stack([cells[rs.randint(0, len(cells), N)] for _ in range(E)]).astype(np.int16)
    kw = dict(obs_window=Wn, obs_knn_agents=K, move_reward=-0.01, stay_reward=-0.02, stay_goal_reward=0.3,
              node_collide_reward=-1, edge_collide_reward=-1.5, env_collide_reward=-1, complete_reward=100,
              complete_fac=1.5, gamma=0.97)
    eng = _engine(E, N, H, W, mode="partial", episode_limit=limit, **kw)
    orc = _oracle(E, N, H, W, MODE_PARTIAL, episode_limit=limit)
    orc.partial_config(**kw)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    assert np.array_equal(_bits(_np(eng.observe()[0])), _bits(orc.partial_observe()))
    assert np.array_equal(_np(eng.avail()), orc.grid_avail())
    for t in range(20):
        a = rs.randint(0, 5, (E, N)).astype(np.uint8)
        out = eng.step_observe(torch.as_tensor(a, device="cuda"), want=PARTIAL_WANT)
        ref = orc.partial_step(a)
        for k in ("terminated", "dones", "node", "edge", "avail"):
            assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        assert np.array_equal(_bits(_np(out["agent_reward"])), _bits(ref["agent_reward"])), t
        assert np.array_equal(_np(eng.positions()), orc.positions()), t
        assert np.array_equal(_bits(_np(out["obs"])), _bits(orc.partial_observe())), t


def test_marl_partial_dropin_class(tmp_path):
    """MARL_PARTIAL_ENV with the reference's constructor, reset/step/get_obs/get_state types."""
    from mapf_marl_b200.marl_partial import MARL_PARTIAL_ENV
    g = load_golden("partial_empty8_yaml")
    mp, sp = _write_movingai(tmp_path, g["obst"])
    N = g["starts"].shape[0]
    kw = _partial_kwargs(g)
    env = MARL_PARTIAL_ENV(mp, sp, n_agents=N, episode_limit=int(g["cfg_episode_limit"]), render="none", **kw)
    env.set_starts_goals(g["starts"], g["goals"])
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.dtype == np.float64 and obs.shape == (N, env.get_obs_size())
    assert np.array_equal(_bits(obs), _bits(g["obs0"]))
    assert env.get_env_info() == {"state_shape": 3, "obs_shape": obs.shape[1], "n_actions": 5, "n_agents": N,
                                  "episode_limit": int(g["cfg_episode_limit"])}
    for t in range(g["actions"].shape[0]):
        reward, terminated, info = env.step(torch.as_tensor(g["actions"][t].astype(np.int64)))
        assert isinstance(reward, float) and isinstance(terminated, bool)
        assert np.float64(reward).view(np.uint64) == g["reward"][t:t + 1].view(np.uint64)[0]
        assert terminated == bool(g["terminated"][t]) and info == {"_step_count": t + 1}
        assert [env.agent_pos(a) for a in range(N)] == [tuple(p) for p in g["pos"][t].tolist()]
        assert env.get_avail_actions() == g["avail"][t].tolist()
        assert np.array_equal(env.get_state(), g["state"][t])
        assert np.array_equal(_bits(env.get_obs()), _bits(g["obs"][t]))
        assert env.episode_done() == bool(g["dones"][t].all())
    # re-sampling reset: starts / goals come from the .scen files and are free cells
    env2 = MARL_PARTIAL_ENV(mp, sp, n_agents=N, render="none")
    random_obs = env2.reset()
    assert random_obs.shape == (N, env2.get_obs_size())
    from mapf_marl_b200.registry import REGISTRY
    env3 = REGISTRY["marl_partial"](grid_file_path=mp, agents_path=sp, n_agents=3, render="none")
    assert env3.reset().shape == (3, env3.get_obs_size())


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("overlap", [False, True])
def test_c4_shape_lifelong_goal_reassignment_matches_oracle(overlap, fused):
    """BASELINE config c4: 64x64 warehouse layout, 128 agents, goals popped from a per-agent queue on arrival
    (mapf_pop_goals), distance maps recomputed only for the reassigned goals (mapf_bfs with the dirty mask: compacted
    list + resident warps), optionally on a side stream overlapped with the next step.  fused: the queues are bound to
    the handle (mapf_lifelong_bind), the step kernel pops them in its write-back and mapf_bfs_popped takes the list the
    kernel collected -- same goals, heads, distance maps and observations, no pop launch."""
    from mapf_marl_b200 import maps
    from mapf_marl_b200.lifelong import LifelongGoals
    from oracle.oracle import MODE_PRIMAL
    E, N, H, W, F, Q = 12, 128, 64, 64, 11, 4
    rs = np.random.RandomState(3)
    obst = maps.warehouse_layout(H, W)
    free = np.argwhere(obst == 0)
    # PRIMAL's `goals` grid holds one id per cell, so goal cells must be distinct at all times: every agent draws its
    # start, first goal and queued goals from a private pool of cells
    assert len(free) >= N * (Q + 2)
    starts = np.zeros((E, N, 2), np.int16)
    goals = np.zeros((E, N, 2), np.int16)
    queue = np.zeros((E, N, Q, 2), np.int16)
    for e in range(E):
        pool = free[rs.permutation(len(free))[:N * (Q + 2)]].reshape(N, Q + 2, 2)
        starts[e] = pool[:, 0]
        goals[e] = pool[:, 1]
        near = rs.rand(N) < 0.5
        goals[e][near] = starts[e][near]             # half of the agents start on their goal: immediate reassignment
        queue[e] = pool[:, 2:]
    eng = _engine(E, N, H, W, mode="primal", fov=F, shared_map=True, goal_dist=True)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=F, shared_map=True)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    dist = torch.full((E, N, H, W), -9, dtype=torch.int16, device="cuda")
    eng.goal_dist(out=dist)
    life = LifelongGoals(eng, queue, dist_out=dist, overlap=overlap, fused=fused)
    ref_dist = orc.goal_dist()
    head = np.zeros((E, N), np.int64)
    cur_goals = goals.copy()
    for t in range(8):
        a = rs.randint(0, 5, (E, N)).astype(np.uint8)
        out = eng.step_observe(torch.as_tensor(a, device="cuda"), want=("dones", "status", "avail", "terminated"))
        ref = orc.primal_sweep(a)
        assert np.array_equal(_np(out["status"]), ref["status"]) and np.array_equal(_np(out["dones"]), ref["dones"])
        l0 = eng.launch_count()
        dirty = life.reassign(out["dones"])
        life.sync()
        ref_dirty = ((ref["dones"] != 0) & (head < Q)).astype(np.uint8)
        if fused:
            assert dirty is None and eng.launch_count() - l0 <= 2      # the BFS over the kernel's list (+ overflow pass)
        else:
            assert np.array_equal(_np(dirty), ref_dirty)
        new_goals = np.take_along_axis(queue, np.minimum(head, Q - 1)[..., None, None].repeat(2, -1), 2)[:, :, 0, :]
        head += ref_dirty
        orc.set_goals(new_goals, ref_dirty)
        orc.goal_dist(dirty=ref_dirty, out=ref_dist)
        assert np.array_equal(_np(dist), ref_dist), t
        cur_goals[ref_dirty != 0] = new_goals[ref_dirty != 0]
        assert np.array_equal(_np(eng.goals()), cur_goals), t
        assert np.array_equal(_np(life.head), head), t
        obs, vec = eng.observe()
        robs, rvec = orc.primal_observe()
        assert np.array_equal(_np(obs), robs) and np.array_equal(_bits(_np(vec)), _bits(rvec)), t
    assert int(head.sum()) > E * N // 4          # several hundred reassignments happened


def test_lifelong_bind_rollout_falls_back_to_one_launch_per_step_and_rejects_other_modes():
    """While goal queues are bound, a T-step mapf_rollout runs as T launches (every launch pops the queues once) and
    equals T single steps + mapf_pop_goals; binding is PRIMAL-only; unbinding restores the single launch."""
    from mapf_marl_b200 import maps
    E, N, H, W, F, Q, T = 32, 8, 20, 20, 11, 3, 6
    obst, starts, goals = maps.synthetic_batch(77, E, H, W, 0.2, N, distinct=0)
    goals = goals.copy()
    goals[:, ::2] = starts[:, ::2]                     # every other agent starts on its goal: pops from step 1 on
    rs = np.random.RandomState(5)
    free = [np.argwhere(obst[e] == 0) for e in range(E)]
    queue = np.stack([np.stack([free[e][rs.permutation(len(free[e]))[:Q]] for _ in range(N)]) for e in range(E)])
    queue = queue.astype(np.int16)
    acts = torch.as_tensor(rs.randint(0, 5, (T, E, N)).astype(np.uint8), device="cuda")
    a = _engine(E, N, H, W, mode="primal", fov=F)
    b = _engine(E, N, H, W, mode="primal", fov=F)
    qd = torch.as_tensor(queue, device="cuda").contiguous()
    head_a = torch.zeros((E, N), dtype=torch.int32, device="cuda")
    head_b = torch.zeros((E, N), dtype=torch.int32, device="cuda")
    a.reset(obst, starts, goals)
    b.reset(obst, starts, goals)
    assert b.rollout_in_one_launch()
    b.lifelong_bind(qd, head_b)
    assert not b.rollout_in_one_launch()
    l0 = b.launch_count()
    ob = b.rollout(acts, want=("reward", "dones", "avail"))
    assert b.launch_count() - l0 == T
    ob = {k: v.clone() for k, v in ob.items()}
    for t in range(T):
        oa = a.step_observe(acts[t], want=("reward", "dones", "avail"))
        for k in ("obs", "vec", "reward", "dones", "avail"):
            assert torch.equal(oa[k].view(torch.uint8), ob[k][t].view(torch.uint8)), (k, t)
        a.pop_goals(qd, head_a)
    assert torch.equal(head_a, head_b) and int(head_a.sum()) > E
    assert torch.equal(a.goals(), b.goals()) and torch.equal(a.positions(), b.positions())
    # (an agent moved onto its start may share that cell with somebody's goal: reset flags the overlap, nothing else)
    assert a.error_flags() == b.error_flags() and (a.error_flags() & ~16) == 0
    b.lifelong_bind(None, None)
    assert b.rollout_in_one_launch()
    g = _engine(4, 4, 10, 10, mode="grid")
    with pytest.raises(Exception):
        g.lifelong_bind(torch.zeros((4, 4, 2, 2), dtype=torch.int16, device="cuda"),
                        torch.zeros((4, 4), dtype=torch.int32, device="cuda"))


@pytest.mark.parametrize("obs_float32", [False, True])
@pytest.mark.parametrize("check_every", [1, 5])
def test_batched_runner_fills_an_episode_batch_like_the_parallel_runner(tmp_path, obs_float32, check_every):
    """The pymarl rollout loop over a device-resident vector env: what lands in the (time-major, kernel-written)
    batch equals a rollout of the CPU oracle driven by the same actions (PARTIAL env, the one the reference
    registers); entries are compared wherever `filled` says pymarl would have data."""
    from mapf_marl_b200.batched_runner import BatchedRunner, RandomMAC
    from mapf_marl_b200.marl_partial import MARL_PARTIAL_ENV
    from oracle.oracle import MODE_PARTIAL
    g = load_golden("partial_empty8_crowd")
    mp, sp = _write_movingai(tmp_path, g["obst"])
    B, N, limit = 24, 6, 12
    kw = dict(obs_window=5, obs_knn_agents=4, move_reward=-0.01, stay_reward=-0.02, stay_goal_reward=0,
              node_collide_reward=-1, edge_collide_reward=-1, env_collide_reward=-1, complete_reward=1000,
              complete_fac=1.5, gamma=0.99)
    env = MARL_PARTIAL_ENV(mp, sp, n_agents=N, episode_limit=limit, render="none", n_envs=B,
                           obs_float32=obs_float32, **kw)
    runner = BatchedRunner(env, RandomMAC(env.engine, seed=5), check_every=check_every)
    l0 = env.engine.launch_count()
    batch = runner.run(test_mode=False)
    T = runner.t                                             # env steps the loop executed
    _cache = {}
    bn = lambda k: _cache.setdefault(k, _np(batch[k]))       # noqa: E731
    # two launches per environment step (tile kernel + PARTIAL observation kernel carrying get_state)
    n_reset = 3 + 2                                          # reset: reset + BFS (+overflow) ; observe + avail
    # + the random policy: one kernel per step; + the runner's two bookkeeping kernels per step
    assert env.engine.launch_count() - l0 <= 2 * T + T + 2 * T + n_reset + 2
    odt = np.float32 if obs_float32 else np.float64
    assert bn("obs").dtype == odt
    H, W = g["obst"].shape
    orc = _oracle(B, N, H, W, MODE_PARTIAL, episode_limit=limit)
    orc.partial_config(**kw)
    orc.reset(np.repeat(g["obst"][None], B, 0), env._starts, env._goals)
    assert np.array_equal(bn("obs")[:, 0], orc.partial_observe().astype(odt))
    assert np.array_equal(bn("state")[:, 0], orc.partial_state())
    assert (bn("filled")[:, 0] == 1).all()
    alive = np.ones(B, bool)
    returns = np.zeros(B)
    for t in range(T):
        a = bn("actions")[:, t, :, 0]
        assert (a[~alive] == 4).all()                        # finished environments idle
        ref = orc.partial_step(a.astype(np.uint8))
        robs = orc.partial_observe()
        bs = np.nonzero(alive)[0]
        assert np.array_equal(bn("filled")[:, t + 1, 0].astype(bool), alive), t
        assert np.array_equal(_bits(bn("reward")[bs, t, 0]), _bits(ref["reward"][bs])), t
        assert np.array_equal(bn("obs")[bs, t + 1], robs[bs].astype(odt)), t
        assert np.array_equal(bn("avail_actions")[bs, t + 1], ref["avail"][bs]), t
        assert np.array_equal(bn("state")[bs, t + 1], orc.partial_state()[bs]), t
        assert np.array_equal(bn("terminated")[bs, t, 0], ref["terminated"][bs]), t
        # the actions the MAC chose were available
        assert (np.take_along_axis(bn("avail_actions")[bs, t], a[bs][..., None], -1) == 1).all(), t
        returns[bs] += ref["reward"][bs]
        alive &= ~ref["terminated"].astype(bool)
    assert not alive.any() or T == limit
    assert runner.t_env == int(bn("filled")[:, 1:, 0].sum())
    assert runner.train_stats["n_episodes"] == B and np.allclose(runner.train_returns, returns)
    # the vector env's own step(): one fused call, getters served from it
    env.reset()
    l1 = env.engine.launch_count()
    r, term, info = env.step(torch.full((B, N), 4, dtype=torch.int64, device="cuda"))
    obs, st, av = env.get_obs(), env.get_state(), env.get_avail_actions()
    assert env.engine.launch_count() - l1 == 2
    orc.reset(np.repeat(g["obst"][None], B, 0), env._starts, env._goals)
    ref = orc.partial_step(np.full((B, N), 4, np.uint8))
    assert np.array_equal(_np(obs), orc.partial_observe().astype(odt)) and np.array_equal(_np(st), orc.partial_state())
    assert np.array_equal(_np(av), ref["avail"]) and np.array_equal(_bits(_np(r)), _bits(ref["reward"]))


def test_batched_runner_primal_vec_env_one_launch_per_step_and_grid_env(tmp_path):
    """PrimalVecEnv (BASELINE config 5's env) through the runner: one engine launch per environment step, the batch
    equals an oracle rollout under the same actions; MAPF_GRID's vector step is one launch too."""
    from mapf_marl_b200 import maps
    from mapf_marl_b200.batched_runner import BatchedRunner, RandomMAC, RNNAgentMAC
    from mapf_marl_b200.vec_env import PrimalVecEnv
    from oracle.oracle import MODE_PRIMAL
    E, N, H, W, F, T = 48, 8, 20, 20, 11, 10
    obst, starts, goals = maps.synthetic_batch(21, E, H, W, 0.2, N, distinct=0)
    env = PrimalVecEnv(obst, starts, goals, fov=F, episode_limit=T)
    for mac, graph, fusedbk in ((RandomMAC(env.engine, seed=3), False, True),
                                (RNNAgentMAC(4 * F * F, 5, "cuda", extra_dim=3), False, True),
                                (RandomMAC(env.engine, seed=4), True, True),
                                (RandomMAC(env.engine, seed=5), False, False)):   # bookkeeping in torch
        runner = BatchedRunner(env, mac, check_every=4, cuda_graph=graph, fused_bookkeeping=fusedbk)
        n_mac = T if isinstance(mac, RandomMAC) else 0          # the random policy is one engine kernel per step
        if graph:
            runner.run()                                         # eager warm-up episode; the next run captures + replays
            mac.episode = -1                                     # same draws as the warm-up: (seed, episode 0, step)
        l0 = env.engine.launch_count()
        n_reset = 3 if env._maps_loaded else 4                   # the obstacle rows are built by the first reset only
        batch = runner.run()
        assert runner.t == T
        if graph:
            assert runner._graphs is not None and len(runner._graphs) == 3
        _cache = {}
        bn = lambda k, _c=_cache, _b=batch: _c.setdefault(k, _np(_b[k]))   # noqa: E731
        # reset (2 kernels the first time, then 1) + observe + avail at t = 0, then ONE launch per environment step (+ the random policy's, +
        # the runner's two bookkeeping kernels: mapf_runner_mask_actions, mapf_runner_account)
        if not graph:
            assert env.engine.launch_count() - l0 == n_reset + T + n_mac + (2 * T if fusedbk else 0)
        orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=F)
        orc.reset(obst, starts, goals)
        robs, rvec = orc.primal_observe()
        assert np.array_equal(bn("obs")[:, 0].reshape(robs.shape), robs)
        alive = np.ones(E, bool)
        for t in range(T):
            a = bn("actions")[:, t, :, 0].astype(np.uint8)
            assert (np.take_along_axis(bn("avail_actions")[:, t], a[..., None].astype(np.int64), -1) == 1)[alive].all()
            ref = orc.primal_sweep(a)
            robs, rvec = orc.primal_observe()
            bs = np.nonzero(alive)[0]
            assert np.array_equal(bn("obs")[bs, t + 1].reshape((len(bs),) + robs.shape[1:]), robs[bs]), t
            assert np.array_equal(_bits(bn("obs_vec")[bs, t + 1]), _bits(rvec[bs])), t
            assert np.array_equal(_bits(bn("state")[bs, t + 1]), _bits(rvec[bs].reshape(len(bs), -1))), t
            assert np.array_equal(bn("avail_actions")[bs, t + 1], ref["avail"][bs]), t
            assert np.array_equal(_bits(bn("reward")[bs, t, 0]), _bits(ref["reward"][bs])), t
            term = ref["terminated"].astype(bool) | (t + 1 >= T)
            assert np.array_equal(bn("terminated")[bs, t, 0].astype(bool), term[bs]), t
            alive &= ~term
    env.close()
    # MAPF_GRID as a vector env: step() is one launch and serves get_obs / get_state / get_avail_actions; the runner
    # stores the shared full-map observation once (obs is a stride-0 view over the agents)
    from mapf_marl_b200.mapf_gridworld import MAPF_GRID
    from oracle.oracle import MODE_GRID
    gobst = (np.random.RandomState(1).rand(10, 10) < 0.1).astype(np.uint8)
    mp, sp = _write_movingai(tmp_path, gobst)
    genv = MAPF_GRID(mp, sp, n_agents=4, episode_limit=9, render="none", n_envs=16)
    genv.reset()
    l1 = genv.engine.launch_count()
    acts = torch.as_tensor(np.random.RandomState(2).randint(0, 5, (16, 4)), device="cuda")
    r, term, info = genv.step(acts)
    st, av, ob = genv.get_state(), genv.get_avail_actions(), genv.get_obs()
    assert genv.engine.launch_count() - l1 == 1
    orc = _oracle(16, 4, 10, 10, MODE_GRID, episode_limit=9)
    orc.reset(np.repeat(gobst[None], 16, 0), genv._starts, genv._goals)
    ref = orc.grid_step(_np(acts).astype(np.uint8))
    assert np.array_equal(_np(st), orc.grid_state()) and np.array_equal(_np(av), ref["avail"])
    assert tuple(ob.shape) == (16, 4, 100) and np.array_equal(_np(ob[:, 2]), orc.grid_state())
    assert np.array_equal(_bits(_np(r)), _bits(ref["reward"]))
    runner = BatchedRunner(genv, RandomMAC(genv.engine, seed=1), check_every=3)
    batch = runner.run()
    assert tuple(batch["obs"].shape) == (16, 10, 4, 100) and tuple(batch["state"].shape) == (16, 10, 100)
    assert runner.t == 9 and bool((batch["terminated"][:, 8] == 1).all())     # the episode limit ends every env


@pytest.mark.parametrize("case", [(3, 255, 40, 40, 11, 0.05), (3, 7, 200, 200, 11, 0.2), (2, 5, 255, 255, 9, 0.1),
                                  (5, 3, 150, 90, 7, 0.3)], ids=lambda c: "E%d_N%d_%dx%d" % c[:4])
def test_extreme_sizes_match_oracle(case):
    """The limits of the ABI (255 agents, 255 x 255 maps) and maps so large that a tile cannot hold the aligned
    number of environments (the kernel's unaligned observation path)."""
    from mapf_marl_b200 import maps
    from oracle.oracle import MODE_PRIMAL
    E, N, H, W, F, dens = case
    obst, starts, goals = maps.synthetic_batch(900 + N, E, H, W, dens, N, distinct=0)
    eng = _engine(E, N, H, W, mode="primal", fov=F)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=F)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    rs = np.random.RandomState(N)
    for t in range(5):
        a = rs.randint(0, 5, (E, N)).astype(np.uint8)
        out = eng.step_observe(torch.as_tensor(a, device="cuda"), want=PRIMAL_WANT)
        ref = orc.primal_sweep(a)
        robs, rvec = orc.primal_observe()
        for k in ("status", "dones", "valid", "done_mid", "next_mid", "avail", "terminated"):
            assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        assert np.array_equal(_np(eng.positions()), orc.positions()), t
        assert np.array_equal(_np(out["obs"]), robs), t
        assert np.array_equal(_bits(_np(out["vec"])), _bits(rvec)), t
        f32, _ = eng.observe(dtype=torch.float32)
        assert np.array_equal(_np(f32), robs.astype(np.float32)), t
    assert np.array_equal(_np(eng.goal_dist()), orc.goal_dist())
    assert eng.error_flags() == 0


# ------------------------------------------------------------------------------------------ PRIMAL blocking reward
@pytest.mark.parametrize("name", golden_names("PRIMALB"))
def test_primal_blocking_reward_matches_reference_trace(name):
    """get_blocking_reward (mapf_primal.py:513-546) through the live reference with od_mstar3 replaced by a
    single-robot BFS (tests/golden/refload.py): rewards and `blocking` flags of every _step call."""
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    F = int(g["fov"])
    eng = _engine(1, N, H, W, mode="primal", fov=F, blocking_reward=True)
    eng.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    want = PRIMAL_WANT + ("blocking",)
    for t in range(g["actions"].shape[0]):
        out = eng.step_observe(torch.as_tensor(g["actions"][t][None]), want=want)
        assert np.array_equal(_np(out["status"])[0], g["status"][t]), t
        assert np.array_equal(_np(out["blocking"])[0], g["blocking"][t]), t
        assert np.array_equal(_bits(_np(out["agent_reward"])[0]), _bits(g["reward"][t])), t
        assert np.array_equal(_np(eng.positions())[0], g["pos"][t]), t
        assert np.array_equal(_np(out["obs"])[0], g["obs"][t]), t
    assert eng.error_flags() == 0
    # the per-agent facade returns the same blocking flags and rewards
    from mapf_marl_b200.mapf_primal import MAPFEnv
    world0 = -g["obst"].astype(int)
    goals0 = np.zeros_like(world0)
    for k in range(N):
        world0[tuple(g["starts"][k])] = k + 1
        goals0[tuple(g["goals"][k])] = k + 1
    env = MAPFEnv(num_agents=N, observation_size=F, world0=world0, goals0=goals0, blocking_reward=True)
    for t in range(min(12, g["actions"].shape[0])):
        for i in range(1, N + 1):
            _, reward, _, _, _, blocking, _ = env._step((i, int(g["actions"][t, i - 1])))
            assert reward == g["reward"][t, i - 1] and blocking == bool(g["blocking"][t, i - 1])


@pytest.mark.parametrize("case", [(48, 9, 10, 10, 11), (16, 20, 24, 40, 7), (8, 40, 64, 64, 11), (40, 14, 40, 20, 9)],
                         ids=lambda c: "E%d_N%d_%dx%d" % c[:4])
def test_primal_blocking_reward_batch_matches_oracle(case):
    from mapf_marl_b200 import maps
    from oracle.oracle import MODE_PRIMAL
    E, N, H, W, F = case
    rs = np.random.RandomState(E + H)
    obst = np.zeros((E, H, W), np.uint8)
    starts = np.zeros((E, N, 2), np.int16)
    goals = np.zeros((E, N, 2), np.int16)
    for e in range(E):
        m = (rs.rand(H, W) < 0.3).astype(np.uint8)          # dense: corridors, where blocking happens
        lab = maps.label_components(m)
        m[lab != np.argmax(np.bincount(lab[lab >= 0]))] = 1
        free = np.argwhere(m == 0)
        obst[e] = m
        starts[e] = free[rs.permutation(len(free))[:N]]
        goals[e] = free[rs.permutation(len(free))[:N]]
        near = rs.rand(N) < 0.6
        goals[e][near] = starts[e][near]                     # many agents already parked on their goals
        # goals must stay distinct cells
        seen = set()
        for k in range(N):
            while tuple(goals[e, k]) in seen:
                goals[e, k] = free[rs.randint(len(free))]
            seen.add(tuple(goals[e, k]))
    eng = _engine(E, N, H, W, mode="primal", fov=F, blocking_reward=True)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=F)
    orc.set_blocking(True)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    want = PRIMAL_WANT + ("blocking",)
    n_events = 0
    for t in range(6):
        a = rs.randint(0, 5, (E, N)).astype(np.uint8)
        a[rs.rand(E, N) < 0.5] = 0                            # lots of "stay"
        out = eng.step(torch.as_tensor(a, device="cuda"), want=want)
        ref = orc.primal_sweep(a)
        for k in ("status", "dones", "valid", "blocking", "avail", "terminated"):
            assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        assert np.array_equal(_bits(_np(out["agent_reward"])), _bits(ref["agent_reward"])), t
        assert np.array_equal(_bits(_np(out["reward"])), _bits(ref["reward"])), t
        n_events += int(ref["blocking"].sum())
    assert n_events > 0


# ------------------------------------------------------------------------------------------ PRIMAL DIAGONAL_MOVEMENT
@pytest.mark.parametrize("name", golden_names("PRIMALD"))
def test_primal_diagonal_engine_matches_reference_trace(name):
    """DIAGONAL_MOVEMENT=True (PRIMAL:175): 9 actions, crossing test, 9-wide masks, 8-connected costs."""
    g = load_golden(name)
    H, W = g["obst"].shape
    N = g["starts"].shape[0]
    F = int(g["fov"])
    eng = _engine(1, N, H, W, mode="primal", fov=F, diagonal_movement=True)
    eng.reset(g["obst"][None], g["starts"][None], g["goals"][None])
    obs, vec = eng.observe()
    assert np.array_equal(_np(obs)[0], g["obs0"])
    assert np.array_equal(_np(eng.avail())[0], g["avail0"]) and g["avail0"].shape[-1] == 9
    nc = g["costs0"].shape[0]
    assert np.array_equal(_np(eng.goal_dist(primal_costs=True))[0, :nc], g["costs0"])
    for t in range(g["actions"].shape[0]):
        a = torch.as_tensor(g["actions"][t][None])
        if t % 2 == 0:
            out = eng.step_observe(a, want=PRIMAL_WANT)
            obs, vec = out["obs"], out["vec"]
        else:
            out = eng.step(a, want=PRIMAL_WANT)
            obs, vec = eng.observe()
        assert np.array_equal(_np(out["status"])[0], g["status"][t]), t
        assert np.array_equal(_bits(_np(out["agent_reward"])[0]), _bits(g["reward"][t])), t
        assert np.array_equal(_np(out["done_mid"])[0], g["done_mid"][t]), t
        assert np.array_equal(_np(out["next_mid"])[0], g["next_mid"][t]), t
        assert np.array_equal(_np(out["dones"])[0], g["on_goal"][t]), t
        assert np.array_equal(_np(out["valid"])[0], g["valid"][t]), t
        assert np.array_equal(_np(eng.positions())[0], g["pos"][t]), t
        assert np.array_equal(_np(out["avail"])[0], g["avail"][t]), t
        assert _np(out["terminated"])[0] == g["done"][t]
        assert np.array_equal(_np(obs)[0], g["obs"][t]), t
        assert np.array_equal(_bits(_np(vec)[0]), _bits(g["vec"][t])), t
    assert np.array_equal(_np(eng.goal_dist(primal_costs=True))[0, :nc], g["costsT"])
    assert eng.error_flags() == 0


DIAG_CASES = [
    # E, N, H, W, F, density, T
    (200, 8, 20, 20, 11, 0.2, 10),
    (64, 32, 32, 32, 11, 0.3, 6),
    (37, 7, 40, 40, 9, 0.25, 6),       # ragged, W > 32
    (64, 12, 8, 8, 5, 0.05, 16),       # crowded: many crossings
    (8, 140, 70, 70, 7, 0.1, 4),       # N > 128, map > 64 (shared-memory BFS with 8 neighbours)
    (16, 9, 16, 16, 6, 0.2, 6),        # generic observation kernel
]


@pytest.mark.parametrize("case", DIAG_CASES, ids=lambda c: "E%d_N%d_%dx%d_F%d" % c[:5])
def test_primal_diagonal_batch_matches_oracle(case):
    from mapf_marl_b200 import maps
    from oracle.oracle import MODE_PRIMAL
    E, N, H, W, F, dens, T = case
    obst, starts, goals = maps.synthetic_batch(300 + E, E, H, W, dens, N, distinct=0)
    eng = _engine(E, N, H, W, mode="primal", fov=F, diagonal_movement=True)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=F)
    orc.set_diagonal(True)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    assert np.array_equal(_np(eng.avail()), orc.primal_avail())
    rs = np.random.RandomState(E + N)
    for t in range(T):
        a = rs.randint(0, 9, (E, N)).astype(np.uint8)
        ad = torch.as_tensor(a.astype(np.int64) if t % 3 == 1 else a, device="cuda")
        if t % 2 == 0:
            out = eng.step_observe(ad, want=PRIMAL_WANT)
            obs, vec = out["obs"], out["vec"]
        else:
            out = eng.step(ad, want=("status", "agent_reward", "reward", "dones", "valid", "avail", "terminated"))
            obs, vec = eng.observe()
        ref = orc.primal_sweep(a)
        robs, rvec = orc.primal_observe()
        for k in ("status", "dones", "valid", "done_mid", "next_mid", "avail", "terminated"):
            if k in out:
                assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        assert np.array_equal(_bits(_np(out["agent_reward"])), _bits(ref["agent_reward"])), t
        assert np.array_equal(_bits(_np(out["reward"])), _bits(ref["reward"])), t
        assert np.array_equal(_np(eng.positions()), orc.positions()), t
        assert np.array_equal(_np(obs), robs), t
        assert np.array_equal(_bits(_np(vec)), _bits(rvec)), t
    # partial sweeps (one _step call at a time) against the same oracle
    a = rs.randint(0, 9, (E, N)).astype(np.uint8)
    for lo in range(0, N, max(1, N // 3)):
        hi = min(N, lo + max(1, N // 3))
        out = eng.step(torch.as_tensor(a, device="cuda"), want=PRIMAL_WANT, agent_range=(lo, hi))
        ref = orc.primal_sweep(a, lo=lo, hi=hi)
        for k in ("status", "dones", "valid", "done_mid", "next_mid", "terminated"):
            assert np.array_equal(_np(out[k])[:, lo:hi] if _np(out[k]).ndim > 1 else _np(out[k]),
                                  ref[k][:, lo:hi] if ref[k].ndim > 1 else ref[k]), (k, lo)
        assert np.array_equal(_np(eng.positions()), orc.positions()), lo
    assert np.array_equal(_np(eng.goal_dist(primal_costs=True)), orc.goal_dist(primal_costs=True))
    assert np.array_equal(_np(eng.goal_dist()), orc.goal_dist())
    assert eng.error_flags() == 0
    bad = a.copy()
    bad[0, 0] = 9
    eng.step(torch.as_tensor(bad, device="cuda"))
    from mapf_marl_b200 import _lib
    assert eng.error_flags() == _lib.FLAG_BAD_ACTION


def test_mapfenv_dropin_class_diagonal_movement():
    from mapf_marl_b200.mapf_primal import MAPFEnv
    g = load_golden("primald_crowd")
    N = g["starts"].shape[0]
    F = int(g["fov"])
    world0 = -g["obst"].astype(int)
    goals0 = np.zeros_like(world0)
    for k in range(N):
        world0[tuple(g["starts"][k])] = k + 1
        goals0[tuple(g["goals"][k])] = k + 1
    env = MAPFEnv(num_agents=N, observation_size=F, world0=world0, goals0=goals0, DIAGONAL_MOVEMENT=True)
    for i in range(1, N + 1):
        assert env._listNextValidActions(i) == [a for a in range(9) if g["avail0"][i - 1, a]]
    for t in range(10):
        for i in range(1, N + 1):
            a = int(g["actions"][t, i - 1])
            state, reward, done, nxt, on_goal, blocking, valid = env._step((i, a))
            assert reward == g["reward"][t, i - 1]
            assert done == bool(g["done_mid"][t, i - 1])
            assert nxt == [k for k in range(9) if g["next_mid"][t, i - 1, k]]
            assert on_goal == bool(g["on_goal"][t, i - 1]) and valid == bool(g["valid"][t, i - 1])
        assert env.getPositions() == [tuple(p) for p in g["pos"][t].tolist()]
    with pytest.raises(AssertionError):
        env._step((1, 9))


def test_fused_step_is_cuda_graph_capturable_and_replays_bit_exactly():
    """The C ABI is stream-ordered and never synchronises: a rollout loop can be captured in a CUDA graph."""
    from mapf_marl_b200 import maps
    from oracle.oracle import MODE_PRIMAL
    E, N, H, W, F, T = 96, 8, 20, 20, 11, 6
    obst, starts, goals = maps.synthetic_batch(77, E, H, W, 0.2, N, distinct=0)
    eng = _engine(E, N, H, W, mode="primal", fov=F)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=F)
    eng.reset(obst, starts, goals)
    orc.reset(obst, starts, goals)
    acts = torch.zeros((E, N), dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):          # warm-up on a side stream (allocates the engine's output buffers)
        eng.observe()
        eng.avail()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    eng.reset(obst, starts, goals)
    with torch.cuda.stream(side):
        out = eng.step_observe(acts, want=PRIMAL_WANT)   # buffers exist before the capture
    torch.cuda.synchronize()
    eng.reset(obst, starts, goals)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        out = eng.step_observe(acts, want=PRIMAL_WANT)
    rs = np.random.RandomState(3)
    for t in range(T):
        a = rs.randint(0, 5, (E, N)).astype(np.uint8)
        acts.copy_(torch.as_tensor(a))
        graph.replay()
        torch.cuda.synchronize()
        ref = orc.primal_sweep(a)
        robs, rvec = orc.primal_observe()
        for k in ("status", "dones", "valid", "done_mid", "next_mid", "avail", "terminated"):
            assert np.array_equal(_np(out[k]), ref[k]), (k, t)
        assert np.array_equal(_bits(_np(out["agent_reward"])), _bits(ref["agent_reward"])), t
        assert np.array_equal(_np(out["obs"]), robs), t
        assert np.array_equal(_bits(_np(out["vec"])), _bits(rvec)), t
    assert eng.error_flags() == 0


# ------------------------------------------------------------------------------------------ mapf_rollout
ROLLOUT_CASES = [
    # (mode, E, N, H, W, F, density, T, obs dtype, engine kwargs)
    ("primal", 300, 8, 20, 20, 11, 0.2, 9, torch.uint8, {}),            # c2 shape, ragged last tile
    ("primal", 64, 32, 32, 32, 11, 0.3, 7, torch.uint8, {}),            # c3 shape
    ("primal", 64, 32, 32, 32, 11, 0.3, 5, "bits", {}),
    ("primal", 16, 32, 32, 32, 11, 0.3, 4, torch.float32, {}),
    ("primal", 48, 6, 12, 12, 5, 0.1, 12, torch.uint8, {}),             # small even-odd FOV, N does not divide 128
    # small batches run through the pipelined kernel (32-agent tiles, step role + observation role):
    ("primal", 37, 4, 12, 12, 7, 0.1, 9, torch.uint8, {}),              # 8 envs per tile, ragged last tile (5 envs)
    ("primal", 5, 16, 24, 24, 9, 0.1, 6, torch.uint8, {}),              # 2 envs per tile, last tile holds one
    ("primal", 16, 2, 8, 8, 3, 0.0, 13, torch.uint8, {}),               # 16 envs per tile, F = 3 (strings share words)
    ("primal", 40, 1, 6, 6, 5, 0.1, 5, "bits", {}),                     # single-agent environments, 32 per tile
    ("primal", 33, 8, 6, 6, 5, 0.05, 20, torch.uint8, {}),              # crowded 6x6 maps: convoys and robot collisions
    ("primal", 12, 32, 40, 40, 11, 0.2, 6, torch.float32, {}),          # one 32-agent environment per tile, W > 32
    ("primal", 8, 128, 64, 64, 11, 0.05, 5, torch.uint8, {}),           # c4 shape: one environment per tile
    ("primal", 4, 140, 40, 40, 11, 0.05, 3, torch.uint8, {}),           # > 128 agents: falls back to T launches
    ("primal", 32, 10, 16, 16, 7, 0.1, 6, torch.uint8, {"diagonal_movement": True}),   # fallback (diagonal mode)
    ("grid", 200, 4, 10, 10, 0, 0.1, 11, None, {"episode_limit": 8}),   # the episode limit is hit inside the rollout
    ("grid", 40, 32, 32, 32, 0, 0.2, 6, None, {}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("mid", [True, False], ids=["mid_outputs", "plain"])
@pytest.mark.parametrize("case", ROLLOUT_CASES, ids=lambda c: "%s_E%d_N%d_%dx%d_T%d_%s" % (c[0], c[1], c[2], c[3], c[4], c[7], str(c[8]).split(".")[-1]))
def test_rollout_equals_consecutive_fused_steps(case, mid):
    """mapf_rollout(T steps, one launch where supported) == T consecutive mapf_step_observe calls, bit for bit, for
    every per-step output and for the state left in the handle."""
    from mapf_marl_b200 import maps
    mode, E, N, H, W, F, dens, T, odt, kw = case
    obst, starts, goals = maps.synthetic_batch(11, E, H, W, dens, N, distinct=0)
    nact = 9 if kw.get("diagonal_movement") else 5
    rs = np.random.RandomState(T * 31 + N)
    acts = torch.as_tensor(rs.randint(0, nact, (T, E, N)).astype(np.uint8), device="cuda")
    # with the mid-sweep outputs the rollout stays in the tile kernel; without them small PRIMAL batches run through
    # the pipelined kernel (mapf_pipe_kernel)
    want = (("reward", "terminated", "agent_reward", "dones", "status", "valid", "avail") +
            (("done_mid", "next_mid") if mid else ())) if mode == "primal" else GRID_WANT
    if mode != "primal" and not mid:
        pytest.skip("GRID has no mid-sweep outputs")
    mk = lambda: _engine(E, N, H, W, mode=mode, fov=F or 11, **kw)   # noqa: E731
    a, b = mk(), mk()
    a.reset(obst, starts, goals)
    b.reset(obst, starts, goals)
    okw = dict(dtype=odt) if odt is not None else {}
    expect_one = mode in ("primal", "grid") and N <= 128 and not kw.get("diagonal_movement")
    assert a.rollout_in_one_launch(odt if odt is not None else torch.uint8) == expect_one
    plan = a.rollout_plan(T, odt if odt is not None else torch.uint8, mid_outputs=mid)
    small_primal = (mode == "primal" and not kw and N <= 32 and F % 2 == 1 and ((32 // N) * N) % 8 == 0)
    assert plan == ("pipelined" if (small_primal and not mid) else ("in_kernel" if expect_one else "per_step"))
    for rnd in range(2):                      # the second rollout starts from the state the first one left
        l0 = a.launch_count()
        ro = a.rollout(acts, want=want, **okw)
        assert a.launch_count() - l0 == (1 if expect_one else T)
        ro = {k: v.clone() for k, v in ro.items()}
        for t in range(T):
            so = b.step_observe(acts[t].clone(), want=want, **okw)   # (a time slice need not be 16-byte aligned)
            for k in so:
                x, y = _np(ro[k][t]), _np(so[k])
                if x.dtype.kind == "f":
                    x, y = _bits(x), _bits(y)
                assert np.array_equal(x.reshape(-1), y.reshape(-1)), (k, t, rnd)
        assert np.array_equal(_np(a.positions()), _np(b.positions()))
        assert np.array_equal(_np(a.dones()), _np(b.dones()))
        assert np.array_equal(_np(a.step_count()), _np(b.step_count()))
        assert np.array_equal(_np(a.avail()), _np(b.avail()))          # prev_action state
    assert a.stats() == b.stats()
    assert a.error_flags() == 0 and b.error_flags() == 0


@pytest.mark.gpu
def test_rollout_partial_mode_and_caller_owned_storage():
    """PARTIAL rollouts run as T launches per kernel but give the same time-major outputs; `out=` writes straight
    into caller-owned (episode-batch style) storage."""
    from mapf_marl_b200 import maps
    E, N, H, W, T = 24, 6, 10, 10, 8
    obst, starts, goals = maps.synthetic_batch(5, E, H, W, 0.0, N, distinct=0)
    kw = dict(mode="partial", episode_limit=30, obs_window=5, obs_knn_agents=3)
    a, b = _engine(E, N, H, W, **kw), _engine(E, N, H, W, **kw)
    a.reset(obst, starts, goals)
    b.reset(obst, starts, goals)
    acts = torch.as_tensor(np.random.RandomState(2).randint(0, 5, (T, E, N)).astype(np.uint8), device="cuda")
    store = {"obs": torch.zeros((T, E, N, a.obs_size), dtype=torch.float32, device="cuda"),
             "reward": torch.zeros((T, E), dtype=torch.float64, device="cuda")}
    ro = a.rollout(acts, want=("reward", "terminated", "avail"), dtype=torch.float32, out=store)
    assert ro["obs"].data_ptr() == store["obs"].data_ptr() and ro["reward"].data_ptr() == store["reward"].data_ptr()
    for t in range(T):
        so = b.step_observe(acts[t], want=("reward", "terminated", "avail"), dtype=torch.float32)
        assert np.array_equal(_np(store["obs"][t]).view(np.uint32), _np(so["obs"]).view(np.uint32)), t
        assert np.array_equal(_bits(_np(store["reward"][t])), _bits(_np(so["reward"]))), t
        assert np.array_equal(_np(ro["terminated"][t]), _np(so["terminated"])), t
        assert np.array_equal(_np(ro["avail"][t]), _np(so["avail"])), t
    # step_observe(out=...) into a time slice of the same storage
    c = _engine(E, N, H, W, **kw)
    c.reset(obst, starts, goals)
    store2 = torch.zeros_like(store["obs"])
    for t in range(T):
        so = c.step_observe(acts[t], want=("reward",), dtype=torch.float32, out={"obs": store2[t]})
        assert so["obs"].data_ptr() == store2[t].data_ptr()
    assert torch.equal(store2, store["obs"])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [("c2", 4096, 0), ("c3", 1024, 3 * 1024)], ids=["c2_full", "c3_shape_E1024"])
def test_pipelined_rollout_matches_oracle_at_full_size(cfg):
    """mapf_rollout through the pipelined kernel (mapf_pipe_kernel) against the CPU oracle, every environment and every
    one of 16 steps: the oracle is stepped first with actions drawn from ITS action masks on odd steps (uniform on even
    ones), the recorded action tensor is handed to one mapf_rollout call, and all time-major outputs are compared."""
    from mapf_marl_b200 import workloads
    from oracle.oracle import MODE_PRIMAL
    name, E, lo = cfg
    wl = workloads.WORKLOADS[name]
    N, H, W, F, T = wl["N"], wl["H"], wl["W"], wl["F"], 16
    obst, starts, goals = workloads.make_world(wl, E, lo, distinct=0)
    orc = _oracle(E, N, H, W, MODE_PRIMAL, fov=F)
    orc.reset(obst, starts, goals)
    avail = orc.primal_avail()
    acts = np.zeros((T, E, N), np.uint8)
    ref = []
    for t in range(T):
        acts[t] = workloads.hash_actions_np(321, range(lo, lo + E), t, N, avail=avail if t % 2 else None)
        out = orc.primal_sweep(acts[t])
        obs, vec = orc.primal_observe()
        ref.append((out, obs, vec, orc.positions().copy()))
        avail = out["avail"]
    eng = _engine(E, N, H, W, mode="primal", fov=F)
    eng.reset(obst, starts, goals)
    want = ("reward", "terminated", "agent_reward", "dones", "status", "valid", "avail")
    assert eng.rollout_plan(T) == "pipelined"
    l0 = eng.launch_count()
    ro = eng.rollout(torch.as_tensor(acts, device="cuda"), want=want)
    assert eng.launch_count() - l0 == 1
    for t in range(T):
        out, obs, vec, pos = ref[t]
        for k in ("terminated", "dones", "status", "valid", "avail"):
            assert np.array_equal(_np(ro[k][t]), out[k]), (k, t)
        assert np.array_equal(_bits(_np(ro["agent_reward"][t])), _bits(out["agent_reward"])), t
        assert np.array_equal(_bits(_np(ro["reward"][t])), _bits(out["reward"])), t
        assert np.array_equal(_np(ro["obs"][t]), obs), t
        assert np.array_equal(_bits(_np(ro["vec"][t])), _bits(vec)), t
    assert np.array_equal(_np(eng.positions()), ref[-1][3])
    assert np.array_equal(_np(eng.step_count()), np.full(E, T, np.int32))
    st = eng.stats()
    assert st["env_steps"] == E * T and st["agent_steps"] == E * N * T
    assert st["goal_arrivals"] > 0 and eng.error_flags() == 0
    # the handle continues from the rollout's final state with single steps
    a = workloads.hash_actions_np(321, range(lo, lo + E), T, N)
    so = eng.step_observe(torch.as_tensor(a, device="cuda"), want=want)
    out = orc.primal_sweep(a)
    assert np.array_equal(_np(so["status"]), out["status"]) and np.array_equal(_np(so["avail"]), out["avail"])
    assert np.array_equal(_np(so["obs"]), orc.primal_observe()[0])


@pytest.mark.gpu
def test_rollout_optional_outputs_single_step_and_long_horizons():
    """The pipelined kernel with outputs left out (no masks, no goal vector, no per-agent outputs), a one-step rollout,
    a 150-step rollout (many buffer hand-overs) and a rollout after a masked reset: all equal consecutive fused steps."""
    from mapf_marl_b200 import maps
    E, N, H, W, F = 96, 8, 14, 14, 7
    obst, starts, goals = maps.synthetic_batch(31, E, H, W, 0.15, N, distinct=0)
    a = _engine(E, N, H, W, mode="primal", fov=F)
    b = _engine(E, N, H, W, mode="primal", fov=F)
    a.reset(obst, starts, goals)
    b.reset(obst, starts, goals)
    rs = np.random.RandomState(9)

    def compare(T, want, want_vec, dtype):
        acts = torch.as_tensor(rs.randint(0, 5, (T, E, N)).astype(np.uint8), device="cuda")
        ro = a.rollout(acts, want=want, dtype=dtype, want_vec=want_vec)
        ro = {k: v.clone() for k, v in ro.items()}
        assert ("vec" in ro) == want_vec
        for t in range(T):
            so = b.step_observe(acts[t].clone(), want=want, dtype=dtype, want_vec=want_vec)
            for k in so:
                assert torch.equal(ro[k][t].reshape(-1).view(torch.uint8), so[k].reshape(-1).view(torch.uint8)), (k, t)
        assert torch.equal(a.positions(), b.positions()) and torch.equal(a.avail(), b.avail())

    assert a.rollout_plan(4) == "pipelined"
    compare(4, ("reward",), False, torch.uint8)                     # no masks, no goal vector
    compare(3, (), True, "bits")                                    # observation + goal vector only
    compare(1, ("reward", "terminated", "dones", "avail"), True, torch.uint8)      # a single step
    compare(150, ("terminated", "avail", "status"), True, torch.uint8)             # long horizon
    mask = (rs.rand(E) < 0.4).astype(np.uint8)
    a.reset(None, None, None, env_mask=mask)
    b.reset(None, None, None, env_mask=mask)
    compare(6, ("reward", "terminated", "dones", "avail", "agent_reward", "valid"), True, torch.float32)
    assert a.stats() == b.stats() and a.error_flags() == 0
    # a bad action inside a rollout raises the device flag like a single step does
    bad = torch.zeros((2, E, N), dtype=torch.uint8, device="cuda")
    bad[1, 5, 3] = 9
    a.rollout(bad, want=("reward",))
    from mapf_marl_b200 import _lib
    assert a.error_flags() == _lib.FLAG_BAD_ACTION


@pytest.mark.gpu
def test_shared_memory_canaries_are_checked():
    """compute-sanitizer is closed on this pool, so the kernels carry their own bounds evidence: guard words between the
    shared-memory regions of a tile, verified before the kernel exits (MAPF_FLAG_INTERNAL).  Every other test asserts
    error_flags() == 0 after stepping; here the self-test hook overwrites one guard on purpose and the flag must appear
    -- for the step kernel, its in-kernel rollout and the pipelined rollout kernel."""
    import ctypes
    from mapf_marl_b200 import _lib, maps
    E, N, H, W, F = 64, 8, 12, 12, 5
    obst, starts, goals = maps.synthetic_batch(2, E, H, W, 0.1, N, distinct=0)
    eng = _engine(E, N, H, W, mode="primal", fov=F)
    eng.reset(obst, starts, goals)
    hook = eng.lib.mapf_debug_corrupt_canary
    hook.argtypes = [ctypes.c_void_p, ctypes.c_int]
    acts = torch.zeros((4, E, N), dtype=torch.uint8, device="cuda")
    for which in range(4):
        eng.step_observe(acts[0])
        assert eng.error_flags() == 0
        assert hook(eng._h, which) == 0
        eng.step_observe(acts[0])
        assert eng.error_flags() == _lib.FLAG_INTERNAL, which
        eng.step_observe(acts[0])
        assert eng.error_flags() == 0                       # the hook is consumed by one launch
    assert eng.rollout_plan(4) == "pipelined"
    for which in range(6):
        assert hook(eng._h, which) == 0
        eng.rollout(acts)
        assert eng.error_flags() == _lib.FLAG_INTERNAL, which
    hook(eng._h, 1)
    eng.rollout(acts, want=("reward", "done_mid"))          # mid outputs: the step kernel's in-kernel loop
    assert eng.error_flags() == _lib.FLAG_INTERNAL
    eng.rollout(acts)
    assert eng.error_flags() == 0
