#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/*.npz from the LIVE reference.

Run in the build container only:   python tests/golden/gen_golden.py
It imports the unmodified reference classes (see refload.py), drives them with seeded
inputs and records every observable of the step/observation path.  The fixtures are
committed; the tests never import the reference.

Recorded per family (all citations relative to /root/reference):
  GRID   mapf_gridworld.py:70-224   reset/step/get_obs/get_state/get_avail_actions
  PRIMAL mapf_primal.py:103-135, 343-386, 407-499, 549-667
  PDIST  MARL-curve-main/src/envs/marl_partial.py:931-955 (per-goal hop-distance maps)
"""
import contextlib
import io
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refload import REF_SRC, load_reference  # noqa: E402

GRID, PRIMAL, PARTIAL = load_reference()


# --------------------------------------------------------------------------- helpers
def write_movingai(tmpdir, grid_rows, n_lines):
    """grid_rows: list[str] of '.'/'@'.  Writes <tmp>/m.map and <tmp>/s-{1..25}.scen."""
    h, w = len(grid_rows), len(grid_rows[0])
    mp = os.path.join(tmpdir, "m.map")
    with open(mp, "w") as f:
        f.write("type octile\nheight %d\nwidth %d\nmap\n" % (h, w))
        for r in grid_rows:
            f.write(r + "\n")
    free = [(i, j) for i in range(h) for j in range(w) if grid_rows[i][j] == "."]
    for k in range(1, 26):
        with open(os.path.join(tmpdir, "s-%d.scen" % k), "w") as f:
            f.write("version 1\n")
            for n in range(n_lines):
                a = free[(n * 7 + k) % len(free)]
                b = free[(n * 13 + 3 * k + 1) % len(free)]
                f.write("0\tm.map\t%d\t%d\t%d\t%d\t%d\t%d\t1.0\n" % (w, h, a[0], a[1], b[0], b[1]))
    return mp, os.path.join(tmpdir, "s-")


def rows_from_map(m):
    return ["".join("@" if v else "." for v in row) for row in m]


def free_cells(obst):
    return [(i, j) for i in range(obst.shape[0]) for j in range(obst.shape[1]) if not obst[i, j]]


# --------------------------------------------------------------------------- GRID
def run_grid(name, obst, starts, goals, actions, step_reward=-0.01, collide_reward=-10,
             episode_limit=10000):
    """obst: bool[H,W]; starts/goals: list of (p0,p1) indexing _full_obs[p0][p1]."""
    n = len(starts)
    T = actions.shape[0]
    with tempfile.TemporaryDirectory() as td:
        mp, sp = write_movingai(td, rows_from_map(obst), max(n + 2, 30))
        env = GRID.MAPF_GRID(mp, sp, n_agents=n, episode_limit=episode_limit, seed=1,
                             render="none", step_reward=step_reward,
                             collide_reward=collide_reward, debug=False)
    # pin starts / goals (the reference samples them from a random .scen)
    for i in range(n):
        env._agent_init_pos[i] = tuple(int(v) for v in starts[i])
        env._agent_goal_pos[i] = tuple(int(v) for v in goals[i])
    env.agent_starts = [env._agent_init_pos[i] for i in range(n)]
    env.agent_goals = [env._agent_goal_pos[i] for i in range(n)]
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        obs0 = env.reset()
        avail0 = np.array(env.get_avail_actions(), dtype=np.uint8)
        H, W = obst.shape
        rec = dict(pos=[], node=[], edge=[], dones=[], reward=[], state=[], avail=[], step_count=[], reward_is_int=[])
        for t in range(T):
            r, dones, info = env.step(actions[t])
            obs = env.get_obs()
            st = env.get_state()
            assert obs.shape == (n, H * W) and all((obs[k] == st).all() for k in range(n))
            rec["pos"].append(np.array(env.agent_positions, dtype=np.int16))
            rec["node"].append(np.array(env._node_collision_agents, dtype=np.int32))
            rec["edge"].append(np.array(env._edge_collision_agents, dtype=np.int32))
            rec["dones"].append(np.array(dones, dtype=np.uint8))
            rec["reward"].append(float(r))
            rec["reward_is_int"].append(isinstance(r, int))
            rec["state"].append(st.astype(np.int8))
            rec["avail"].append(np.array(env.get_avail_actions(), dtype=np.uint8))
            rec["step_count"].append(info["_step_count"])
    out = dict(
        family="GRID", obst=obst.astype(np.uint8), starts=np.array(starts, dtype=np.int16),
        goals=np.array(goals, dtype=np.int16), actions=actions.astype(np.uint8),
        step_reward=np.float64(step_reward), collide_reward=np.float64(collide_reward),
        step_is_int=np.int64(isinstance(step_reward, int)), collide_is_int=np.int64(isinstance(collide_reward, int)),
        py_sum_mode=np.int64(sys.version_info >= (3, 12)), reward_is_int=np.array(rec["reward_is_int"], dtype=np.uint8),
        episode_limit=np.int64(episode_limit), obs0=np.asarray(obs0)[0].astype(np.int8),
        avail0=avail0,
        pos=np.array(rec["pos"]), node=np.array(rec["node"]), edge=np.array(rec["edge"]),
        dones=np.array(rec["dones"]), reward=np.array(rec["reward"], dtype=np.float64),
        state=np.array(rec["state"]), avail=np.array(rec["avail"]),
        step_count=np.array(rec["step_count"], dtype=np.int32))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "T=%d N=%d HxW=%dx%d sum_reward=%.6f" % (T, n, obst.shape[0], obst.shape[1],
                                                                 float(np.sum(out["reward"]))))


def gen_grid():
    # KAT B.1
    obst = np.zeros((10, 10), bool)
    obst[5, 5] = True
    run_grid("grid_kat_b1", obst, [(2, 2), (3, 2), (5, 4), (0, 0)], [(9, 9), (2, 2), (9, 0), (0, 1)],
             np.array([[1, 0, 3, 0], [4, 4, 4, 3], [0, 4, 4, 1], [4, 4, 4, 4]]))
    # c1: BASELINE config 1 -- 10x10 empty, 4 agents, seeded random actions
    perm = np.random.RandomState(0).permutation(100)
    cells = [(int(p) // 10, int(p) % 10) for p in perm]
    run_grid("grid_c1", np.zeros((10, 10), bool), cells[:4], cells[4:8],
             np.random.RandomState(1).randint(0, 5, [200, 4]))
    # helloworld_v1 fixture (8x8, 2x2 block, corner starts, diagonal goals), episode limit hit
    obst = np.zeros((8, 8), bool)
    obst[3:5, 3:5] = True
    run_grid("grid_hello8", obst, [(7, 7), (0, 0), (7, 0), (0, 7)], [(0, 0), (7, 7), (0, 7), (7, 0)],
             np.random.RandomState(2).randint(0, 5, [60, 4]), episode_limit=40)
    # random 16x16, float rewards (exercises the f64 add order), 12 agents
    rs = np.random.RandomState(3)
    obst = rs.rand(16, 16) < 0.2
    fc = free_cells(obst)
    idx = rs.permutation(len(fc))
    starts = [fc[i] for i in idx[:12]]
    goals = [fc[i] for i in idx[6:18]]  # some goals are other agents' starts
    run_grid("grid_rand16", obst, starts, goals, rs.randint(0, 5, [80, 12]),
             step_reward=-0.013, collide_reward=-0.7)
    # crowded 6x6, overlapping starts allowed (edge counts > 1, node groups > 2)
    rs = np.random.RandomState(4)
    obst = np.zeros((6, 6), bool)
    obst[2, 2] = obst[3, 4] = True
    fc = free_cells(obst)
    starts = [fc[i] for i in rs.randint(0, len(fc), 10)]
    goals = [fc[i] for i in rs.randint(0, len(fc), 10)]
    run_grid("grid_crowd6", obst, starts, goals, rs.randint(0, 5, [60, 10]))
    # all-int rewards (the sum stays a Python int) and float-collide / int-step (type tracking in sum())
    rs = np.random.RandomState(6)
    obst = rs.rand(9, 9) < 0.15
    fc = free_cells(obst)
    starts = [fc[i] for i in rs.randint(0, len(fc), 7)]
    goals = [fc[i] for i in rs.randint(0, len(fc), 7)]
    acts = rs.randint(0, 5, [50, 7])
    run_grid("grid_intint", obst, starts, goals, acts, step_reward=-1, collide_reward=-10)
    run_grid("grid_fcis", obst, starts, goals, acts, step_reward=-1, collide_reward=-0.3)
    run_grid("grid_ff", obst, starts, goals, acts, step_reward=-0.1, collide_reward=-0.3, episode_limit=30)
    # agents standing ON wall cells (the reference allows it: the .scen x/y transposition produces such starts).  A wall
    # cell holding k agents reads -1 + k in _full_obs, so __is_cell_obstacle (GRID:278) no longer sees an obstacle:
    # neighbours may walk onto it and the avail mask opens up.  4x4 known-answer case first, then a dense random one.
    obst = np.zeros((4, 4), bool)
    obst[1, 1] = True
    run_grid("grid_onwall4", obst, [(1, 1), (0, 1), (2, 1), (3, 3)], [(3, 0), (1, 1), (0, 0), (1, 1)],
             np.array([[4, 1, 0, 0], [4, 4, 4, 0], [3, 4, 4, 0], [4, 0, 1, 2], [0, 1, 4, 1], [1, 4, 3, 0]]))
    rs = np.random.RandomState(8)
    obst = rs.rand(7, 7) < 0.35
    allc = [(i, j) for i in range(7) for j in range(7)]
    starts = [allc[i] for i in rs.randint(0, len(allc), 12)]      # anywhere, walls included
    goals = [allc[i] for i in rs.randint(0, len(allc), 12)]
    run_grid("grid_onwall7", obst, starts, goals, rs.randint(0, 5, [80, 12]), step_reward=-0.5, collide_reward=-3)
    # real MovingAI map + scen through the reference's own parser (x/y quirk included)
    random.seed(11)
    mp = os.path.join(REF_SRC, "mapf_baseline", "mapf-map", "random-32-32-20.map")
    sp = os.path.join(REF_SRC, "mapf_baseline", "scen-random", "random-32-32-20-random-")
    env = GRID.MAPF_GRID(mp, sp, n_agents=32, seed=1, render="none")
    obst = np.array([[c != "." for c in row] for row in env._original_grid])
    run_grid("grid_real32", obst, list(env.agent_starts), list(env.agent_goals),
             np.random.RandomState(5).randint(0, 5, [40, 32]))


# --------------------------------------------------------------------------- PRIMAL
def primal_world(obst, starts, goals):
    world = -obst.astype(int)
    g = np.zeros_like(world)
    for k, (s, gg) in enumerate(zip(starts, goals)):
        world[s] = k + 1
        g[gg] = k + 1
    return world, g


def mask5(lst, n_act=5):
    m = np.zeros(n_act, np.uint8)
    for a in lst:
        m[a] = 1
    return m


def run_primal(name, obst, starts, goals, fov, actions, n_costs=3, blocking=False, greedy=None, diagonal=False):
    n = len(starts)
    T = actions.shape[0]
    world0, goals0 = primal_world(obst, starts, goals)
    env = PRIMAL.MAPFEnv(num_agents=n, observation_size=fov, world0=world0.copy(), goals0=goals0.copy(),
                         DIAGONAL_MOVEMENT=diagonal)
    n_act = 9 if diagonal else 5
    statuses = []
    orig_act = env.world.act

    def act_spy(action, agent_id):
        s = orig_act(action, agent_id)
        statuses.append(s)
        return s

    env.world.act = act_spy

    def observe_all():
        o = np.zeros((n, 4, fov, fov), np.uint8)
        v = np.zeros((n, 3), np.float64)
        for i in range(1, n + 1):
            maps, vec = env._observe(i)
            for c in range(4):
                assert set(np.unique(maps[c])) <= {0.0, 1.0}
                o[i - 1, c] = maps[c].astype(np.uint8)
            v[i - 1] = vec
        return o, v

    obs0, vec0 = observe_all()
    avail0 = np.array([mask5(env._listNextValidActions(i), n_act) for i in range(1, n + 1)])
    costs0 = np.array([env.getAstarCosts(env.world.getPos(i), env.world.getGoal(i))
                       for i in range(1, min(n, n_costs) + 1)], dtype=np.int32)
    rec = {k: [] for k in ("status", "reward", "done_mid", "next_mid", "on_goal", "valid",
                           "pos", "obs", "vec", "avail", "done", "blocking")}
    if blocking:
        os.environ["MAPF_REF_BFS_MSTAR"] = "1"
    else:
        os.environ.pop("MAPF_REF_BFS_MSTAR", None)
    actions = actions.copy()
    for t in range(T):
        row = {k: [] for k in ("reward", "done_mid", "next_mid", "on_goal", "valid", "blocking")}
        statuses.clear()
        if greedy is not None:      # walk down the goal-distance map with probability `greedy`, stay on the goal
            rsg = np.random.RandomState(1000 + t)
            for i in range(1, n + 1):
                if rsg.rand() < greedy:
                    p = env.world.getPos(i)
                    if p == env.world.getGoal(i):
                        actions[t, i - 1] = 0
                    else:
                        costs = env.getAstarCosts(p, env.world.getGoal(i))
                        best, bd = 0, costs[p]
                        for a_, (dx, dy) in ((1, (0, 1)), (2, (1, 0)), (3, (0, -1)), (4, (-1, 0))):
                            q = (p[0] + dx, p[1] + dy)
                            if 0 <= q[0] < obst.shape[0] and 0 <= q[1] < obst.shape[1] and not obst[q] \
                                    and env.world.state[q] == 0 and costs[q] < bd:
                                best, bd = a_, costs[q]
                        actions[t, i - 1] = best
        for i in range(1, n + 1):
            state, reward, done, nxt, on_goal, blk, valid = env._step((i, int(actions[t, i - 1])))
            assert blocking or blk is False
            row["blocking"].append(bool(blk))
            row["reward"].append(float(reward))
            row["done_mid"].append(bool(done))
            row["next_mid"].append(mask5(nxt, n_act))
            row["on_goal"].append(bool(on_goal))
            row["valid"].append(bool(valid))
        rec["status"].append(np.array(statuses, dtype=np.int8))
        rec["reward"].append(np.array(row["reward"], dtype=np.float64))
        rec["done_mid"].append(np.array(row["done_mid"], dtype=np.uint8))
        rec["next_mid"].append(np.array(row["next_mid"], dtype=np.uint8))
        rec["on_goal"].append(np.array(row["on_goal"], dtype=np.uint8))
        rec["valid"].append(np.array(row["valid"], dtype=np.uint8))
        rec["blocking"].append(np.array(row["blocking"], dtype=np.uint8))
        rec["pos"].append(np.array(env.getPositions(), dtype=np.int16))
        o, v = observe_all()
        rec["obs"].append(o)
        rec["vec"].append(v)
        rec["avail"].append(np.array([mask5(env._listNextValidActions(i, int(actions[t, i - 1])), n_act)
                                      for i in range(1, n + 1)]))
        rec["done"].append(bool(env.world.done()))
    costsT = np.array([env.getAstarCosts(env.world.getPos(i), env.world.getGoal(i))
                       for i in range(1, min(n, n_costs) + 1)], dtype=np.int32)
    out = dict(family="PRIMAL", obst=obst.astype(np.uint8), starts=np.array(starts, dtype=np.int16),
               goals=np.array(goals, dtype=np.int16), fov=np.int64(fov), actions=actions.astype(np.uint8),
               obs0=obs0, vec0=vec0, avail0=avail0, costs0=costs0, costsT=costsT,
               blocking_enabled=np.int64(bool(blocking)), diagonal=np.int64(bool(diagonal)),
               **{k: np.array(v) for k, v in rec.items()})
    out["done"] = out["done"].astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    os.environ.pop("MAPF_REF_BFS_MSTAR", None)
    print("wrote", name, "T=%d N=%d HxW=%dx%d F=%d moved_frac=%.2f collisions=%d blocking_events=%d" % (
        T, n, obst.shape[0], obst.shape[1], fov, float(np.mean(out["status"] >= 0)),
        int(np.sum(out["status"] == -3)), int(np.sum(out["blocking"]))))


def rand_primal_case(seed, h, w, density, n):
    rs = np.random.RandomState(seed)
    obst = rs.rand(h, w) < density
    fc = free_cells(obst)
    idx = rs.permutation(len(fc))
    starts = [fc[i] for i in idx[:n]]
    idg = rs.permutation(len(fc))
    goals = [fc[i] for i in idg[:n]]
    return rs, obst, starts, goals


def gen_primal():
    # KAT B.2
    obst = np.zeros((7, 7), bool)
    obst[3, 3] = True
    run_primal("primal_kat_b2", obst, [(1, 1), (1, 2), (5, 5)], [(1, 3), (1, 1), (5, 5)], 5,
               np.array([[1, 3, 0], [2, 3, 1], [0, 1, 4]]))
    rs, obst, s, g = rand_primal_case(1000, 20, 20, 0.2, 8)        # c2 shape
    run_primal("primal_c2", obst, s, g, 11, rs.randint(0, 5, [60, 8]))
    rs, obst, s, g = rand_primal_case(2000, 32, 32, 0.3, 32)       # c3 shape
    run_primal("primal_c3", obst, s, g, 11, rs.randint(0, 5, [24, 32]))
    rs, obst, s, g = rand_primal_case(3000, 7, 7, 0.1, 12)         # crowded: many robot collisions
    run_primal("primal_crowd", obst, s, g, 5, rs.randint(0, 5, [60, 12]))
    rs, obst, s, g = rand_primal_case(4000, 12, 12, 0.15, 5)       # even FOV (the ctor default is 10)
    run_primal("primal_f10", obst, s, g, 10, rs.randint(0, 5, [30, 5]))
    rs, obst, s, g = rand_primal_case(5000, 9, 9, 0.1, 6)          # tiny FOV
    run_primal("primal_f3", obst, s, g, 3, rs.randint(0, 5, [30, 6]))
    rs, obst, s, g = rand_primal_case(6000, 10, 14, 0.2, 6)        # rectangular world
    run_primal("primal_rect", obst, s, g, 7, rs.randint(0, 5, [30, 6]))
    rs, obst, s, g = rand_primal_case(7000, 40, 40, 0.25, 7)       # N not a multiple of 4, W > 32
    run_primal("primal_n7w40", obst, s, g, 11, rs.randint(0, 5, [20, 7]))


def gen_primal_diagonal():
    """PRIMAL with DIAGONAL_MOVEMENT=True (9 actions, diagonalCollision mapf_primal.py:77-100, 8-connected costs)."""
    rs, obst, s_, g_ = rand_primal_case(9000, 12, 12, 0.15, 8)
    run_primal("primald_12", obst, s_, g_, 7, rs.randint(0, 9, [60, 8]), diagonal=True)
    rs, obst, s_, g_ = rand_primal_case(9100, 7, 7, 0.05, 14)          # crowded: many crossings
    run_primal("primald_crowd", obst, s_, g_, 5, rs.randint(0, 9, [80, 14]), diagonal=True)
    rs, obst, s_, g_ = rand_primal_case(9200, 20, 24, 0.2, 12)
    acts = rs.randint(0, 9, [40, 12])
    acts[rs.rand(40, 12) < 0.5] = rs.randint(5, 9)                     # mostly diagonal moves
    run_primal("primald_rect", obst, s_, g_, 11, acts, diagonal=True)
    obst = np.zeros((6, 6), bool)
    run_primal("primald_open6", obst, [(r, c) for r in range(3) for c in range(4)],
               [(5 - r, 5 - c) for r in range(3) for c in range(4)], 3,
               np.random.RandomState(9300).randint(4, 9, [50, 12]), diagonal=True)


def gen_primal_blocking():
    """PRIMAL with the blocking reward ON (mapf_primal.py:513-546), od_mstar3 replaced by the BFS stand-in of
    refload.py: corridor maps where agents parked on their goals lengthen or cut other agents' paths."""
    rows = ["..........",
            ".@@@@.@@@.",
            ".@......@.",
            ".@.@@@@.@.",
            "...@..@...",
            ".@.@..@.@.",
            ".@.@@.@.@.",
            ".@......@.",
            ".@@@.@@@@.",
            ".........."]
    obst = np.array([[c == "@" for c in r] for r in rows])
    fc = free_cells(obst)
    for k, (seed, n, fov, T) in enumerate(((11, 6, 9, 60), (12, 9, 11, 50), (13, 4, 5, 40))):
        rs = np.random.RandomState(seed)
        idx = rs.permutation(len(fc))
        starts = [fc[i] for i in idx[:n]]
        goals = [fc[i] for i in rs.permutation(len(fc))[:n]]
        run_primal("primalb_corridor%d" % k, obst, starts, goals, fov, rs.randint(0, 5, [T, n]),
                   blocking=True, greedy=0.9)
    rs, obst, s_, g_ = rand_primal_case(8000, 20, 20, 0.25, 10)
    run_primal("primalb_rand20", obst, s_, g_, 11, rs.randint(0, 5, [60, 10]), blocking=True, greedy=0.85)


# --------------------------------------------------------------------------- PARTIAL distance maps
def run_pdist(name, map_file, scen_prefix, n, seed):
    random.seed(seed)
    mp = os.path.join(REF_SRC, "mapf_baseline", "mapf-map", map_file)
    sp = os.path.join(REF_SRC, "mapf_baseline", "scen-random", scen_prefix)
    env = PARTIAL.MARL_PARTIAL_ENV(mp, sp, n_agents=n, obs_window=5, obs_knn_agents=min(5, n), render="none")
    grid = env._original_grid
    H, W = len(grid), len(grid[0])
    obst = np.array([[c != "." for c in row] for row in grid])
    dist = np.full((n, H, W), -1, np.int32)
    for a in range(n):
        for num, d in env._goal_dist[a].items():
            dist[a, num // W, num % W] = d
    np.savez_compressed(os.path.join(HERE, name + ".npz"), family="PDIST", obst=obst.astype(np.uint8),
                        goals=np.array(env._agent_goal_pos, dtype=np.int16),
                        starts=np.array(env._agent_init_pos, dtype=np.int16), dist=dist)
    print("wrote", name, "N=%d HxW=%dx%d maxdist=%d" % (n, H, W, dist.max()))


def gen_pdist():
    run_pdist("pdist_empty8", "empty-8-8.map", "empty-8-8-random-", 10, 21)
    run_pdist("pdist_rand32", "random-32-32-20.map", "random-32-32-20-random-", 4, 22)


# --------------------------------------------------------------------------- PARTIAL step / observation traces
def run_partial(name, map_file, scen_prefix, n, T, seed, policy, **kw):
    """Traces of MARL_PARTIAL_ENV (the env the reference actually registers): step :169-310, get_obs :319-382,
    get_state :384-393, avail :409-434.  policy: 'random' or 'greedy' (walk down the goal-distance map)."""
    random.seed(seed)
    mp = os.path.join(REF_SRC, "mapf_baseline", "mapf-map", map_file)
    sp = os.path.join(REF_SRC, "mapf_baseline", "scen-random", scen_prefix)
    env = PARTIAL.MARL_PARTIAL_ENV(mp, sp, n_agents=n, render="none", **kw)
    obs0 = env.reset()
    grid = env._original_grid
    H, W = len(grid), len(grid[0])
    obst = np.array([[c != "." for c in row] for row in grid])
    starts = np.array(env._agent_init_pos, dtype=np.int16)
    goals = np.array(env._agent_goal_pos, dtype=np.int16)
    dist = np.full((n, H, W), -1, np.int32)
    for a in range(n):
        for num, d in env._goal_dist[a].items():
            dist[a, num // W, num % W] = d
    rs = np.random.RandomState(seed)
    delta = {0: (-1, 0), 1: (1, 0), 2: (0, -1), 3: (0, 1)}
    rec = {k: [] for k in ("actions", "pos", "node", "edge", "at_goal", "dones", "reward", "terminated", "obs",
                           "state", "avail", "goal_cost", "agent_steps")}
    avail0 = np.array(env.get_avail_actions(), dtype=np.uint8)
    state0 = np.array(env.get_state(), dtype=np.int64)
    extra = 0
    for t in range(T):
        acts = rs.randint(0, 5, n)
        if policy == "greedy":
            for a in range(n):
                if rs.rand() < 0.85:
                    p = env._agent_positions[a]
                    if p == env._agent_goal_pos[a]:
                        acts[a] = 4
                    else:
                        best, bd = 4, dist[a, p[0], p[1]]
                        for k, (dr, dc) in delta.items():
                            q = (p[0] + dr, p[1] + dc)
                            if 0 <= q[0] < H and 0 <= q[1] < W and not obst[q] and dist[a, q[0], q[1]] < bd:
                                best, bd = k, dist[a, q[0], q[1]]
                        acts[a] = best
        r, term, info = env.step(acts)
        rec["actions"].append(acts.astype(np.uint8))
        rec["pos"].append(np.array(env._agent_positions, dtype=np.int16))
        rec["node"].append(np.array(env._node_collision_agents, dtype=np.int32))
        rec["edge"].append(np.array(env._edge_collision_agents, dtype=np.int32))
        rec["at_goal"].append(np.array(env._agent_at_goals, dtype=np.uint8))
        rec["dones"].append(np.array(env._agent_dones, dtype=np.uint8))
        rec["reward"].append(float(r))
        rec["terminated"].append(bool(term))
        rec["obs"].append(np.array(env.get_obs(), dtype=np.float64))
        rec["state"].append(np.array(env.get_state(), dtype=np.int64))
        rec["avail"].append(np.array(env.get_avail_actions(), dtype=np.uint8))
        rec["goal_cost"].append(np.array(env._each_goal_cost, dtype=np.int32))
        rec["agent_steps"].append(np.array(env._agent_step_count, dtype=np.int32))
        assert info == {"_step_count": t + 1}
        if term:
            extra += 1
            if extra > 2:
                break
    cfg = dict(obs_window=5, obs_knn_agents=5, episode_limit=100, move_reward=-0.01, stay_reward=-0.02,
               stay_goal_reward=0, node_collide_reward=-1, edge_collide_reward=-1, env_collide_reward=-1,
               complete_reward=1000, complete_fac=1.5, gamma=0.99)
    cfg.update(kw)
    out = dict(family="PARTIAL", obst=obst.astype(np.uint8), starts=starts, goals=goals, dist=dist,
               obs0=np.array(obs0, dtype=np.float64), avail0=avail0, state0=state0,
               py_sum_mode=np.int64(sys.version_info >= (3, 12)),
               **{k: np.array(v) for k, v in rec.items()})
    for k, v in cfg.items():
        out["cfg_" + k] = np.float64(v)
        out["cfgint_" + k] = np.int64(isinstance(v, int))
    out["terminated"] = out["terminated"].astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "T=%d N=%d HxW=%dx%d terminated_at=%s collisions=%d sum_reward=%.4f" % (
        len(rec["reward"]), n, H, W, (int(np.argmax(out["terminated"])) + 1) if out["terminated"].any() else None,
        int(out["state"][-1][0]), float(np.sum(out["reward"]))))


def gen_partial():
    yaml_like = dict(move_reward=0, stay_reward=-0.1, stay_goal_reward=1, node_collide_reward=-2000,
                     edge_collide_reward=-2000, env_collide_reward=-2000, complete_reward=1000, complete_fac=1.5,
                     gamma=0.99)
    run_partial("partial_empty8_yaml", "empty-8-8.map", "empty-8-8-random-", 6, 60, 31, "greedy",
                obs_window=5, obs_knn_agents=5, episode_limit=100, **yaml_like)
    run_partial("partial_empty8_crowd", "empty-8-8.map", "empty-8-8-random-", 12, 40, 32, "random",
                obs_window=3, obs_knn_agents=3, episode_limit=25)
    run_partial("partial_rand32", "random-32-32-20.map", "random-32-32-20-random-", 5, 40, 33, "greedy",
                obs_window=11, obs_knn_agents=8, episode_limit=60)
    run_partial("partial_k1", "empty-8-8.map", "empty-8-8-random-", 4, 30, 34, "greedy",
                obs_window=4, obs_knn_agents=1, episode_limit=40, move_reward=-0.05, stay_reward=-0.25,
                stay_goal_reward=0.5, node_collide_reward=-1.5, edge_collide_reward=-2.5, env_collide_reward=-0.75,
                complete_reward=10.0, complete_fac=1.1, gamma=0.9)


if __name__ == "__main__":
    which = sys.argv[1:] or ["grid", "primal", "primalb", "primald", "pdist", "partial"]
    if "grid" in which:
        gen_grid()
    if "primal" in which:
        gen_primal()
    if "primalb" in which:
        gen_primal_blocking()
    if "primald" in which:
        gen_primal_diagonal()
    if "pdist" in which:
        gen_pdist()
    if "partial" in which:
        gen_partial()
