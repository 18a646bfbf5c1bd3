"""MAPFEnv: drop-in for the reference's mapf_primal.MAPFEnv (PRIMAL's mapf_gym), executed by the
B200 engine.

Reference surface kept (mapf_primal.py): MAPFEnv(num_agents, observation_size, world0, goals0,
DIAGONAL_MOVEMENT, SIZE, PROB, FULL_HELP, blank_world) :175, _reset :389, _step((id, action)) :549,
_observe(id) :343, _listNextValidActions(id, prev_action) :639, getAstarCosts(start, goal) :407,
getObstacleMap/getGoals/getPositions :233-246, _complete :404, and a `world` view with
getPos/getGoal/done/state/goals.

Batched surface added (n_envs >= 1, device tensors): step_sweep(actions[E,N]) = one
`for id in 1..N: _step((id, a))` sweep per environment; observe_all(); astar_costs().

DIAGONAL_MOVEMENT=True switches the engine to the 9-action sweep with the crossing test
(State.diagonalCollision, :77-100) and 8-connected getAstarCosts.  The blocking reward (row P7) depends on the
un-vendored od_mstar3 planner in the reference; `blocking_reward=True` computes it with BFS path lengths (what a
single-robot M* returns), otherwise rewards equal the reference's with blocking == 0.
"""

import numpy as np
import torch

from . import maps
from .engine import MapfEngine

ACTION_COST, IDLE_COST, GOAL_REWARD, COLLISION_REWARD, FINISH_REWARD, BLOCKING_COST = -0.3, -.5, 0.0, -2., 20., -1.
opposite_actions = {0: -1, 1: 3, 2: 4, 3: 1, 4: 2, 5: 7, 6: 8, 7: 5, 8: 6}
dirDict = {0: (0, 0), 1: (0, 1), 2: (1, 0), 3: (0, -1), 4: (-1, 0), 5: (1, 1), 6: (1, -1), 7: (-1, -1), 8: (-1, 1)}
actionDict = {v: k for k, v in dirDict.items()}

_SWEEP_WANT = ("status", "agent_reward", "dones", "valid", "done_mid", "next_mid", "avail", "terminated",
               "blocking")


class _WorldView(object):
    """Read-only stand-in for the reference's State object (mapf_primal.py:32-165), environment 0."""

    def __init__(self, env):
        self._env = env

    @property
    def num_agents(self):
        return self._env.num_agents

    @property
    def agents(self):
        return self._env.getPositions()

    @property
    def agent_goals(self):
        return self._env.getGoals()

    @property
    def state(self):
        st = -self._env._obst0.astype(int)
        for k, (x, y) in enumerate(self._env.getPositions()):
            st[x, y] = k + 1
        return st

    @property
    def goals(self):
        g = np.zeros(self._env._obst0.shape, int)
        for k, (x, y) in enumerate(self._env.getGoals()):
            g[x, y] = k + 1
        return g

    def getPos(self, agent_id):
        return self._env.getPositions()[agent_id - 1]

    def getGoal(self, agent_id):
        return self._env.getGoals()[agent_id - 1]

    def getDir(self, action):
        return dirDict[action]

    def getAction(self, direction):
        return actionDict[direction]

    def done(self):
        return self._env._complete()


class MAPFEnv(object):
    metadata = {"render.modes": ["human", "ansi"]}

    def getFinishReward(self):
        return FINISH_REWARD

    def __init__(self, num_agents=1, observation_size=10, world0=None, goals0=None, DIAGONAL_MOVEMENT=False,
                 SIZE=(10, 40), PROB=(0, .5), FULL_HELP=False, blank_world=False, n_envs=1, device=None,
                 goal_dist=False, blocking_reward=False):
        if DIAGONAL_MOVEMENT and blocking_reward:
            raise NotImplementedError("blocking_reward with DIAGONAL_MOVEMENT is not implemented by the B200 engine")
        self.num_agents = num_agents
        self.n_envs = int(n_envs)
        self.individual_rewards = [0 for _ in range(num_agents)]
        self.observation_size = observation_size
        self.SIZE = SIZE
        self.PROB = PROB
        self.fresh = True
        self.FULL_HELP = FULL_HELP
        self.finished = False
        self.DIAGONAL_MOVEMENT = bool(DIAGONAL_MOVEMENT)
        self.n_actions = 9 if DIAGONAL_MOVEMENT else 5                  # action_space, :198-201
        self._device = device
        self._goal_dist = goal_dist
        # get_blocking_reward (mapf_primal.py:513-546) with BFS path lengths instead of the un-vendored od_mstar3;
        # off by default because the fixtures pinned against the reference were recorded with it fenced off
        self._blocking_reward = bool(blocking_reward)
        self.engine = None
        self.viewer = None
        self.world = _WorldView(self)
        self._setWorld(world0, goals0, blank_world=blank_world)

    # ------------------------------------------------------------------ world set-up (host side, reset time)
    def _setWorld(self, world0=None, goals0=None, blank_world=False):
        """world0: int [H,W] (or [E,H,W]) with -1 walls and agent ids 1..N; goals0: ids at goal cells
        (mapf_primal.py:248-340).  Without world0 a random world is drawn with the reference's recipe
        (triangular obstacle density, 3-way size choice, goals inside the agent's connected region)."""
        E, N = self.n_envs, self.num_agents
        if world0 is not None:
            if goals0 is None and not blank_world:
                raise Exception("you gave a world with no goals!")
            w = np.asarray(world0)
            w = w.reshape((-1,) + w.shape[-2:])
            if w.shape[0] not in (1, E):
                raise ValueError("world0 must be [H,W] or [n_envs,H,W]")
            obst = (w == -1).astype(np.uint8)
            starts = np.zeros((E, N, 2), np.int16)
            goals = np.zeros((E, N, 2), np.int16)
            if blank_world:
                for e in range(E):
                    rs = np.random.RandomState(np.random.randint(0, 2 ** 31 - 1))
                    starts[e], goals[e] = maps.place_agents_and_goals(rs, obst[e % obst.shape[0]], N)
            else:
                g = np.asarray(goals0)
                g = g.reshape((-1,) + g.shape[-2:])
                for e in range(E):
                    starts[e] = self._scan(w[e % w.shape[0]], N)       # State.scanForAgents, :53-66
                    goals[e] = self._scan(g[e % g.shape[0]], N)
            obst = np.broadcast_to(obst, (E,) + obst.shape[1:]) if obst.shape[0] == 1 else obst
        else:
            prob = np.random.triangular(self.PROB[0], .33 * self.PROB[0] + .66 * self.PROB[1], self.PROB[1])
            size = int(np.random.choice([self.SIZE[0], self.SIZE[0] * .5 + self.SIZE[1] * .5, self.SIZE[1]],
                                        p=[.5, .25, .25]))
            obst = np.zeros((E, size, size), np.uint8)
            starts = np.zeros((E, N, 2), np.int16)
            goals = np.zeros((E, N, 2), np.int16)
            for e in range(E):
                rs = np.random.RandomState(np.random.randint(0, 2 ** 31 - 1))
                obst[e] = maps.random_obstacles(rs, size, size, prob)
                starts[e], goals[e] = maps.place_agents_and_goals(rs, obst[e], N)
        obst = np.ascontiguousarray(obst)
        H, W = obst.shape[-2:]
        if self.engine is None or (self.engine.H, self.engine.W) != (H, W):
            if self.engine is not None:
                self.engine.close()
            self.engine = MapfEngine(E, N, H, W, mode="primal", fov=self.observation_size, device=self._device,
                                     goal_dist=self._goal_dist, action_cost=ACTION_COST, idle_cost=IDLE_COST,
                                     goal_reward=GOAL_REWARD, collision_reward=COLLISION_REWARD,
                                     blocking_reward=self._blocking_reward, blocking_cost=BLOCKING_COST,
                                     diagonal_movement=self.DIAGONAL_MOVEMENT)
        self._obst0 = obst[0]
        self.initial_world = world0
        self.initial_goals = goals0
        self.engine.reset(obst, starts, goals)
        flags = self.engine.error_flags()
        if flags:
            raise AssertionError("invalid world0/goals0 (device flags 0x%x)" % flags)
        self._cache = {}

    @staticmethod
    def _scan(grid, n):
        out = np.full((n, 2), -1, np.int16)
        for (i, j) in np.argwhere(grid > 0):
            out[grid[i, j] - 1] = (i, j)
        assert (out >= 0).all(), "every agent id 1..N must appear exactly once"
        return out

    # ------------------------------------------------------------------ reference surface (environment 0)
    def _positions_np(self):
        if "pos" not in self._cache:
            self._cache["pos"] = self.engine.positions().cpu().numpy()
        return self._cache["pos"]

    def _goals_np(self):
        if "goal" not in self._cache:
            self._cache["goal"] = self.engine.goals().cpu().numpy()
        return self._cache["goal"]

    def getObstacleMap(self):
        return self._obst0.astype(int)

    def getGoals(self):
        g = self._goals_np()[0]
        return [(int(x), int(y)) for x, y in g]

    def getPositions(self):
        p = self._positions_np()[0]
        return [(int(x), int(y)) for x, y in p]

    def _complete(self):
        return bool((self._positions_np()[0] == self._goals_np()[0]).all())

    def _observe(self, agent_id):
        assert agent_id > 0
        if "obs" not in self._cache:
            obs, vec = self.engine.observe()
            self._cache["obs"] = (obs[0].cpu().numpy(), vec[0].cpu().numpy())
        obs, vec = self._cache["obs"]
        maps4 = [obs[agent_id - 1, c].astype(np.float64) for c in range(4)]
        v = vec[agent_id - 1]
        p, g = self.getPositions()[agent_id - 1], self.getGoals()[agent_id - 1]
        dx, dy = g[0] - p[0], g[1] - p[1]
        # the reference returns Python ints for (dx, dy) when the agent stands on its goal (mag == 0)
        return (maps4, [float(v[0]), float(v[1]), float(v[2])] if v[2] != 0 else [dx, dy, float(v[2])])

    def _reset(self, agent_id, world0=None, goals0=None):
        self.finished = False
        self._setWorld(world0, goals0)
        self.fresh = True
        on_goal = self.getPositions()[agent_id - 1] == self.getGoals()[agent_id - 1]
        return self._listNextValidActions(agent_id), on_goal, False

    def _listNextValidActions(self, agent_id, prev_action=0, episode=0):
        # a read-only query like the reference's (prev_action is an argument, mapf_primal.py:639): the engine's stored
        # last actions -- which later avail() calls and sweeps rely on -- stay as they are
        prev = torch.zeros((self.n_envs, self.num_agents), dtype=torch.uint8)
        prev[:, agent_id - 1] = int(prev_action)
        mask = self.engine.avail(prev=prev)[0, agent_id - 1].cpu().tolist()
        return [a for a in range(self.n_actions) if mask[a]]

    def _step(self, action_input, episode=0):
        self.fresh = False
        assert len(action_input) == 2, 'Action input should be a tuple with the form (agent_id, action)'
        assert action_input[1] in range(self.n_actions), 'Invalid action'
        assert action_input[0] in range(1, self.num_agents + 1)
        agent_id, action = int(action_input[0]), int(action_input[1])
        acts = torch.zeros((self.n_envs, self.num_agents), dtype=torch.uint8)
        acts[:, agent_id - 1] = action
        out = self.engine.step(acts, want=_SWEEP_WANT, agent_range=(agent_id - 1, agent_id))
        self._cache = {}
        i = agent_id - 1
        reward = float(out["agent_reward"][0, i].item())
        self.individual_rewards[i] = reward
        state = self._observe(agent_id)
        done = bool(out["done_mid"][0, i].item())
        self.finished |= done
        mask = out["next_mid"][0, i].cpu().tolist()
        nextActions = [a for a in range(self.n_actions) if mask[a]]
        on_goal = bool(out["dones"][0, i].item())
        valid_action = bool(out["valid"][0, i].item())
        blocking = bool(out["blocking"][0, i].item()) if self._blocking_reward else False
        return state, reward, done, nextActions, on_goal, blocking, valid_action

    def getAstarCosts(self, start, goal):
        """Distance-to-`goal` map with getAstarCosts' conventions (walls -1, unreached cells keep `state`).
        `goal` must be the goal of one of the agents (how PRIMAL calls it)."""
        goals = self.getGoals()
        goal = (int(goal[0]), int(goal[1]))
        if goal not in goals:
            raise ValueError("getAstarCosts: goal %s is not the goal of any agent" % (goal,))
        costs = self.engine.goal_dist(primal_costs=True)
        return costs[0, goals.index(goal)].cpu().numpy().astype(int)

    # ------------------------------------------------------------------ batched surface (device tensors)
    def step_sweep(self, actions, want=("agent_reward", "dones", "valid", "avail", "terminated")):
        """One `for id in 1..N: _step((id, actions[:, id-1]))` sweep in every environment."""
        self.fresh = False
        self._cache = {}
        return self.engine.step(actions, want=want)

    def step_sweep_observe(self, actions, want=("agent_reward", "dones", "valid", "avail", "terminated"),
                           dtype=torch.uint8):
        self.fresh = False
        self._cache = {}
        return self.engine.step_observe(actions, want=want, dtype=dtype)

    def observe_all(self, dtype=torch.uint8):
        """(obs [E,N,4,F,F], vec [E,N,3]) == `_observe(id)` for every id, after the sweep."""
        return self.engine.observe(dtype=dtype)

    def astar_costs(self, primal_costs=False):
        """int16 [E,N,H,W] hop distance to every agent's goal."""
        return self.engine.goal_dist(primal_costs=primal_costs)

    def close(self):
        if self.engine is not None:
            self.engine.close()
