"""MAPF_GRID: drop-in for the reference's mapf_gridworld.MAPF_GRID, executed by the B200 engine.

Same constructor keywords (mapf_gridworld.py:21-32), same methods and return conventions
(`reset() -> obs`, `step(actions) -> (reward, dones, info)`, the pymarl MultiAgentEnv getters).
With n_envs == 1 every method returns the same Python / numpy types and shapes as the reference,
so pymarl's EpisodeRunner and the parity tests run unchanged.  With n_envs > 1 the same methods
return device tensors with a leading environment dimension (vector API).

All arithmetic (moves, collisions, rewards, done flags, masks, the map observation) runs in the
CUDA kernels behind MapfEngine; this file only converts types.
"""
import os
import random

import numpy as np
import torch

from . import _lib, maps
from .engine import MapfEngine
from .multiagentenv import MultiAgentEnv

ACTION_MEANING = {0: "LEFT", 1: "RIGHT", 2: "UP", 3: "DOWN", 4: "STAY"}

_STEP_WANT = ("reward", "terminated", "dones", "node", "edge", "status", "avail")


class MAPF_GRID(MultiAgentEnv):
    def __init__(self, grid_file_path, agents_path, n_agents=4, episode_limit: int = 10000, seed=None,
                 render='human', step_reward=-0.01, collide_reward=-10, debug=False,
                 n_envs=1, device=None, starts=None, goals=None, strict=True):
        assert os.path.exists(grid_file_path)
        self._grid_file_path = grid_file_path
        self._agent_path = agents_path
        # same draws, in the same order, as the reference constructor (mapf_gridworld.py:37-38)
        self._seed = random.randint(0, 9999)
        np.random.seed(self._seed)
        if seed:
            self._seed = seed
        self._render_mode = render
        self._debug_mode = debug
        self._strict = strict
        self._n_agents = self.n_agents = n_agents
        self.n_envs = int(n_envs)
        self.agents = [a for a in range(n_agents)]
        self.episode_limit = episode_limit
        self._actions = [0, 1, 2, 3, 4]
        self._step_rew = step_reward
        self._collide_rew = collide_reward
        self._obst = maps.read_movingai_map(grid_file_path)            # __setup_grid, :421-428
        self._grid_shape = self._obst.shape
        H, W = self._grid_shape
        E, N = self.n_envs, n_agents
        st = np.zeros((E, N, 2), np.int16)
        gl = np.zeros((E, N, 2), np.int16)
        for e in range(E):
            if starts is not None and goals is not None:
                s_e = np.asarray(starts, dtype=np.int16).reshape(-1, N, 2)[e if np.ndim(starts) == 3 else 0]
                g_e = np.asarray(goals, dtype=np.int16).reshape(-1, N, 2)[e if np.ndim(goals) == 3 else 0]
            else:
                s_e, g_e = self._sample_scen()
            st[e], gl[e] = s_e, g_e
        self._starts, self._goals = st, gl
        self._publish_starts_goals()
        self.engine = MapfEngine(E, N, H, W, mode="grid", shared_map=True, episode_limit=episode_limit,
                                 step_reward=step_reward, collide_reward=collide_reward, device=device)
        self._loaded = False
        self._step_count = None
        self._agent_dones = None
        self.agent_positions = [(-1, -1) for _ in self.agents]
        self._node_collision_agents = None
        self._edge_collision_agents = None
        self._avail_actions = None
        self._last = None

    # -- __setup_agent, mapf_gridworld.py:430-449: a random .scen (1..25), n random records, fields 4..7
    #    taken as (pos[0], pos[1]) in file order (x first), exactly like the reference.
    def _sample_scen(self):
        path = self._agent_path + str(random.randint(1, 25)) + '.scen'
        assert os.path.exists(path)
        lines = maps.read_scen_lines(path)
        sampled = random.sample(lines, self._n_agents)
        s = np.zeros((self._n_agents, 2), np.int16)
        g = np.zeros((self._n_agents, 2), np.int16)
        for k, line in enumerate(sampled):
            sx, sy, gx, gy = maps.scen_fields(line)
            s[k] = (sx, sy)
            g[k] = (gx, gy)
        return s, g

    def _publish_starts_goals(self):
        s, g = self._starts[0], self._goals[0]
        self._agent_init_pos = {a: (int(s[a, 0]), int(s[a, 1])) for a in self.agents}
        self._agent_goal_pos = {a: (int(g[a, 0]), int(g[a, 1])) for a in self.agents}
        self.agent_starts = [self._agent_init_pos[a] for a in self.agents]
        self.agent_goals = [self._agent_goal_pos[a] for a in self.agents]

    def set_starts_goals(self, starts, goals):
        """Pin start / goal cells ([N,2] or [E,N,2]) instead of the sampled ones; takes effect at reset()."""
        self._starts = np.broadcast_to(np.asarray(starts, np.int16).reshape(-1, self._n_agents, 2),
                                       (self.n_envs, self._n_agents, 2)).copy()
        self._goals = np.broadcast_to(np.asarray(goals, np.int16).reshape(-1, self._n_agents, 2),
                                      (self.n_envs, self._n_agents, 2)).copy()
        self._publish_starts_goals()

    # ------------------------------------------------------------------ MultiAgentEnv
    def reset(self):
        """Returns initial observations (mapf_gridworld.py:70-83)."""
        self.engine.reset(self._obst, self._starts, self._goals)
        flags = self.engine.error_flags()
        # a start cell on an obstacle is legal in the reference (the .scen x/y transposition produces them,
        # mapf_gridworld.py:443-445); only out-of-range cells, which would raise IndexError there, are rejected
        if (flags & _lib.FLAG_BAD_POSITION) and self._strict:
            raise AssertionError("start/goal cells outside the map (device flags 0x%x)" % flags)
        self._step_count = 0
        self._agent_dones = [False for _ in self.agents]
        self._node_collision_agents = [0 for _ in self.agents]
        self._edge_collision_agents = [0 for _ in self.agents]
        self._avail_actions = None
        self._last = None
        self._refresh_positions()
        return self.get_obs()

    def _refresh_positions(self):
        if self.n_envs == 1:
            p = self.engine.positions()[0].cpu().numpy()
            self.agent_positions = [(int(p[a, 0]), int(p[a, 1])) for a in self.agents]

    def step(self, agents_action):
        """Returns reward, terminated, info (mapf_gridworld.py:85-141)."""
        if self.n_envs == 1:
            if isinstance(agents_action, torch.Tensor):
                acts = agents_action.detach().reshape(-1)
            else:
                acts = torch.as_tensor(np.asarray(agents_action)).reshape(-1)
            assert len(acts) == self._n_agents
            if self._strict:   # the reference asserts before touching any state (:91-92)
                assert all(int(a) in ACTION_MEANING for a in acts.cpu().tolist())
            out = self.engine.step(acts.reshape(1, -1), want=_STEP_WANT)
            self._step_count += 1
            reward = out["reward"][0].item()
            if isinstance(self._step_rew, int) and isinstance(self._collide_rew, int):
                reward = int(reward)                      # all-int rewards stay a Python int in the reference
            dones = out["dones"][0].cpu().tolist()
            for a in self.agents:
                self._agent_dones[a] = bool(dones[a])     # the reference returns this very list (:141)
            self._node_collision_agents = [int(v) for v in out["node"][0].cpu().tolist()]
            self._edge_collision_agents = [int(v) for v in out["edge"][0].cpu().tolist()]
            self._avail_actions = out["avail"][0].cpu().tolist()
            self._refresh_positions()
            return reward, self._agent_dones, {'_step_count': self._step_count}
        # vector envs: ONE launch steps and produces what get_obs / get_state / get_avail_actions return
        out = self.engine.step_observe(agents_action, want=_STEP_WANT)
        self._step_count += 1
        self._last = out
        return out["reward"], out["terminated"], {'_step_count': self._step_count, 'dones': out["dones"],
                                                  'node': out["node"], 'edge': out["edge"]}

    # ------------------------------------------------------------------ BatchedRunner protocol (vector envs)
    stay_action = 4

    def rollout_spec(self):
        N, HW = self._n_agents, self.get_state_size()
        # every agent observes the same flattened map (:165-182): stored once as `state`, `obs` is a stride-0 view
        return {"state": ((HW,), torch.int8), "avail_actions": ((N, 5), torch.uint8), "reward": ((1,), torch.float64),
                "terminated": ((1,), torch.uint8),
                "_views": {"obs": lambda tm: tm["state"].unsqueeze(2).expand(-1, -1, N, -1)}}

    def reset_into(self, batch):
        self.reset()
        batch.tm["state"][0].copy_(self.get_state())
        batch.tm["avail_actions"][0].copy_(self.get_avail_actions())

    def step_into(self, actions, t, batch):
        tm = batch.tm
        self._last = self.engine.step_observe(actions, want=("reward", "terminated", "avail"),
                                              out={"obs": tm["state"][t + 1], "avail": tm["avail_actions"][t + 1],
                                                   "reward": tm["reward"][t], "terminated": tm["terminated"][t]})
        self._step_count += 1

    def _state_tensor(self):
        if self.n_envs > 1 and self._last is not None and "obs" in self._last:
            return self._last["obs"]
        return self.engine.observe()[0]

    def get_obs(self):
        """[N, H*W] int64 (n_envs == 1) -- every agent sees the same flattened map (:143-183)."""
        state = self._state_tensor()
        if self.n_envs == 1:
            row = state[0].cpu().numpy().astype(np.int64)
            return np.repeat(row[None, :], self._n_agents, axis=0)
        return state[:, None, :].expand(self.n_envs, self._n_agents, state.shape[1])

    def get_obs_agent(self, agent_id):
        state = self._state_tensor()
        if self.n_envs == 1:
            return state[0].cpu().numpy().astype(np.int64)
        return state

    def get_obs_size(self):
        return self._grid_shape[0] * self._grid_shape[1]

    def get_state(self):
        state = self._state_tensor()
        if self.n_envs == 1:
            return state[0].cpu().numpy().astype(np.int64)
        return state

    def get_state_size(self):
        return self._grid_shape[0] * self._grid_shape[1]

    def get_avail_actions(self):
        if self.n_envs > 1 and self._last is not None and "avail" in self._last:
            return self._last["avail"]
        av = self.engine.avail()
        if self.n_envs == 1:
            self._avail_actions = av[0].cpu().tolist()
            return self._avail_actions
        return av

    def get_avail_agent_actions(self, agent_id):
        av = self.engine.avail()
        if self.n_envs == 1:
            return av[0, agent_id].cpu().tolist()
        return av[:, agent_id]

    def get_total_actions(self):
        return len(self._actions)

    def render(self):
        return None

    def close(self):
        self.engine.close()

    def seed(self):
        pass

    def save_replay(self):
        pass

    def get_env_info(self):
        return {"state_shape": self.get_state_size(),
                "obs_shape": self.get_obs_size(),
                "n_actions": self.get_total_actions(),
                "n_agents": self._n_agents,
                "episode_limit": self.episode_limit}

    def episode_done(self):
        if self.n_envs == 1:
            return sum(self._agent_dones) == self._n_agents
        return self.engine.dones().all(dim=1)

    def get_stats(self):
        """Needed by pymarl's ParallelRunner (parallel_runner.py:255); the reference envs lack it."""
        return self.engine.stats()

    @property
    def _full_obs(self):
        state, _ = self.engine.observe()
        H, W = self._grid_shape
        return state[0].cpu().numpy().astype(np.int64).reshape(H, W).tolist()
