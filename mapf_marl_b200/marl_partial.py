"""MARL_PARTIAL_ENV: drop-in for MARL-curve-main/src/envs/marl_partial.py::MARL_PARTIAL_ENV -- the environment the
reference's pymarl registry actually registers (src/envs/__init__.py:63) -- executed by the B200 engine.

Same constructor keywords (marl_partial.py:26-47) and the same MultiAgentEnv surface: reset() -> obs,
step(actions) -> (reward, terminated, info), get_obs() float64 [N, 2*W*W + 13*K], get_state() =
[total collisions, step count, sum of goal costs], get_avail_actions(), get_env_info(), episode_done().
With n_envs > 1 the same calls return device tensors with a leading environment dimension.

reset() re-samples starts / goals from a random .scen file exactly like the reference (:907-927, (row, col) =
fields (5, 4) / (7, 6)) and rebuilds the per-goal distance maps on the device (the reference spends 0.65 s per
agent there in networkx A*, :931-955).  `output=True` (randomised collision repair) and `visual=True` are not
supported.
"""
import os
import random

import numpy as np
import torch

from . import _lib, maps
from .engine import MapfEngine
from .multiagentenv import MultiAgentEnv

ACTION_MEANING = {0: "LEFT", 1: "RIGHT", 2: "UP", 3: "DOWN", 4: "STAY"}

_WANT = ("reward", "terminated", "dones", "node", "edge", "avail")


class MARL_PARTIAL_ENV(MultiAgentEnv):
    def __init__(self, grid_file_path, agents_path, n_agents=4, obs_window=5, obs_knn_agents=5, episode_limit=100,
                 seed=None, render='human', move_reward=-0.01, stay_reward=-0.02, stay_goal_reward=0,
                 node_collide_reward=-1, edge_collide_reward=-1, env_collide_reward=-1, complete_reward=1000,
                 complete_fac=1.5, debug=False, visual=False, gamma=0.99, output=False, n_envs=1, device=None,
                 strict=True, obs_float32=False):
        assert os.path.exists(grid_file_path)
        if output:
            raise NotImplementedError("output=True (randomised collision repair, marl_partial.py:645-820) is "
                                      "non-deterministic and not implemented by the B200 engine")
        if visual:
            raise NotImplementedError("visual=True writes json files for the reference's web visualiser")
        self._grid_file_path = grid_file_path
        self._agent_path = agents_path
        self._render_mode = render
        self._debug_mode = debug
        self._output_mode = output
        self._strict = strict
        # vector envs only: observations as float32 device tensors (the dtype pymarl's episode batch stores,
        # src/run.py:133-140), written by the kernel itself instead of float64 followed by a cast
        self._obs_dtype = torch.float32 if (obs_float32 and int(n_envs) > 1) else torch.float64
        self._n_agents = self.n_agents = n_agents
        self.n_envs = int(n_envs)
        self._seed = random.randint(0, 9999)         # same draws, in the same order, as marl_partial.py:59-62
        np.random.seed(self._seed)
        if seed:
            self._seed = seed
        self.agents = [a for a in range(n_agents)]
        self._n_features = 13
        self._actions = [0, 1, 2, 3, 4]
        self.episode_limit = episode_limit
        self._obs_knn_agents = obs_knn_agents
        self._obs_window = obs_window
        self._gamma = gamma
        self._obst = maps.read_movingai_map(grid_file_path)
        self._grid_shape = self._obst.shape
        H, W = self._grid_shape
        self._starts = np.zeros((self.n_envs, n_agents, 2), np.int16)
        self._goals = np.zeros((self.n_envs, n_agents, 2), np.int16)
        self._pinned = False
        self._setup_agent()                          # the reference samples once in __init__ (:107) ...
        self.engine = MapfEngine(self.n_envs, n_agents, H, W, mode="partial", shared_map=True,
                                 episode_limit=episode_limit, device=device, obs_window=obs_window,
                                 obs_knn_agents=obs_knn_agents, move_reward=move_reward, stay_reward=stay_reward,
                                 stay_goal_reward=stay_goal_reward, node_collide_reward=node_collide_reward,
                                 edge_collide_reward=edge_collide_reward, env_collide_reward=env_collide_reward,
                                 complete_reward=complete_reward, complete_fac=complete_fac, gamma=gamma)
        self._step_count = None
        self._terminated = False
        self._agent_dones = None
        self._agent_positions = [(-1, -1) for _ in self.agents]
        self._node_collision_agents = None
        self._edge_collision_agents = None
        self._avail_actions = None
        self._vlast = None       # vector envs: outputs of the last fused step (what the getters return)
        self._vstate = None

    # __setup_agent, marl_partial.py:907-929
    def _setup_agent(self):
        if self._pinned:
            return
        for e in range(self.n_envs):
            path = self._agent_path + str(random.randint(1, 25)) + '.scen'
            assert os.path.exists(path)
            lines = maps.read_scen_lines(path)
            assert len(lines) > self._n_agents
            for k, line in enumerate(random.sample(lines, self._n_agents)):
                s_col, s_row, f_col, f_row = maps.scen_fields(line)
                self._starts[e, k] = (s_row, s_col)
                self._goals[e, k] = (f_row, f_col)
        self._publish()

    def _publish(self):
        self._agent_init_pos = [tuple(int(v) for v in p) for p in self._starts[0]]
        self._agent_goal_pos = [tuple(int(v) for v in p) for p in self._goals[0]]

    def set_starts_goals(self, starts, goals):
        """Pin (row, col) starts / goals ([N,2] or [E,N,2]); reset() then stops re-sampling."""
        self._starts = np.broadcast_to(np.asarray(starts, np.int16).reshape(-1, self._n_agents, 2),
                                       (self.n_envs, self._n_agents, 2)).copy()
        self._goals = np.broadcast_to(np.asarray(goals, np.int16).reshape(-1, self._n_agents, 2),
                                      (self.n_envs, self._n_agents, 2)).copy()
        self._pinned = True
        self._publish()

    # ------------------------------------------------------------------ MultiAgentEnv
    def reset(self):
        """Returns initial observations (marl_partial.py:125-167)."""
        self._setup_agent()                          # ... and again at every reset (:130)
        self.engine.reset(self._obst, self._starts, self._goals)     # includes the goal-distance maps
        flags = self.engine.error_flags()
        if (flags & (_lib.FLAG_BAD_POSITION | _lib.FLAG_START_ON_WALL)) and self._strict:
            raise KeyError("start/goal cell is not a free cell of the map (device flags 0x%x)" % flags)
        self._terminated = False
        self._step_count = 0
        self._agent_dones = [False for _ in self.agents]
        self._node_collision_agents = [0 for _ in self.agents]
        self._edge_collision_agents = [0 for _ in self.agents]
        self._vlast = None
        self._refresh()
        return self.get_obs()

    def _refresh(self):
        if self.n_envs == 1:
            p = self.engine.positions()[0].cpu().numpy()
            self._agent_positions = [(int(p[a, 0]), int(p[a, 1])) for a in self.agents]
            self._avail_actions = self.engine.avail()[0].cpu().tolist()

    def step(self, agents_action):
        """Returns reward, terminated, info (marl_partial.py:169-310)."""
        if self.n_envs == 1:
            if isinstance(agents_action, torch.Tensor):
                acts = agents_action.detach().reshape(-1)
            else:
                acts = torch.as_tensor(np.asarray(agents_action)).reshape(-1)
            assert len(acts) == self._n_agents
            if self._strict:
                assert all(int(a) in ACTION_MEANING for a in acts.cpu().tolist())
            out = self.engine.step(acts.reshape(1, -1), want=_WANT)
            self._step_count += 1
            reward = float(out["reward"][0].item())
            self._terminated = bool(out["terminated"][0].item())
            self._agent_dones = [bool(v) for v in out["dones"][0].cpu().tolist()]
            self._node_collision_agents = [int(v) for v in out["node"][0].cpu().tolist()]
            self._edge_collision_agents = [int(v) for v in out["edge"][0].cpu().tolist()]
            self._refresh()
            return reward, self._terminated, {'_step_count': self._step_count}
        # vector envs: ONE call produces everything the getters return (the tile kernel steps, the observation kernel
        # writes obs and the get_state() triple): two launches per environment step instead of four
        if self._vstate is None:
            self._vstate = torch.zeros((self.n_envs, 3), dtype=torch.int64, device=self.engine.device)
        self.engine.bind_partial_state_out(self._vstate)
        out = self.engine.step_observe(agents_action, want=_WANT, dtype=self._obs_dtype)
        self._step_count += 1
        self._vlast = out
        return out["reward"], out["terminated"], {'_step_count': self._step_count, 'dones': out["dones"],
                                                  'node': out["node"], 'edge': out["edge"]}

    # ------------------------------------------------------------------ BatchedRunner protocol (vector envs)
    stay_action = 4

    def rollout_spec(self):
        N = self._n_agents
        return {"obs": ((N, self.get_obs_size()), self._obs_dtype), "state": ((3,), torch.int64),
                "avail_actions": ((N, 5), torch.uint8), "reward": ((1,), torch.float64),
                "terminated": ((1,), torch.uint8)}

    def reset_into(self, batch):
        self.engine.bind_partial_state_out(batch.tm["state"][0])
        obs = self.reset()                                   # engine-owned tensors of the t = 0 observation
        batch.tm["obs"][0].copy_(obs)
        batch.tm["avail_actions"][0].copy_(self.get_avail_actions())
        self._vstate = batch.tm["state"][0]

    def step_into(self, actions, t, batch):
        tm = batch.tm
        self.engine.bind_partial_state_out(tm["state"][t + 1])
        self._vlast = self.engine.step_observe(actions, want=("reward", "terminated", "avail"), dtype=self._obs_dtype,
                                               out={"obs": tm["obs"][t + 1], "avail": tm["avail_actions"][t + 1],
                                                    "reward": tm["reward"][t], "terminated": tm["terminated"][t]})
        self._vstate = tm["state"][t + 1]
        self._step_count += 1

    def get_obs(self):
        if self.n_envs > 1 and self._vlast is not None and "obs" in self._vlast:
            return self._vlast["obs"]
        obs, _ = self.engine.observe(dtype=self._obs_dtype)
        return obs[0].cpu().numpy() if self.n_envs == 1 else obs

    def get_obs_agent(self, agent_id):
        assert agent_id > -1
        obs, _ = self.engine.observe()
        return obs[0, agent_id].cpu().numpy() if self.n_envs == 1 else obs[:, agent_id]

    def get_obs_size(self):
        return 2 * (self._obs_window ** 2) + self._obs_knn_agents * self._n_features

    def get_state(self):
        if self.n_envs > 1 and self._vlast is not None and self._vstate is not None:
            return self._vstate
        st = self.engine.partial_state(want=("state",))["state"]
        return st[0].cpu().numpy() if self.n_envs == 1 else st

    def get_state_size(self):
        return 3

    def get_avail_actions(self):
        if self.n_envs == 1:
            return self._avail_actions
        if self._vlast is not None and "avail" in self._vlast:
            return self._vlast["avail"]
        return self.engine.avail()

    def get_avail_agent_actions(self, agent_id):
        if self.n_envs == 1:
            return self._avail_actions[agent_id]
        return self.engine.avail()[:, agent_id]

    def get_total_actions(self):
        return len(self._actions)

    def agent_pos(self, agent_id):
        assert -1 < agent_id < self._n_agents
        return self._agent_positions[agent_id]

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
        return self.engine.stats()
