"""PrimalVecEnv: the PRIMAL environment (mapf_primal.MAPFEnv semantics: ordered sweep, 4-channel FOV observation,
goal vector) as a device-resident VECTOR environment with the MultiAgentEnv-style surface pymarl's runners use.

The reference's MAPFEnv is a gym-style single-agent-step class (mapf_primal.py:175-667), not a MultiAgentEnv, so this
adapter defines the joint-step surface the way the reference's pymarl envs do (MARL-curve-main/src/envs/
multiagentenv.py:1-68): one `step(actions[E, N])` = the sweep `for id in 1..N: _step((id, a))`, reward = the team
reward of the sweep, terminated = State.done() (all agents on goal, mapf_primal.py:159-165) or the episode limit.

    obs        uint8  [E, N, 4*F*F]   poss / goal / goals / obstacle maps of `_observe` (:343-386), flattened
    obs_vec    float64[E, N, 3]       [dx/mag, dy/mag, mag] (:380-385)
    state      float64[E, N*3]        the goal vectors of all agents (a view of obs_vec: MAPFEnv has no get_state)
    avail      uint8  [E, N, 5]       `_listNextValidActions(id, last action)` (:639-667)

This is the env behind BASELINE config 5 ("feeding the pymarl QMIX rollout"): BatchedRunner steps it with ONE kernel
launch per environment step, the kernel writing straight into the episode batch.
"""
import numpy as np
import torch

from .engine import MapfEngine


class PrimalVecEnv:
    stay_action = 0                      # dirDict[0] = (0, 0), mapf_primal.py:28
    graph_safe_reset = True              # reset_into() launches kernels on device tensors only (BatchedRunner)

    def __init__(self, obst, starts, goals, fov=11, episode_limit=256, device=None, shared_map=False, **engine_kwargs):
        starts = np.asarray(starts)
        self.n_envs, self.n_agents = int(starts.shape[0]), int(starts.shape[1])
        H, W = (obst.shape if shared_map else obst.shape[1:])
        self.fov = int(fov)
        self.episode_limit = int(episode_limit)
        self.engine = MapfEngine(self.n_envs, self.n_agents, H, W, mode="primal", fov=fov, shared_map=shared_map,
                                 device=device, **engine_kwargs)
        # the world stays on the device: an episode reset is two small kernels, no host copy
        dev = self.engine.device
        self._world = (torch.as_tensor(np.asarray(obst), device=dev).to(torch.int8).contiguous(),
                       torch.as_tensor(starts, device=dev).to(torch.int16).contiguous(),
                       torch.as_tensor(np.asarray(goals), device=dev).to(torch.int16).contiguous())
        self._last = None
        self._t = 0
        self._maps_loaded = False        # the obstacle rows are built by the first reset only (the world is constant)

    def _engine_reset(self):
        obst, starts, goals = self._world
        self.engine.reset(None if self._maps_loaded else obst, starts, goals)
        self._maps_loaded = True

    # ------------------------------------------------------------------ MultiAgentEnv-style surface (vector)
    def get_env_info(self):
        return {"state_shape": 3 * self.n_agents, "obs_shape": 4 * self.fov * self.fov + 3, "n_actions": 5,
                "n_agents": self.n_agents, "episode_limit": self.episode_limit}

    def reset(self):
        self._engine_reset()
        self._t = 0
        obs, vec = self.engine.observe()
        self._last = {"obs": obs, "vec": vec, "avail": self.engine.avail()}
        return self.get_obs()

    def step(self, actions):
        """-> (reward float64 [E], terminated uint8 [E], info); ONE launch, observation and masks included."""
        out = self.engine.step_observe(actions, want=("reward", "terminated", "dones", "avail"))
        self._t += 1
        self._last = out
        term = out["terminated"] if self._t < self.episode_limit else torch.ones_like(out["terminated"])
        return out["reward"], term, {"dones": out["dones"]}

    def get_obs(self):
        return self._last["obs"].reshape(self.n_envs, self.n_agents, -1)

    def get_obs_vec(self):
        return self._last["vec"]

    def get_state(self):
        return self._last["vec"].reshape(self.n_envs, -1)

    def get_avail_actions(self):
        return self._last["avail"]

    def get_stats(self):
        return self.engine.stats()

    def close(self):
        self.engine.close()

    # ------------------------------------------------------------------ BatchedRunner protocol
    def rollout_spec(self):
        N, F = self.n_agents, self.fov
        return {"obs": ((N, 4 * F * F), torch.uint8), "obs_vec": ((N, 3), torch.float64),
                "avail_actions": ((N, 5), torch.uint8), "reward": ((1,), torch.float64),
                "terminated": ((1,), torch.uint8),
                "_views": {"state": lambda tm: tm["obs_vec"].reshape(tm["obs_vec"].shape[0], self.n_envs, 3 * N)}}

    def reset_into(self, batch):
        self._engine_reset()
        self._t = 0
        self.engine.observe(out={"obs": batch.tm["obs"][0], "vec": batch.tm["obs_vec"][0]})   # straight into the batch
        batch.tm["avail_actions"][0].copy_(self.engine.avail())

    def step_into(self, actions, t, batch):
        tm = batch.tm
        self.engine.step_observe(actions, want=("reward", "terminated", "avail"),
                                 out={"obs": tm["obs"][t + 1], "vec": tm["obs_vec"][t + 1],
                                      "avail": tm["avail_actions"][t + 1], "reward": tm["reward"][t],
                                      "terminated": tm["terminated"][t]})
        self._t = t + 1
        if self._t >= self.episode_limit:
            tm["terminated"][t].fill_(1)
