"""BatchedRunner: pymarl's ParallelRunner loop (MARL-curve-main/src/runners/parallel_runner.py:91-206) over ONE
device-resident vector environment instead of `batch_size_run` worker processes and pipes.

The reference's rollout pays, per env step, a pickle over a Pipe in both directions, an actions GPU->CPU copy and
obs/state/avail numpy->torch copies.  Here the environments live on the GPU: actions are
consumed from the controller without `.cpu()`, and obs / state / avail_actions / reward / terminated are written
into the episode batch as device tensors.

Protocols (the same calls ParallelRunner makes):
    env     : a vector env of this package with n_envs == batch size (MARL_PARTIAL_ENV, MAPF_GRID):
              reset(), step(actions[B, N]) -> (reward[B], terminated[B], info), get_obs(), get_state(),
              get_avail_actions(), get_env_info(), get_stats()
    mac     : init_hidden(batch_size); select_actions(batch, t_ep, t_env, bs, test_mode) -> LongTensor [len(bs), N]
    batch   : new_batch() returns an object with update(data, bs=..., ts=..., mark_filled=...) -- pymarl's
              EpisodeBatch, or DeviceEpisodeBatch below.
"""
import torch


class DeviceEpisodeBatch:
    """Minimal device-resident stand-in for pymarl's EpisodeBatch (components/episode_buffer.py:7-134): one tensor
    [B, T+1, ...] per scheme key with the dtypes of run.py:133-140 (obs/state/reward float32, avail_actions int32,
    actions int64, terminated uint8) plus `filled`."""

    def __init__(self, env_info, batch_size, device):
        B, T, N = batch_size, env_info["episode_limit"] + 1, env_info["n_agents"]
        self.batch_size, self.max_seq_length, self.device = B, T, device
        z = lambda shape, dt: torch.zeros((B, T) + shape, dtype=dt, device=device)  # noqa: E731
        self.data = {
            "state": z((env_info["state_shape"],), torch.float32),
            "obs": z((N, env_info["obs_shape"]), torch.float32),
            "avail_actions": z((N, env_info["n_actions"]), torch.int32),
            "actions": z((N, 1), torch.int64),
            "reward": z((1,), torch.float32),
            "terminated": z((1,), torch.uint8),
            "filled": z((1,), torch.int64),
        }

    def update(self, data, bs=slice(None), ts=slice(None), mark_filled=True):
        if isinstance(bs, list):
            bs = torch.as_tensor(bs, dtype=torch.int64, device=self.device)
        for k, v in data.items():
            dst = self.data[k]
            v = torch.as_tensor(v, device=self.device).to(dst.dtype)
            dst[bs, ts] = v.reshape((-1,) + dst.shape[2:]) if not isinstance(ts, slice) else v
        if mark_filled:
            self.data["filled"][bs, ts] = 1

    def __getitem__(self, k):
        return self.data[k]


class BatchedRunner:
    def __init__(self, env, mac, new_batch=None, test_nepisode=0):
        self.env, self.mac = env, mac
        self.batch_size = env.n_envs
        self.env_info = env.get_env_info()
        self.episode_limit = self.env_info["episode_limit"]
        self.device = env.engine.device
        self.new_batch = new_batch or (lambda: DeviceEpisodeBatch(self.env_info, self.batch_size, self.device))
        self.t = 0
        self.t_env = 0
        self.train_returns, self.test_returns = [], []
        self.train_stats, self.test_stats = {}, {}

    def get_env_info(self):
        return self.env_info

    def reset(self):
        self.batch = self.new_batch()
        self.env.reset()
        self.batch.update({"state": self.env.get_state(), "avail_actions": self.env.get_avail_actions(),
                           "obs": self.env.get_obs()}, ts=0)
        self.t = 0
        self.env_steps_this_run = 0

    def run(self, test_mode=False):
        """One episode in every environment; returns the filled batch (parallel_runner.py:91-206)."""
        self.reset()
        B, dev = self.batch_size, self.device
        episode_returns = torch.zeros(B, dtype=torch.float64, device=dev)
        episode_lengths = torch.zeros(B, dtype=torch.int64, device=dev)
        terminated = torch.zeros(B, dtype=torch.bool, device=dev)
        self.mac.init_hidden(batch_size=B)
        N = self.env_info["n_agents"]
        stay = torch.full((B, N), 4, dtype=torch.int64, device=dev)
        while True:
            alive = (~terminated).nonzero(as_tuple=False).flatten()
            if alive.numel() == 0 or self.t >= self.episode_limit:
                break
            actions = self.mac.select_actions(self.batch, t_ep=self.t, t_env=self.t_env, bs=alive.tolist(),
                                              test_mode=test_mode)
            actions = actions.to(dev).reshape(alive.numel(), N)
            self.batch.update({"actions": actions.unsqueeze(-1)}, bs=alive, ts=self.t, mark_filled=False)
            full = stay.clone()
            full[alive] = actions                                  # finished environments idle; their data is masked
            reward, term, info = self.env.step(full)
            term = term.bool()
            episode_returns[alive] += reward[alive].double()
            episode_lengths[alive] += 1
            if not test_mode:
                self.env_steps_this_run += int(alive.numel())
            # parallel_runner.py:150-153: a termination is recorded unless the env flags it as `episode_limit` in its
            # info; the MAPF envs of the reference never set that key, so limit terminations are recorded too
            limit_flag = info.get("episode_limit") if isinstance(info, dict) else None
            rec = term if limit_flag is None else (term & ~torch.as_tensor(limit_flag, device=dev).bool())
            self.batch.update({"reward": reward[alive].float().unsqueeze(-1),
                               "terminated": rec[alive].to(torch.uint8).unsqueeze(-1)},
                              bs=alive, ts=self.t, mark_filled=False)
            terminated = terminated | term
            self.t += 1
            self.batch.update({"state": self.env.get_state()[alive], "avail_actions": self.env.get_avail_actions()[alive],
                               "obs": self.env.get_obs()[alive]}, bs=alive, ts=self.t, mark_filled=True)
        if not test_mode:
            self.t_env += self.env_steps_this_run
        stats = self.test_stats if test_mode else self.train_stats
        stats["n_episodes"] = B + stats.get("n_episodes", 0)
        stats["ep_length"] = int(episode_lengths.sum().item()) + stats.get("ep_length", 0)
        for k, v in self.env.get_stats().items():
            stats["env_" + k] = v
        (self.test_returns if test_mode else self.train_returns).extend(episode_returns.tolist())
        return self.batch
