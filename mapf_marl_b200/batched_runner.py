"""BatchedRunner: pymarl's ParallelRunner loop (MARL-curve-main/src/runners/parallel_runner.py:91-206) over ONE
device-resident vector environment instead of `batch_size_run` worker processes and pipes.

The reference's rollout pays, per env step, a pickle over a Pipe in both directions, an actions GPU->CPU copy and
obs/state/avail numpy->torch copies.  Here the environments live on the GPU and the loop never
touches the host:

  * the episode batch is TIME-MAJOR device storage ([T+1, B, ...] per key, DeviceEpisodeBatch) and the fused kernel
    writes obs / avail_actions / state / reward / terminated of a step straight into its time slices (the `out=`
    argument of MapfEngine.step_observe): no staging tensors, no copies;
  * one engine launch per environment step for the PRIMAL and GRID envs (two for PARTIAL, whose observation is its own
    kernel; its get_state() triple rides along with it);
  * no `.item()`, `.tolist()`, `.nonzero()` in the loop: finished environments keep stepping with the STAY action and
    their entries are masked by `filled`, exactly the mask pymarl's learners apply; whether every environment has
    finished is checked every `check_every` steps (ONE host sync per check);
  * `cuda_graph=True`: the loop body of every block of `check_every` steps -- controller, masking, the engine launch,
    bookkeeping -- is captured once into a CUDA graph and replayed, so an environment step costs no Python or launch
    latency at all (controllers must then keep their per-episode state in device tensors they update in place;
    RandomMAC's draws are a function of (seed, step) and therefore repeat from replay to replay).

Protocols:
    env : a vector env of this package (MARL_PARTIAL_ENV / MAPF_GRID with n_envs > 1, PrimalVecEnv) providing
          rollout_spec(), reset_into(batch), step_into(actions, t, batch), stay_action, get_env_info(), get_stats()
    mac : init_hidden(batch_size); select_actions(batch, t_ep, t_env, bs, test_mode) -> integer tensor [B, N] on the
          device (bs is always slice(None): the whole batch, like pymarl's controllers accept)
"""
import torch


class DeviceEpisodeBatch:
    """Device-resident stand-in for pymarl's EpisodeBatch (components/episode_buffer.py:7-134).

    Storage is time-major, `tm[key]` = [T+1, B, ...], so that one time step of one key is a contiguous block a kernel
    can write; `batch[key]` is the [B, T+1, ...] view pymarl code indexes (`batch["avail_actions"][:, t]`).  Keys and
    per-step shapes come from the env's rollout_spec(); dtypes are the kernel's own (uint8 masks / flags, float64
    rewards, the env's observation dtype) -- the values pymarl's scheme (run.py:133-148) would hold, before its casts.
    Entries of (env, t) pairs with filled == 0 are undefined (pymarl leaves zeros there); consumers mask with `filled`.
    """

    def __init__(self, spec, batch_size, max_seq_length, n_agents, device):
        self.batch_size, self.max_seq_length, self.device = batch_size, max_seq_length, device
        self.views = dict(spec.get("_views", {}))
        self.tm = {}
        for k, v in spec.items():
            if k != "_views":
                shape, dt = v
                self.tm[k] = torch.zeros((max_seq_length, batch_size) + tuple(shape), dtype=dt, device=device)
        self.tm["actions"] = torch.zeros((max_seq_length, batch_size, n_agents, 1), dtype=torch.int64, device=device)
        self.tm["filled"] = torch.zeros((max_seq_length, batch_size, 1), dtype=torch.uint8, device=device)

    def __getitem__(self, k):
        if k in self.views:
            return self.views[k](self.tm).transpose(0, 1)
        return self.tm[k].transpose(0, 1)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tm.values())


class BatchedRunner:
    def __init__(self, env, mac, check_every=8, max_steps=None, cuda_graph=False, fused_bookkeeping=True):
        self.env, self.mac = env, mac
        self.fused_bookkeeping = bool(fused_bookkeeping)
        self.cuda_graph = bool(cuda_graph)
        self._graphs = None
        self._reset_graph = None
        self.batch_size = env.n_envs
        self.env_info = env.get_env_info()
        self.episode_limit = self.env_info["episode_limit"]
        self.max_steps = min(self.episode_limit, max_steps) if max_steps else self.episode_limit
        self.check_every = max(int(check_every), 1)
        self.device = env.engine.device
        self.t = 0
        self._t_env = 0
        self._returns = {False: [], True: []}     # per-episode return tensors (device); lists are built on demand
        self._stats = {False: {}, True: {}}
        self._pending = []                        # (test_mode, device scalar: env steps of a finished run()), unresolved
        self._env_stats_stale = {False: False, True: False}
        self.batch = None

    # What run() learns about an episode only on the device -- how many environment steps it took, the engine's
    # counters -- is resolved when somebody reads it, not at the end of run(): two host synchronisations per episode
    # kept the GPU idle while the next episode's launches were still being issued (14 % of a 16-step c3 episode).
    def _resolve(self):
        if self._pending:
            for test_mode, steps in self._pending:
                n = int(steps.item())
                if not test_mode:
                    self._t_env += n
                st = self._stats[test_mode]
                st["ep_length"] = n + st.get("ep_length", 0)
            self._pending = []

    @property
    def t_env(self):
        self._resolve()
        return self._t_env

    @t_env.setter
    def t_env(self, v):
        self._resolve()
        self._t_env = int(v)

    def _stats_view(self, test_mode):
        self._resolve()
        st = self._stats[test_mode]
        if self._env_stats_stale[test_mode]:
            for k, v in self.env.get_stats().items():
                st["env_" + k] = v
            self._env_stats_stale[test_mode] = False
        return st

    @property
    def train_stats(self):
        return self._stats_view(False)

    @property
    def test_stats(self):
        return self._stats_view(True)

    def get_env_info(self):
        return self.env_info

    @property
    def train_returns(self):
        """Episode returns of the training runs as a Python list (pymarl logs their mean, parallel_runner.py:196-204);
        kept as device tensors until somebody asks: a million-environment batch must not pay a host list per episode."""
        return torch.cat(self._returns[False]).tolist() if self._returns[False] else []

    @property
    def test_returns(self):
        return torch.cat(self._returns[True]).tolist() if self._returns[True] else []

    def new_batch(self):
        return DeviceEpisodeBatch(self.env.rollout_spec(), self.batch_size, self.max_steps + 1,
                                  self.env_info["n_agents"], self.device)

    def reset(self, reuse_batch=False):
        if not (reuse_batch and self.batch is not None):
            self.batch = self.new_batch()
        else:
            self.batch.tm["filled"].zero_()
        self.env.reset_into(self.batch)                      # obs / avail_actions / state of t = 0
        self.batch.tm["filled"][0] = 1
        self.t = 0

    def _step_body(self, t, st, test_mode):
        """Environment step t of the episode; every tensor it touches is persistent (graph-capturable)."""
        B, batch = self.batch_size, self.batch
        alive = st["alive"]
        actions = self.mac.select_actions(batch, t_ep=t, t_env=self._t_env, bs=slice(None), test_mode=test_mode)
        actions = actions.reshape(B, -1)
        eng = getattr(self.env, "engine", None)
        if (self.fused_bookkeeping and eng is not None and actions.dtype in (torch.uint8, torch.int64)
                and batch.tm["reward"].dtype == torch.float64 and batch.tm["terminated"].dtype == torch.uint8):
            # the same bookkeeping as below in two library kernels instead of a dozen element-wise launches
            # (36 us of a 89 us environment step at c3)
            if "act_u8" not in st:
                st["act_u8"] = torch.empty((B, actions.shape[1]), dtype=torch.uint8, device=self.device)
            eng.runner_mask_actions(actions.contiguous(), alive, self.env.stay_action, st["act_u8"],
                                    batch.tm["actions"][t])
            self.env.step_into(st["act_u8"], t, batch)
            eng.runner_account(batch.tm["reward"][t], batch.tm["terminated"][t], alive, st["returns"], st["lengths"],
                               batch.tm["filled"][t + 1])
            return
        # finished environments idle; their entries stay masked (filled == 0)
        actions = torch.where(alive.unsqueeze(1), actions, torch.full_like(actions, self.env.stay_action))
        batch.tm["actions"][t] = actions.unsqueeze(-1)
        self.env.step_into(actions, t, batch)                # reward / terminated -> t ; obs / avail / state -> t + 1
        # parallel_runner.py:150-153: a termination is recorded unless the env flags it as `episode_limit` in its
        # info; the MAPF envs of the reference never set that key, so limit terminations are recorded too
        term = batch.tm["terminated"][t].reshape(B) != 0
        st["returns"] += torch.where(alive, batch.tm["reward"][t].reshape(B).double(), st["zero"])
        st["lengths"] += alive
        batch.tm["filled"][t + 1] = alive.unsqueeze(-1)      # the step's data exists for envs that were still running
        alive.logical_and_(~term)

    def run(self, test_mode=False, reuse_batch=False):
        """One episode in every environment; returns the filled batch (parallel_runner.py:91-206)."""
        reuse_batch = reuse_batch or self.cuda_graph         # captured graphs hold the batch's addresses
        if self._graphs is None:
            self._resolve()                                  # eager steps (and a capture) read t_env on the host
        if self._graphs is not None and self._reset_graph is not None:
            self._reset_graph.replay()                       # the same launches as reset(True), without the host side
            self.t = 0
        else:
            self.reset(reuse_batch)
        B, dev = self.batch_size, self.device
        if getattr(self, "_st", None) is None:
            self._st = {"returns": torch.zeros(B, dtype=torch.float64, device=dev),
                        "lengths": torch.zeros(B, dtype=torch.int64, device=dev),
                        "alive": torch.ones(B, dtype=torch.bool, device=dev),
                        "zero": torch.zeros(B, dtype=torch.float64, device=dev)}
        st = self._st
        st["returns"].zero_()
        st["lengths"].zero_()
        st["alive"].fill_(True)
        self.mac.init_hidden(batch_size=B)
        C, T = self.check_every, self.max_steps
        blocks = [range(t0, min(t0 + C, T)) for t0 in range(0, T, C)]
        if self.cuda_graph and self._graphs is None and getattr(self, "_warm", False):
            # capture after one eager episode (library handles, controller buffers and the batch all exist by now)
            self._graphs = []
            pool = None
            torch.cuda.synchronize()
            if getattr(self.env, "graph_safe_reset", False):
                # the episode reset as a graph of its own (envs whose reset_into touches nothing but device tensors)
                g0 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g0):
                    self.reset(True)
                self._reset_graph, pool = g0, g0.pool()
            for blk in blocks:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    for t in blk:
                        self._step_body(t, st, test_mode)
                pool = g.pool()
                self._graphs.append(g)
            # the capture pass does not execute anything: restore the start-of-episode state and replay below
            st["returns"].zero_()
            st["lengths"].zero_()
            st["alive"].fill_(True)
            self.mac.init_hidden(batch_size=B)
        for k, blk in enumerate(blocks):
            if self._graphs is not None:
                self._graphs[k].replay()
            else:
                for t in blk:
                    self._step_body(t, st, test_mode)
            self.t = blk[-1] + 1
            if self.t < T and not bool(st["alive"].any()):
                break                                        # the only host synchronisation of the loop
        self._warm = True
        test_mode = bool(test_mode)
        self._pending.append((test_mode, st["lengths"].sum()))        # a device scalar: resolved when somebody asks
        stats = self._stats[test_mode]
        stats["n_episodes"] = B + stats.get("n_episodes", 0)
        self._env_stats_stale[test_mode] = True
        self._returns[test_mode].append(st["returns"].clone())
        return self.batch


# ----------------------------------------------------------------------------------------------------------------
# Controllers for benchmarks and tests (the learner side of pymarl is out of scope; these only produce actions)
# ----------------------------------------------------------------------------------------------------------------
class RandomMAC:
    """Uniform over the available actions, drawn on the device by the engine's counter-hash kernel
    (mapf_random_actions): one launch, no host round trip, a function of (seed, episode, step, env, agent)."""

    def __init__(self, engine, seed=0, env_offset=0):
        self.engine, self.seed, self.env_offset, self.episode = engine, int(seed), int(env_offset), -1

    def init_hidden(self, batch_size):
        self.episode += 1

    def select_actions(self, batch, t_ep, t_env, bs=slice(None), test_mode=False):
        return self.engine.random_actions(self.seed + 7919 * self.episode, t_ep, avail=batch.tm["avail_actions"][t_ep],
                                          env_offset=self.env_offset)


class RNNAgentMAC:
    """pymarl's recurrent agent (MARL-curve-main/src/modules/agents/rnn_agent.py:12-21: Linear -> GRUCell -> Linear,
    shared by all agents) with greedy masked action selection, in plain torch: random-init weights, bf16 first layer.
    The network is a caller of the environment path, not part of it."""

    def __init__(self, obs_dim, n_actions, device, hidden=64, extra_dim=0, seed=0):
        g = torch.Generator(device="cpu").manual_seed(seed)
        mk = lambda *s: (torch.randn(*s, generator=g) * 0.05).to(device)   # noqa: E731
        self.w1 = mk(obs_dim, hidden).to(torch.bfloat16)
        self.w1x = mk(extra_dim, hidden) if extra_dim else None
        self.b1 = mk(hidden)
        self.gru = torch.nn.GRUCell(hidden, hidden).to(device)
        self.w2, self.b2 = mk(hidden, n_actions), mk(n_actions)
        self.hidden = hidden
        self.h = None

    def init_hidden(self, batch_size):
        if self.h is not None:
            self.h.zero_()                  # in place: a captured rollout graph keeps reading this tensor

    @torch.no_grad()
    def select_actions(self, batch, t_ep, t_env, bs=slice(None), test_mode=False):
        obs = batch.tm["obs"][t_ep]                                       # [B, N, D]
        B, N = obs.shape[0], obs.shape[1]
        x = obs.reshape(B * N, -1).to(torch.bfloat16) @ self.w1
        x = x.float() + self.b1
        if self.w1x is not None:
            x = x + batch.tm["obs_vec"][t_ep].reshape(B * N, -1).float() @ self.w1x
        x = torch.relu(x)
        if self.h is None:
            self.h = torch.zeros(B * N, self.hidden, device=x.device)
        self.h.copy_(self.gru(x, self.h))
        q = self.h @ self.w2 + self.b2
        q = q.masked_fill(batch.tm["avail_actions"][t_ep].reshape(B * N, -1) == 0, float("-inf"))
        return q.argmax(-1).reshape(B, N)
