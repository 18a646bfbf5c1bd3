"""MapfEngine: thousands of MAPF grid-world environments resident on one B200.

Thin host-side wrapper over the C ABI (include/mapf_b200.h): torch is used only for device
memory, streams and dtype plumbing.  All arithmetic happens in csrc/mapf_kernels.cu.

Ownership of outputs: step / step_observe / observe / avail / positions / ... return ENGINE-OWNED tensors that the
next call of the same method overwrites in place (no allocation on the hot path).  Keep a value across calls with
`.clone()`, or hand the call your own storage (`out=` of step_observe / observe, `rollout()`).

Vector API (device tensors, no host round trips):
    eng = MapfEngine(n_envs=E, n_agents=N, height=H, width=W, mode="primal", fov=11)
    eng.reset(obst, starts, goals)                       # obst [E,H,W] (or [H,W] shared), starts/goals [E,N,2]
    out = eng.step_observe(actions)                      # actions uint8/int64 [E,N] on the device
    out["obs"] [E,N,4,F,F] u8, out["vec"] [E,N,3] f64, out["reward"], out["terminated"], out["avail"], ...

Semantics follow the reference envs (citations in include/mapf_b200.h):
    mode "grid"   = mapf_gridworld.MAPF_GRID.step (detect-and-penalise collisions)
    mode "primal" = mapf_primal.MAPFEnv._step swept over agents 1..N (sequential claim)
"""
import ctypes
import sys

import numpy as np
import torch

from . import _lib
from ._lib import (BITS, F32, F64, I8, I64, MODE_GRID, MODE_PARTIAL, MODE_PRIMAL, OBS_FULLMAP, OBS_PARTIAL_WINDOW,
                   OBS_PRIMAL_FOV, STAT_NAMES, STEP_OUT_FIELDS, U8, MapfCfg, MapfHostIO, MapfStepOut)

_OUT_SPECS = {
    # name: (dtype, per-env shape suffix builder)
    "reward": (torch.float64, lambda N: ()),
    "terminated": (torch.uint8, lambda N: ()),
    "agent_reward": (torch.float64, lambda N: (N,)),
    "dones": (torch.uint8, lambda N: (N,)),
    "status": (torch.int8, lambda N: (N,)),
    "node": (torch.int16, lambda N: (N,)),
    "edge": (torch.int16, lambda N: (N,)),
    "valid": (torch.uint8, lambda N: (N,)),
    "done_mid": (torch.uint8, lambda N: (N,)),
    "next_mid": (torch.uint8, lambda N: (N, -1)),   # -1: the engine's n_actions (5, or 9 with diagonal movement)
    "avail": (torch.uint8, lambda N: (N, -1)),
    "blocking": (torch.uint8, lambda N: (N,)),
}

DEFAULT_WANT = ("reward", "terminated", "dones", "avail")


class MapfError(RuntimeError):
    pass


def python_sum_mode():
    """1 when this interpreter's builtin sum() compensates float sums (CPython >= 3.12), else 0."""
    return 1 if sys.version_info >= (3, 12) else 0


def magnitude_lut(height, width):
    """mag[s] for s = dx*dx + dy*dy, evaluated with the reference's expression (mapf_primal.py:382)."""
    n = (height - 1) ** 2 + (width - 1) ** 2 + 1
    return np.array([s ** .5 for s in range(n)], dtype=np.float64)


class MapfEngine:
    def __init__(self, n_envs, n_agents, height, width, mode="primal", obs_mode=None, fov=11, shared_map=False,
                 episode_limit=10000, step_reward=-0.01, collide_reward=-10, action_cost=-0.3, idle_cost=-0.5,
                 goal_reward=0.0, collision_reward=-2.0, goal_dist=False, collect_stats=True, device=None,
                 reward_sum_mode=None, obs_window=5, obs_knn_agents=5, move_reward=-0.01, stay_reward=-0.02,
                 stay_goal_reward=0, node_collide_reward=-1, edge_collide_reward=-1, env_collide_reward=-1,
                 complete_reward=1000, complete_fac=1.5, gamma=0.99, blocking_reward=False, blocking_cost=-1.0,
                 diagonal_movement=False):
        if not torch.cuda.is_available():
            raise MapfError("MapfEngine needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        if self.device.type != "cuda":
            raise MapfError("MapfEngine needs a CUDA device; got %s" % self.device)
        self.E, self.N, self.H, self.W = int(n_envs), int(n_agents), int(height), int(width)
        self.mode = ({"grid": MODE_GRID, "primal": MODE_PRIMAL, "partial": MODE_PARTIAL}[mode]
                     if isinstance(mode, str) else int(mode))
        if obs_mode is None:
            obs_mode = {MODE_PRIMAL: OBS_PRIMAL_FOV, MODE_PARTIAL: OBS_PARTIAL_WINDOW}.get(self.mode, OBS_FULLMAP)
        elif isinstance(obs_mode, str):
            obs_mode = {"fullmap": OBS_FULLMAP, "fov": OBS_PRIMAL_FOV, "window": OBS_PARTIAL_WINDOW}[obs_mode]
        self.obs_mode = int(obs_mode)
        self.F = int(fov)
        self.shared_map = bool(shared_map)
        self.has_goal_dist = bool(goal_dist) or self.mode == MODE_PARTIAL
        self.obs_window, self.obs_knn_agents = int(obs_window), int(obs_knn_agents)
        self.obs_size = 2 * self.obs_window ** 2 + 13 * self.obs_knn_agents
        cfg = MapfCfg()
        self.lib.mapf_default_cfg(ctypes.byref(cfg))
        cfg.n_envs, cfg.n_agents, cfg.height, cfg.width = self.E, self.N, self.H, self.W
        cfg.mode, cfg.obs_mode, cfg.fov = self.mode, self.obs_mode, self.F
        cfg.shared_map = int(self.shared_map)
        cfg.episode_limit = int(episode_limit)
        cfg.goal_dist = int(self.has_goal_dist)
        cfg.collect_stats = int(bool(collect_stats))
        cfg.step_reward, cfg.collide_reward = float(step_reward), float(collide_reward)
        cfg.action_cost, cfg.idle_cost = float(action_cost), float(idle_cost)
        cfg.goal_reward, cfg.collision_reward = float(goal_reward), float(collision_reward)
        cfg.reward_sum_mode = python_sum_mode() if reward_sum_mode is None else int(reward_sum_mode)
        cfg.step_reward_is_int = int(isinstance(step_reward, int))
        cfg.collide_reward_is_int = int(isinstance(collide_reward, int))
        self._lut = magnitude_lut(self.H, self.W)
        cfg.mag_lut_host = self._lut.ctypes.data
        cfg.mag_lut_len = int(self._lut.size)
        cfg.blocking_reward = int(bool(blocking_reward))
        cfg.blocking_cost = float(blocking_cost)
        cfg.diagonal_movement = int(bool(diagonal_movement))   # DIAGONAL_MOVEMENT, mapf_primal.py:175
        self.n_actions = 9 if diagonal_movement else 5
        if self.mode == MODE_PARTIAL:
            cfg.obs_window, cfg.obs_knn_agents = self.obs_window, self.obs_knn_agents
            cfg.move_reward, cfg.stay_reward = float(move_reward), float(stay_reward)
            cfg.stay_goal_reward = float(stay_goal_reward)
            cfg.node_collide_reward, cfg.edge_collide_reward = float(node_collide_reward), float(edge_collide_reward)
            cfg.env_collide_reward = float(env_collide_reward)
            # the completion bonus per step index, with the reference's own expression (marl_partial.py:296)
            limit = int(episode_limit)
            def bonus(t):
                try:
                    return (complete_reward / (gamma ** (limit - t))) * complete_fac
                except (ZeroDivisionError, OverflowError):
                    # gamma ** (limit - t) underflowed / overflowed: the reference raises at this step count
                    return float("inf")
            self._clut = np.array([bonus(t) for t in range(limit + 65)], dtype=np.float64)
            cfg.complete_lut_host = self._clut.ctypes.data
            cfg.complete_lut_len = int(self._clut.size)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.mapf_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != 0:
            self._h = None
            raise MapfError("mapf_create failed (%d): %s" % (rc, _lib.last_error(self.lib)))
        self._bufs = {}
        self._keep = []   # tensors referenced by asynchronous launches

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self.lib.mapf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc, what):
        if rc != 0:
            raise MapfError("%s failed (%d): %s" % (what, rc, _lib.last_error(self.lib, self._h)))

    def _buf(self, name, shape, dtype):
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    def _to_dev(self, x, dtype, shape=None):
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            t = x.to(device=self.device, dtype=dtype).contiguous()
        else:
            t = torch.as_tensor(np.array(x, copy=True), device=self.device).to(dtype).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError("expected shape %s, got %s" % (tuple(shape), tuple(t.shape)))
        return t

    @staticmethod
    def _ptr(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _actions(self, actions):
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.dtype not in (torch.uint8, torch.int64):
            actions = actions.to(torch.int64)
        a = actions.to(self.device).contiguous()
        if tuple(a.shape) != (self.E, self.N):
            raise ValueError("actions must have shape (%d, %d), got %s" % (self.E, self.N, tuple(a.shape)))
        return a, (U8 if a.dtype == torch.uint8 else I64)

    def _step_out(self, want, out=None, T=None):
        """The mapf_step_out descriptor for the outputs in `want`.  out: optional dict of CALLER-OWNED tensors (e.g.
        slices of an episode batch) the kernel writes into directly; T: leading time dimension of a rollout."""
        so = MapfStepOut()
        outs = {}
        lead = (self.E,) if T is None else (int(T), self.E)
        for name in want:
            if name not in _OUT_SPECS:
                raise KeyError("unknown step output %r (choose from %s)" % (name, ", ".join(STEP_OUT_FIELDS)))
            dtype, suffix = _OUT_SPECS[name]
            shape = lead + tuple(self.n_actions if k == -1 else k for k in suffix(self.N))
            t = out.get(name) if out else None
            if t is None:
                t = self._buf(("out_" if T is None else "roll_") + name, shape, dtype)
            else:
                self._check_out(name, t, shape, dtype)
            setattr(so, name + "_dev", t.data_ptr())
            outs[name] = t
        return so, outs

    def _check_out(self, name, t, shape, dtype):
        if (not isinstance(t, torch.Tensor) or t.device != self.device or t.dtype != dtype or not t.is_contiguous()
                or t.numel() != int(np.prod(shape))):
            raise ValueError("out[%r] must be a contiguous %s tensor with %d elements on %s" %
                             (name, dtype, int(np.prod(shape)), self.device))

    # ------------------------------------------------------------------ state
    def reset(self, obst=None, starts=None, goals=None, env_mask=None):
        """obst: non-zero = obstacle, [E,H,W] or [H,W] when shared_map.  starts/goals: [E,N,2] (p0, p1).
        Any argument may be None to keep the stored value (MAPF_GRID.reset re-uses its start positions)."""
        mshape = (self.H, self.W) if self.shared_map else (self.E, self.H, self.W)
        m = self._to_dev(obst, torch.int8, mshape)
        s = self._to_dev(starts, torch.int16, (self.E, self.N, 2))
        g = self._to_dev(goals, torch.int16, (self.E, self.N, 2))
        k = self._to_dev(env_mask, torch.uint8, (self.E,))
        self._keep = [m, s, g, k]
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_reset(self._h, self._ptr(m), self._ptr(s), self._ptr(g), self._ptr(k),
                                            self._stream()), "mapf_reset")

    def set_goals(self, goals, dirty=None):
        g = self._to_dev(goals, torch.int16, (self.E, self.N, 2))
        dmask = self._to_dev(dirty, torch.uint8, (self.E, self.N))
        self._keep = [g, dmask]
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_set_goals(self._h, self._ptr(g), self._ptr(dmask), self._stream()),
                        "mapf_set_goals")

    def pop_goals(self, queue, head, dirty_out=None):
        """Lifelong hand-out (MAPF-490-main/Global.cpp:85-94): agents standing on their goal take queue[head] as
        their new goal.  queue int16 [E,N,Q,2], head int32 [E,N] (advanced in place); returns the uint8 [E,N] mask of
        re-assigned agents."""
        assert queue.dtype == torch.int16 and queue.is_cuda and queue.is_contiguous()
        assert head.dtype == torch.int32 and head.is_cuda and head.is_contiguous()
        assert tuple(queue.shape[:2]) == (self.E, self.N) and queue.shape[3] == 2 and tuple(head.shape) == (self.E, self.N)
        if dirty_out is None:
            dirty_out = self._buf("pop_dirty", (self.E, self.N), torch.uint8)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_pop_goals(self._h, self._ptr(queue), self._ptr(head), int(queue.shape[2]),
                                                self._ptr(dirty_out), self._stream()), "mapf_pop_goals")
        return dirty_out

    def lifelong_bind(self, queue, head):
        """Fuses the lifelong hand-out into the step: from now on every full-range step launch pops the queue of every
        agent that ends the step on its goal (exactly pop_goals() behind the step, without the launch).  queue int16
        [E,N,Q,2], head int32 [E,N], both kept alive and read by every step; queue=None unbinds."""
        if queue is None:
            self._life = None
            self._check(self.lib.mapf_lifelong_bind(self._h, None, None, 0), "mapf_lifelong_bind")
            return
        assert queue.dtype == torch.int16 and queue.is_cuda and queue.is_contiguous()
        assert head.dtype == torch.int32 and head.is_cuda and head.is_contiguous()
        assert tuple(queue.shape[:2]) == (self.E, self.N) and queue.shape[3] == 2 and tuple(head.shape) == (self.E, self.N)
        self._life = (queue, head)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_lifelong_bind(self._h, self._ptr(queue), self._ptr(head), int(queue.shape[2])),
                        "mapf_lifelong_bind")

    def bfs_popped(self, out=None):
        """Goal-distance maps of the agents whose goals the most recent step re-assigned (lifelong_bind), into `out`
        (int16 [E,N,H,W]) or the handle's own maps."""
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_bfs_popped(self._h, self._ptr(out), self._stream()), "mapf_bfs_popped")

    def set_prev_actions(self, prev):
        p = self._to_dev(prev, torch.uint8, (self.E, self.N))
        self._keep = [p]
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_set_prev_actions(self._h, self._ptr(p), self._stream()), "mapf_set_prev_actions")

    def positions(self):
        out = self._buf("positions", (self.E, self.N, 2), torch.int16)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_get_positions(self._h, self._ptr(out), self._stream()), "mapf_get_positions")
        return out

    def goals(self):
        out = self._buf("goals", (self.E, self.N, 2), torch.int16)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_get_goals(self._h, self._ptr(out), self._stream()), "mapf_get_goals")
        return out

    def dones(self):
        out = self._buf("dones", (self.E, self.N), torch.uint8)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_get_dones(self._h, self._ptr(out), self._stream()), "mapf_get_dones")
        return out

    def step_count(self):
        out = self._buf("step_count", (self.E,), torch.int32)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_get_step_count(self._h, self._ptr(out), self._stream()), "mapf_get_step_count")
        return out

    # ------------------------------------------------------------------ hot path
    def step(self, actions, want=DEFAULT_WANT, agent_range=None):
        """One environment step (GRID) / one sweep over agents (PRIMAL).  Returns a dict of device tensors."""
        a, adt = self._actions(actions)
        so, outs = self._step_out(want)
        self._keep = [a]
        with torch.cuda.device(self.device):
            if agent_range is None:
                rc = self.lib.mapf_step(self._h, self._ptr(a), adt, ctypes.byref(so), self._stream())
            else:
                rc = self.lib.mapf_step_agents(self._h, self._ptr(a), adt, int(agent_range[0]), int(agent_range[1]),
                                               ctypes.byref(so), self._stream())
        self._check(rc, "mapf_step")
        return outs

    def _obs_buffers(self, dtype, want_vec, out=None, T=None):
        have_all = bool(out) and out.get("obs") is not None and (not want_vec or out.get("vec") is not None
                                                                  or self.obs_mode != OBS_PRIMAL_FOV)
        # caller-owned storage for everything: only the shapes are needed (no engine buffer is allocated)
        mk = (lambda name, shape, dt: torch.empty(shape, dtype=dt, device="meta")) if have_all else self._buf
        obs, vec, odt = self._obs_buffers_own(dtype, want_vec, T, mk)
        if out:
            if out.get("obs") is not None:
                self._check_out("obs", out["obs"], tuple(obs.shape), obs.dtype)
                obs = out["obs"]
            if vec is not None and out.get("vec") is not None:
                self._check_out("vec", out["vec"], tuple(vec.shape), vec.dtype)
                vec = out["vec"]
        return obs, vec, odt

    def _obs_buffers_own(self, dtype, want_vec, T=None, mk=None):
        mk = mk or self._buf
        if T is not None:
            # time-major rollout storage: one engine-owned tensor per (dtype, T)
            lead = (int(T),)
            if self.obs_mode == OBS_PRIMAL_FOV:
                vec = mk("roll_vec", lead + (self.E, self.N, 3), torch.float64) if want_vec else None
                if dtype == "bits":
                    cells = self.E * self.N * 4 * self.F * self.F
                    if cells % 32:
                        raise ValueError("bit-packed rollouts need E*N*4*F*F to be a multiple of 32")
                    return mk("roll_obs_bits", lead + (cells // 8,), torch.uint8), vec, BITS
                if dtype not in (torch.uint8, torch.float32):
                    raise ValueError("FOV observations are uint8, float32 or 'bits'")
                return (mk("roll_obs_%s" % dtype, lead + (self.E, self.N, 4, self.F, self.F), dtype), vec,
                        U8 if dtype == torch.uint8 else F32)
            if self.obs_mode == OBS_PARTIAL_WINDOW:
                pdt = torch.float32 if dtype == torch.float32 else torch.float64
                return (mk("roll_obs_partial_%s" % pdt, lead + (self.E, self.N, self.obs_size), pdt), None,
                        F32 if pdt == torch.float32 else F64)
            return mk("roll_state", lead + (self.E, self.H * self.W), torch.int8), None, I8
        if self.obs_mode == OBS_PRIMAL_FOV:
            if dtype == "bits":
                # one bit per cell, bit i of the stream == element i of the uint8 tensor (little-endian bit order):
                # np.unpackbits(obs.cpu().numpy(), bitorder="little")[:E*N*4*F*F] gives the cells back
                nwords = (self.E * self.N * 4 * self.F * self.F + 31) // 32
                obs = mk("obs_bits", (nwords * 4,), torch.uint8)
                vec = mk("vec", (self.E, self.N, 3), torch.float64) if want_vec else None
                return obs, vec, BITS
            if dtype not in (torch.uint8, torch.float32):
                raise ValueError("FOV observations are uint8, float32 or 'bits'")
            obs = mk("obs_%s" % dtype, (self.E, self.N, 4, self.F, self.F), dtype)
            vec = mk("vec", (self.E, self.N, 3), torch.float64) if want_vec else None
            return obs, vec, (U8 if dtype == torch.uint8 else F32)
        if self.obs_mode == OBS_PARTIAL_WINDOW:
            # float64 like the reference's get_obs; float32 = the same values rounded once at the store (what
            # pymarl's episode batch keeps, src/run.py:133-140)
            if dtype == torch.float32:
                return mk("obs_partial32", (self.E, self.N, self.obs_size), torch.float32), None, F32
            return mk("obs_partial", (self.E, self.N, self.obs_size), torch.float64), None, F64
        obs = mk("state", (self.E, self.H * self.W), torch.int8)
        return obs, None, I8

    def observe(self, dtype=torch.uint8, want_vec=True, out=None):
        """FOV: (obs [E,N,4,F,F], vec [E,N,3]); full map: (state int8 [E,H*W], None).  out: optional dict with
        caller-owned "obs" / "vec" tensors to write into (e.g. time slice 0 of an episode batch)."""
        obs, vec, odt = self._obs_buffers(dtype, want_vec, out)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_observe(self._h, self._ptr(obs), odt, self._ptr(vec), self._stream()),
                        "mapf_observe")
        return obs, vec

    def step_observe(self, actions, want=DEFAULT_WANT, dtype=torch.uint8, want_vec=True, out=None):
        """Fused step + observation: one kernel launch.  out: optional dict name -> caller-owned contiguous tensor
        ("obs", "vec" and any step output) that the kernel writes into directly, e.g. the time slice t of a
        time-major episode batch; everything else lands in engine-owned tensors (valid until the next call)."""
        a, adt = self._actions(actions)
        so, outs = self._step_out(want, out)
        obs, vec, odt = self._obs_buffers(dtype, want_vec, out)
        self._keep = [a]
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_step_observe(self._h, self._ptr(a), adt, ctypes.byref(so), self._ptr(obs), odt,
                                                   self._ptr(vec), self._stream()), "mapf_step_observe")
        outs["obs"] = obs
        if vec is not None:
            outs["vec"] = vec
        return outs

    def rollout(self, actions, want=DEFAULT_WANT, dtype=torch.uint8, want_vec=True, out=None, observe=True):
        """T consecutive fused steps with pre-supplied actions [T,E,N] in one call (mapf_rollout) -- one kernel launch
        for PRIMAL / GRID batches whose tiles hold one thread per agent (rollout_in_one_launch()).  Returns
        time-major tensors: out["obs"] [T,E,N,4,F,F], out["reward"] [T,E], ...; step t's entries equal what the t-th
        of T consecutive step_observe() calls returns."""
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.dtype not in (torch.uint8, torch.int64):
            actions = actions.to(torch.int64)
        a = actions.to(self.device).contiguous()
        if a.dim() != 3 or tuple(a.shape[1:]) != (self.E, self.N):
            raise ValueError("actions must have shape (T, %d, %d), got %s" % (self.E, self.N, tuple(a.shape)))
        T = int(a.shape[0])
        so, outs = self._step_out(want, out, T)
        obs = vec = None
        odt = U8
        if observe:
            obs, vec, odt = self._obs_buffers(dtype, want_vec, out, T)
        self._keep = [a]
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_rollout(self._h, self._ptr(a), U8 if a.dtype == torch.uint8 else I64, T,
                                              ctypes.byref(so), self._ptr(obs), odt, self._ptr(vec), self._stream()),
                        "mapf_rollout")
        if obs is not None:
            outs["obs"] = obs
        if vec is not None:
            outs["vec"] = vec
        return outs

    def random_actions(self, seed, step, avail=None, env_offset=0, out=None, dtype=torch.int64):
        """Counter-hash random policy on the device (mapf_random_actions): uniform over the actions, or over the set
        bits of `avail` (uint8 [E,N,A]).  Same values as workloads.hash_actions_np(seed, env_offset + arange(E), step,
        N, avail=...)."""
        if out is None:
            out = self._buf("rand_actions_%s" % dtype, (self.E, self.N), dtype)
        else:
            self._check_out("actions", out, (self.E, self.N), dtype)
        if avail is not None:
            self._check_out("avail", avail, (self.E, self.N, self.n_actions), torch.uint8)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_random_actions(self._h, self._ptr(avail), int(seed) & 0xFFFFFFFF,
                                                     int(step) & 0xFFFFFFFF, int(env_offset), self._ptr(out),
                                                     U8 if dtype == torch.uint8 else I64, self._stream()),
                        "mapf_random_actions")
        return out

    def runner_mask_actions(self, actions, alive, stay_action, out_u8, out_i64=None):
        """Rollout bookkeeping (mapf_runner_mask_actions): actions [E,N] uint8 / int64 -> out_u8 (the step's input) and
        out_i64 (the episode batch's int64 [E,N] slice), finished environments (alive bool/uint8 [E] == 0) STAY."""
        assert actions.is_cuda and actions.is_contiguous() and actions.dtype in (torch.uint8, torch.int64)
        assert alive.is_contiguous() and alive.dtype in (torch.bool, torch.uint8) and alive.numel() == self.E
        assert out_u8.dtype == torch.uint8 and out_u8.is_contiguous() and out_u8.numel() == self.E * self.N
        if out_i64 is not None:
            assert out_i64.dtype == torch.int64 and out_i64.is_contiguous() and out_i64.numel() == self.E * self.N
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_runner_mask_actions(self._h, self._ptr(actions),
                                                          U8 if actions.dtype == torch.uint8 else I64, self._ptr(alive),
                                                          int(stay_action), self._ptr(out_u8), self._ptr(out_i64),
                                                          self._stream()), "mapf_runner_mask_actions")
        return out_u8

    def runner_account(self, reward, terminated, alive, returns, lengths, filled_next):
        """Rollout bookkeeping after a step (mapf_runner_account): returns / lengths of running environments, the
        `filled` flag of the next time slot, alive &= not terminated; all [E], in place."""
        for t, dt in ((reward, torch.float64), (returns, torch.float64), (lengths, torch.int64)):
            assert t.dtype == dt and t.is_contiguous() and t.numel() == self.E
        for t in (terminated, alive, filled_next):
            assert t.dtype in (torch.uint8, torch.bool) and t.is_contiguous() and t.numel() == self.E
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_runner_account(self._h, self._ptr(reward), self._ptr(terminated), self._ptr(alive),
                                                     self._ptr(returns), self._ptr(lengths), self._ptr(filled_next),
                                                     self._stream()), "mapf_runner_account")

    def rollout_in_one_launch(self, dtype=torch.uint8):
        odt = BITS if dtype == "bits" else (F32 if dtype == torch.float32 else U8)
        return bool(self.lib.mapf_rollout_in_one_launch(self._h, odt))

    def rollout_plan(self, n_steps, dtype=torch.uint8, mid_outputs=False):
        """'pipelined' / 'in_kernel' / 'per_step': how rollout() would run (mapf_rollout_plan)."""
        odt = BITS if dtype == "bits" else (F32 if dtype == torch.float32 else U8)
        return ("per_step", "in_kernel", "pipelined")[int(self.lib.mapf_rollout_plan(self._h, int(n_steps), odt,
                                                                                  int(bool(mid_outputs))))]

    def avail(self, prev=None):
        """Action masks of the current state.  prev (uint8 [E,N]): evaluate `_listNextValidActions(id, prev_action)`
        with these previous actions instead of the stored ones -- a pure query, nothing in the handle changes."""
        out = self._buf("out_avail", (self.E, self.N, self.n_actions), torch.uint8)
        with torch.cuda.device(self.device):
            if prev is None:
                self._check(self.lib.mapf_avail(self._h, self._ptr(out), self._stream()), "mapf_avail")
            else:
                p = self._to_dev(prev, torch.uint8, (self.E, self.N))
                self._keep = [p]
                self._check(self.lib.mapf_avail_prev(self._h, self._ptr(p), self._ptr(out), self._stream()),
                            "mapf_avail_prev")
        return out

    def goal_dist(self, dirty=None, primal_costs=False, out=None):
        """int16 [E,N,H,W] hop distance to every agent's goal (walls -1, unreachable -2), computed into `out`
        (or an engine-owned tensor)."""
        dmask = self._to_dev(dirty, torch.uint8, (self.E, self.N))
        if out is None:
            out = self._buf("goal_dist", (self.E, self.N, self.H, self.W), torch.int16)
        self._keep = [dmask]
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_bfs(self._h, self._ptr(dmask), self._ptr(out), int(bool(primal_costs)),
                                          self._stream()), "mapf_bfs")
        return out

    def refresh_goal_dist(self, dirty=None):
        """Recompute the handle's own distance maps (cfg.goal_dist / mode PARTIAL) for the flagged agents."""
        dmask = self._to_dev(dirty, torch.uint8, (self.E, self.N))
        self._keep = [dmask]
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_bfs(self._h, self._ptr(dmask), None, 0, self._stream()), "mapf_bfs")

    def bind_partial_state_out(self, state):
        """PARTIAL: every later observation launch also writes get_state() into `state` (int64 [E,3], caller-owned);
        None unbinds.  One launch less per environment step than partial_state()."""
        if state is not None:
            self._check_out("state", state, (self.E, 3), torch.int64)
        self._bound_state = state          # keeps the storage alive while it is bound
        self._check(self.lib.mapf_partial_bind_state_out(self._h, self._ptr(state)), "mapf_partial_bind_state_out")

    def partial_state(self, want=("state", "at_goal", "goal_cost", "agent_steps")):
        """MARL_PARTIAL_ENV bookkeeping: state int64 [E,3] = get_state(); at_goal u8, goal_cost / agent_steps i32 [E,N]."""
        out = {}
        if "state" in want:
            out["state"] = self._buf("p_state", (self.E, 3), torch.int64)
        if "at_goal" in want:
            out["at_goal"] = self._buf("p_at_goal", (self.E, self.N), torch.uint8)
        if "goal_cost" in want:
            out["goal_cost"] = self._buf("p_goal_cost", (self.E, self.N), torch.int32)
        if "agent_steps" in want:
            out["agent_steps"] = self._buf("p_agent_steps", (self.E, self.N), torch.int32)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_partial_state(self._h, self._ptr(out.get("state")), self._ptr(out.get("at_goal")),
                                                    self._ptr(out.get("goal_cost")), self._ptr(out.get("agent_steps")),
                                                    self._stream()), "mapf_partial_state")
        return out

    # ------------------------------------------------------------------ host-buffer path (what e2e times)
    def make_host_io(self, obs_dtype=torch.uint8, want=("reward", "terminated", "dones", "avail", "obs", "vec")):
        """Pinned host buffers + the mapf_host_io descriptor for step_observe_host()."""
        E, N = self.E, self.N
        bufs = {"actions": torch.empty((E, N), dtype=torch.uint8).pin_memory()}
        if "reward" in want:
            bufs["reward"] = torch.empty((E,), dtype=torch.float64).pin_memory()
        if "terminated" in want:
            bufs["terminated"] = torch.empty((E,), dtype=torch.uint8).pin_memory()
        if "dones" in want:
            bufs["dones"] = torch.empty((E, N), dtype=torch.uint8).pin_memory()
        if "avail" in want:
            bufs["avail"] = torch.empty((E, N, self.n_actions), dtype=torch.uint8).pin_memory()
        odt = I8
        if "obs" in want:
            if self.obs_mode == OBS_PRIMAL_FOV and obs_dtype == "bits":
                # the bit stream itself (bit i = cell i of [E,N,4,F,F]); int32 words, no host expansion
                bufs["obs"] = torch.empty((self.packed_obs_bytes() // 4,), dtype=torch.int32).pin_memory()
                odt = BITS
            elif self.obs_mode == OBS_PRIMAL_FOV:
                bufs["obs"] = torch.empty((E, N, 4, self.F, self.F), dtype=obs_dtype).pin_memory()
                odt = U8 if obs_dtype == torch.uint8 else F32
            elif self.obs_mode == OBS_PARTIAL_WINDOW:
                pdt = torch.float32 if obs_dtype == torch.float32 else torch.float64
                bufs["obs"] = torch.empty((E, N, self.obs_size), dtype=pdt).pin_memory()
                odt = F32 if pdt == torch.float32 else F64
            else:
                bufs["obs"] = torch.empty((E, self.H * self.W), dtype=torch.int8).pin_memory()
        if "vec" in want and self.obs_mode == OBS_PRIMAL_FOV:
            bufs["vec"] = torch.empty((E, N, 3), dtype=torch.float64).pin_memory()
        io = MapfHostIO()
        io.actions_host = bufs["actions"].data_ptr()
        for k in ("reward", "terminated", "dones", "avail", "obs", "vec"):
            setattr(io, k + "_host", bufs[k].data_ptr() if k in bufs else None)
        io.obs_dtype = odt
        h2d = bufs["actions"].numel()
        d2h = sum(t.numel() * t.element_size() for k, t in bufs.items() if k != "actions")
        if "obs" in bufs and odt in (U8, F32) and self.obs_mode == OBS_PRIMAL_FOV and self.host_transport() == 1:
            # FOV observations cross PCIe as packed bits and are expanded by the library's host threads
            d2h += self.packed_obs_bytes() - bufs["obs"].numel() * bufs["obs"].element_size()
        return io, bufs, h2d, d2h

    def unpack_host_obs(self, bits, env_lo=0, env_hi=None, out=None, dtype=torch.uint8):
        """Lazy view of a MAPF_BITS host observation (the "obs" buffer of make_host_io(obs_dtype="bits")): expands
        only environments [env_lo, env_hi) into a [n, N, 4, F, F] host tensor (mapf_host_unpack, calling thread)."""
        env_hi = self.E if env_hi is None else int(env_hi)
        per_env = self.N * 4 * self.F * self.F
        n = env_hi - int(env_lo)
        if out is None:
            out = torch.empty((n, self.N, 4, self.F, self.F), dtype=dtype)
        assert out.is_contiguous() and out.numel() == n * per_env and out.dtype == dtype and not out.is_cuda
        rc = self.lib.mapf_host_unpack(ctypes.c_void_p(bits.data_ptr()), int(env_lo) * per_env, n * per_env,
                                       ctypes.c_void_p(out.data_ptr()), U8 if dtype == torch.uint8 else F32)
        if rc != 0:
            raise MapfError("mapf_host_unpack failed (%d)" % rc)
        return out

    def bits_supported(self):
        return bool(self.lib.mapf_obs_bits_supported(self._h))

    def host_transport(self, packed=None):
        """Query (packed=None) or set the PCIe transport of step_observe_host: 1 = bit-packed + host expansion,
        0 = dense copies.  Returns the mode in effect."""
        if packed is None:
            return int(self.lib.mapf_host_transport_get(self._h))       # side-effect free
        return int(self.lib.mapf_host_transport(self._h, int(bool(packed))))

    def packed_obs_bytes(self):
        """Bytes of one bit-packed FOV observation as the host entry point copies it (whole tiles)."""
        return (self.E * self.N * 4 * self.F * self.F + 31) // 32 * 4

    def step_observe_host(self, io):
        """Host buffers in, host buffers out; returns when the outputs are on the host."""
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_step_observe_host(self._h, ctypes.byref(io), self._stream()),
                        "mapf_step_observe_host")

    # ------------------------------------------------------------------ diagnostics
    def stats(self):
        buf = (ctypes.c_int64 * _lib.N_STATS)()
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_stats(self._h, buf, self._stream()), "mapf_stats")
        return dict(zip(STAT_NAMES, [int(v) for v in buf]))

    def error_flags(self):
        v = ctypes.c_uint32(0)
        with torch.cuda.device(self.device):
            self._check(self.lib.mapf_error_flags(self._h, ctypes.byref(v), self._stream()), "mapf_error_flags")
        return int(v.value)

    def launch_count(self):
        return int(self.lib.mapf_launch_count(self._h))
