"""ctypes wrapper around oracle/libmapf_oracle.so (test infrastructure, see __init__.py)."""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODE_GRID = 0
MODE_PRIMAL = 1
MODE_PARTIAL = 2


def build_oracle(force=False):
    so = os.path.join(_HERE, "libmapf_oracle.so")
    src = os.path.join(_HERE, "mapf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s", "libmapf_oracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_oracle())
        vp, i, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        lib.oracle_create.restype = vp
        lib.oracle_create.argtypes = [i, i, i, i, i, i, i, d, d, d, d, d, d, i, i, i, i]
        lib.oracle_destroy.argtypes = [vp]
        lib.oracle_max_threads.restype = i
        for name, n in (("oracle_get_positions", 2), ("oracle_get_dones", 2), ("oracle_get_step_count", 2),
                        ("oracle_grid_reset", 4), ("oracle_grid_avail", 2), ("oracle_grid_state", 2),
                        ("oracle_primal_reset", 4), ("oracle_primal_set_goals", 3),
                        ("oracle_primal_avail", 3), ("oracle_primal_observe", 3)):
            getattr(lib, name).argtypes = [vp] * n
            getattr(lib, name).restype = None
        lib.oracle_grid_step.argtypes = [vp] * 9
        lib.oracle_grid_step.restype = i
        lib.oracle_primal_sweep.argtypes = [vp, vp, i, i] + [vp] * 10
        lib.oracle_primal_set_blocking.argtypes = [vp, i]
        lib.oracle_primal_set_blocking.restype = None
        lib.oracle_primal_set_diagonal.argtypes = [vp, i]
        lib.oracle_primal_set_diagonal.restype = None
        lib.oracle_primal_sweep.restype = i
        lib.oracle_goal_dist.argtypes = [vp, vp, i, vp]
        lib.oracle_goal_dist.restype = None
        lib.oracle_partial_config.argtypes = [vp, i, i] + [d] * 9
        lib.oracle_partial_config.restype = None
        lib.oracle_partial_reset.argtypes = [vp] * 4
        lib.oracle_partial_reset.restype = None
        lib.oracle_partial_step.argtypes = [vp] * 6
        lib.oracle_partial_step.restype = i
        lib.oracle_partial_get.argtypes = [vp] * 6
        lib.oracle_partial_get.restype = None
        lib.oracle_partial_state.argtypes = [vp] * 2
        lib.oracle_partial_state.restype = None
        lib.oracle_partial_observe.argtypes = [vp] * 2
        lib.oracle_partial_observe.restype = None
        _LIB = lib
    return _LIB


def oracle_max_threads():
    return int(_lib().oracle_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """Batched CPU environments with the reference's semantics (numpy in / numpy out)."""

    def __init__(self, n_envs, n_agents, height, width, mode, fov=11, shared_map=False, episode_limit=10000,
                 step_reward=-0.01, collide_reward=-10, action_cost=-0.3, idle_cost=-0.5, goal_reward=0.0,
                 collision_reward=-2.0, threads=0, sum_mode=None, step_is_int=None, collide_is_int=None):
        # sum_mode None: behave like the interpreter running this process would on the reference
        # (CPython >= 3.12 builtin sum() compensates float sums, see py_sum in mapf_oracle.c)
        if sum_mode is None:
            sum_mode = 1 if sys.version_info >= (3, 12) else 0
        if step_is_int is None:
            step_is_int = isinstance(step_reward, int)
        if collide_is_int is None:
            collide_is_int = isinstance(collide_reward, int)
        self.E, self.N, self.H, self.W, self.F = n_envs, n_agents, height, width, fov
        self.mode = mode
        self.n_actions = 5
        self.shared_map = bool(shared_map)
        self._lib = _lib()
        self._h = self._lib.oracle_create(n_envs, n_agents, height, width, fov, int(self.shared_map),
                                          int(episode_limit), float(step_reward), float(collide_reward),
                                          float(action_cost), float(idle_cost), float(goal_reward),
                                          float(collision_reward), int(threads), int(sum_mode),
                                          int(bool(step_is_int)), int(bool(collide_is_int)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.oracle_destroy(self._h)
            self._h = None

    # ---- state
    def reset(self, obst=None, starts=None, goals=None):
        m = None if obst is None else np.ascontiguousarray(obst, dtype=np.int8)
        s = None if starts is None else np.ascontiguousarray(starts, dtype=np.int16)
        g = None if goals is None else np.ascontiguousarray(goals, dtype=np.int16)
        if m is not None:
            assert m.size == (1 if self.shared_map else self.E) * self.H * self.W
        for a in (s, g):
            if a is not None:
                assert a.shape == (self.E, self.N, 2)
        fn = {MODE_GRID: self._lib.oracle_grid_reset, MODE_PRIMAL: self._lib.oracle_primal_reset,
              MODE_PARTIAL: self._lib.oracle_partial_reset}[self.mode]
        fn(self._h, _p(m), _p(s), _p(g))

    def set_goals(self, goals, dirty=None):
        g = np.ascontiguousarray(goals, dtype=np.int16)
        dm = None if dirty is None else np.ascontiguousarray(dirty, dtype=np.uint8)
        self._lib.oracle_primal_set_goals(self._h, _p(g), _p(dm))

    def positions(self):
        out = np.empty((self.E, self.N, 2), np.int16)
        self._lib.oracle_get_positions(self._h, _p(out))
        return out

    def dones(self):
        out = np.empty((self.E, self.N), np.uint8)
        self._lib.oracle_get_dones(self._h, _p(out))
        return out

    def step_count(self):
        out = np.empty((self.E,), np.int32)
        self._lib.oracle_get_step_count(self._h, _p(out))
        return out

    # ---- GRID
    def grid_step(self, actions, want=("reward", "terminated", "agent_reward", "status", "node", "edge", "avail")):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        assert a.shape == (self.E, self.N)
        E, N = self.E, self.N
        out = dict(reward=np.empty(E, np.float64), terminated=np.empty(E, np.uint8),
                   agent_reward=np.empty((E, N), np.float64), status=np.empty((E, N), np.int8),
                   node=np.empty((E, N), np.int16), edge=np.empty((E, N), np.int16),
                   avail=np.empty((E, N, 5), np.uint8))
        out = {k: v for k, v in out.items() if k in want}
        bad = self._lib.oracle_grid_step(self._h, _p(a), _p(out.get("reward")), _p(out.get("terminated")),
                                         _p(out.get("agent_reward")), _p(out.get("status")), _p(out.get("node")),
                                         _p(out.get("edge")), _p(out.get("avail")))
        out["bad_actions"] = bad
        out["dones"] = self.dones()
        return out

    def grid_state(self):
        out = np.empty((self.E, self.H * self.W), np.int8)
        self._lib.oracle_grid_state(self._h, _p(out))
        return out

    def grid_avail(self):
        out = np.empty((self.E, self.N, 5), np.uint8)
        self._lib.oracle_grid_avail(self._h, _p(out))
        return out

    # ---- PRIMAL
    def primal_sweep(self, actions, lo=0, hi=None,
                     want=("status", "agent_reward", "dones", "valid", "done_mid", "next_mid", "avail",
                           "terminated", "reward", "blocking")):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        assert a.shape == (self.E, self.N)
        hi = self.N if hi is None else hi
        E, N = self.E, self.N
        out = dict(status=np.zeros((E, N), np.int8), agent_reward=np.zeros((E, N), np.float64),
                   dones=np.zeros((E, N), np.uint8), valid=np.zeros((E, N), np.uint8),
                   done_mid=np.zeros((E, N), np.uint8), next_mid=np.zeros((E, N, self.n_actions), np.uint8),
                   avail=np.zeros((E, N, self.n_actions), np.uint8), terminated=np.zeros(E, np.uint8),
                   reward=np.zeros(E, np.float64), blocking=np.zeros((E, N), np.uint8))
        out = {k: v for k, v in out.items() if k in want}
        bad = self._lib.oracle_primal_sweep(self._h, _p(a), int(lo), int(hi), _p(out.get("status")),
                                            _p(out.get("agent_reward")), _p(out.get("dones")),
                                            _p(out.get("valid")), _p(out.get("done_mid")), _p(out.get("next_mid")),
                                            _p(out.get("avail")), _p(out.get("terminated")), _p(out.get("reward")),
                                            _p(out.get("blocking")))
        out["bad_actions"] = bad
        return out

    def set_diagonal(self, on=True):
        """PRIMAL DIAGONAL_MOVEMENT (mapf_primal.py:175): 9 actions; call before reset()."""
        self._lib.oracle_primal_set_diagonal(self._h, int(bool(on)))
        self.n_actions = 9 if on else 5

    def set_blocking(self, on=True):
        """PRIMAL blocking reward (mapf_primal.py:513-546) with single-robot BFS path lengths."""
        self._lib.oracle_primal_set_blocking(self._h, int(bool(on)))

    def primal_avail(self, prev_action=None):
        out = np.empty((self.E, self.N, self.n_actions), np.uint8)
        pa = None if prev_action is None else np.ascontiguousarray(prev_action, dtype=np.uint8)
        self._lib.oracle_primal_avail(self._h, _p(pa), _p(out))
        return out

    def primal_observe(self, want_vec=True):
        obs = np.empty((self.E, self.N, 4, self.F, self.F), np.uint8)
        vec = np.empty((self.E, self.N, 3), np.float64) if want_vec else None
        self._lib.oracle_primal_observe(self._h, _p(obs), _p(vec))
        return obs, vec

    # ---- distance maps
    def goal_dist(self, dirty=None, primal_costs=False, out=None):
        if out is None:
            out = np.full((self.E, self.N, self.H, self.W), -9, np.int16)
        dm = None if dirty is None else np.ascontiguousarray(dirty, dtype=np.uint8)
        self._lib.oracle_goal_dist(self._h, _p(dm), int(bool(primal_costs)), _p(out))
        return out

    # ---- PARTIAL (marl_partial.py)
    def partial_config(self, obs_window=5, obs_knn_agents=5, move_reward=-0.01, stay_reward=-0.02, stay_goal_reward=0,
                       node_collide_reward=-1, edge_collide_reward=-1, env_collide_reward=-1, complete_reward=1000,
                       complete_fac=1.5, gamma=0.99):
        self.pW, self.pK = int(obs_window), int(obs_knn_agents)
        self._lib.oracle_partial_config(self._h, self.pW, self.pK, float(move_reward), float(stay_reward),
                                        float(stay_goal_reward), float(node_collide_reward),
                                        float(edge_collide_reward), float(env_collide_reward), float(complete_reward),
                                        float(complete_fac), float(gamma))

    def partial_step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        assert a.shape == (self.E, self.N)
        E, N = self.E, self.N
        out = dict(reward=np.empty(E, np.float64), terminated=np.empty(E, np.uint8),
                   agent_reward=np.empty((E, N), np.float64), avail=np.empty((E, N, 5), np.uint8))
        out["bad_actions"] = self._lib.oracle_partial_step(self._h, _p(a), _p(out["reward"]), _p(out["terminated"]),
                                                           _p(out["agent_reward"]), _p(out["avail"]))
        out.update(self.partial_get())
        out["dones"] = self.dones()
        return out

    def partial_get(self):
        E, N = self.E, self.N
        out = dict(at_goal=np.empty((E, N), np.uint8), goal_cost=np.empty((E, N), np.int32),
                   agent_steps=np.empty((E, N), np.int32), node=np.empty((E, N), np.int16),
                   edge=np.empty((E, N), np.int16))
        self._lib.oracle_partial_get(self._h, _p(out["at_goal"]), _p(out["goal_cost"]), _p(out["agent_steps"]),
                                     _p(out["node"]), _p(out["edge"]))
        return out

    def partial_state(self):
        out = np.empty((self.E, 3), np.int64)
        self._lib.oracle_partial_state(self._h, _p(out))
        return out

    def partial_observe(self):
        obs = np.empty((self.E, self.N, 2 * self.pW * self.pW + 13 * self.pK), np.float64)
        self._lib.oracle_partial_observe(self._h, _p(obs))
        return obs
