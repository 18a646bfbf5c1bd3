"""ctypes binding of libmapf_b200.so (the C ABI declared in include/mapf_b200.h).

The library is built in-tree by `build()` with nvcc for sm_100a.  There is no CPU
fallback: if the shared library is missing or no CUDA device is present the product
path raises.
"""
import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("MAPF_B200_LIB") or os.path.join(_HERE, "libmapf_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mapf_b200.h")

ABI_VERSION = 1

MAPF_OK = 0
MODE_GRID, MODE_PRIMAL, MODE_PARTIAL = 0, 1, 2
OBS_FULLMAP, OBS_PRIMAL_FOV, OBS_PARTIAL_WINDOW = 0, 1, 2
U8, I64, F32, I8, F64, BITS = 0, 1, 2, 3, 4, 5
FLAG_BAD_ACTION, FLAG_BAD_POSITION, FLAG_START_ON_WALL, FLAG_START_OVERLAP, FLAG_GOAL_OVERLAP = 1, 2, 4, 8, 16
FLAG_INTERNAL = 32
N_STATS = 8
STAT_NAMES = ("env_steps", "agent_steps", "env_collisions", "node_collisions", "edge_collisions",
              "goal_arrivals", "episodes_done", "reserved")

_vp = ctypes.c_void_p


class MapfCfg(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("n_envs", ctypes.c_int32), ("n_agents", ctypes.c_int32),
        ("height", ctypes.c_int32), ("width", ctypes.c_int32), ("mode", ctypes.c_int32),
        ("obs_mode", ctypes.c_int32), ("fov", ctypes.c_int32), ("shared_map", ctypes.c_int32),
        ("episode_limit", ctypes.c_int32), ("goal_dist", ctypes.c_int32), ("collect_stats", ctypes.c_int32),
        ("step_reward", ctypes.c_double), ("collide_reward", ctypes.c_double),
        ("action_cost", ctypes.c_double), ("idle_cost", ctypes.c_double),
        ("goal_reward", ctypes.c_double), ("collision_reward", ctypes.c_double),
        ("mag_lut_host", _vp), ("mag_lut_len", ctypes.c_int32),
        ("reward_sum_mode", ctypes.c_int32), ("step_reward_is_int", ctypes.c_int32),
        ("collide_reward_is_int", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("obs_window", ctypes.c_int32), ("obs_knn_agents", ctypes.c_int32),
        ("move_reward", ctypes.c_double), ("stay_reward", ctypes.c_double), ("stay_goal_reward", ctypes.c_double),
        ("node_collide_reward", ctypes.c_double), ("edge_collide_reward", ctypes.c_double),
        ("env_collide_reward", ctypes.c_double),
        ("complete_lut_host", _vp), ("complete_lut_len", ctypes.c_int32), ("blocking_reward", ctypes.c_int32),
        ("blocking_cost", ctypes.c_double),
        ("diagonal_movement", ctypes.c_int32), ("reserved3", ctypes.c_int32),
    ]


STEP_OUT_FIELDS = ("reward", "terminated", "agent_reward", "dones", "status", "node", "edge", "valid",
                   "done_mid", "next_mid", "avail", "blocking")


class MapfStepOut(ctypes.Structure):
    _fields_ = [(name + "_dev", _vp) for name in STEP_OUT_FIELDS]


class MapfHostIO(ctypes.Structure):
    _fields_ = [
        ("actions_host", _vp), ("reward_host", _vp), ("terminated_host", _vp), ("dones_host", _vp),
        ("avail_host", _vp), ("obs_host", _vp), ("vec_host", _vp), ("obs_dtype", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


# name -> (restype, argtypes); every function include/mapf_b200.h declares
_i, _i64 = ctypes.c_int, ctypes.c_int64
PROTOTYPES = {
    "mapf_last_error": (ctypes.c_char_p, [_vp]),
    "mapf_default_cfg": (None, [ctypes.POINTER(MapfCfg)]),
    "mapf_create": (_i, [ctypes.POINTER(MapfCfg), ctypes.POINTER(_vp)]),
    "mapf_destroy": (_i, [_vp]),
    "mapf_reset": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mapf_set_goals": (_i, [_vp, _vp, _vp, _vp]),
    "mapf_pop_goals": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "mapf_lifelong_bind": (_i, [_vp, _vp, _vp, _i]),
    "mapf_bfs_popped": (_i, [_vp, _vp, _vp]),
    "mapf_step": (_i, [_vp, _vp, _i, ctypes.POINTER(MapfStepOut), _vp]),
    "mapf_step_agents": (_i, [_vp, _vp, _i, _i, _i, ctypes.POINTER(MapfStepOut), _vp]),
    "mapf_observe": (_i, [_vp, _vp, _i, _vp, _vp]),
    "mapf_step_observe": (_i, [_vp, _vp, _i, ctypes.POINTER(MapfStepOut), _vp, _i, _vp, _vp]),
    "mapf_rollout": (_i, [_vp, _vp, _i, _i, ctypes.POINTER(MapfStepOut), _vp, _i, _vp, _vp]),
    "mapf_rollout_in_one_launch": (_i, [_vp, _i]),
    "mapf_rollout_plan": (_i, [_vp, _i, _i, _i]),
    "mapf_step_observe_host": (_i, [_vp, ctypes.POINTER(MapfHostIO), _vp]),
    "mapf_host_unpack": (_i, [_vp, ctypes.c_uint64, ctypes.c_uint64, _vp, _i]),
    "mapf_obs_bits_supported": (_i, [_vp]),
    "mapf_host_transport": (_i, [_vp, _i]),
    "mapf_host_transport_get": (_i, [_vp]),
    "mapf_avail": (_i, [_vp, _vp, _vp]),
    "mapf_avail_prev": (_i, [_vp, _vp, _vp, _vp]),
    "mapf_bfs": (_i, [_vp, _vp, _vp, _i, _vp]),
    "mapf_set_prev_actions": (_i, [_vp, _vp, _vp]),
    "mapf_get_positions": (_i, [_vp, _vp, _vp]),
    "mapf_get_goals": (_i, [_vp, _vp, _vp]),
    "mapf_get_dones": (_i, [_vp, _vp, _vp]),
    "mapf_get_step_count": (_i, [_vp, _vp, _vp]),
    "mapf_partial_state": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mapf_random_actions": (_i, [_vp, _vp, ctypes.c_uint32, ctypes.c_uint32, _i64, _vp, _i, _vp]),
    "mapf_partial_bind_state_out": (_i, [_vp, _vp]),
    "mapf_runner_mask_actions": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "mapf_runner_account": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mapf_stats": (_i, [_vp, _vp, _vp]),
    "mapf_error_flags": (_i, [_vp, _vp, _vp]),
    "mapf_launch_count": (_i64, [_vp]),
    "mapf_abi_version": (_i, []),
    "mapf_build_arch": (ctypes.c_char_p, []),
}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]
SOURCES = ["mapf_kernels.cu", "mapf_capi.cu", "mapf_host_unpack.cpp"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmapf_b200.so cannot be built")


def build(force=False, verbose=False):
    """Compile csrc/*.cu into mapf_marl_b200/libmapf_b200.so for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(_CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(_CSRC, "mapf_internal.h"), HEADER_PATH]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(p) for p in deps):
        return LIB_PATH
    cmd = [_nvcc()]
    if os.path.exists("/usr/bin/g++"):
        cmd += ["-ccbin", "/usr/bin/g++"]
    cmd += NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_LIB = None


def load():
    """Load the shared library and attach prototypes.  Raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmapf_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or mapf_marl_b200.build(); there is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.mapf_abi_version() != ABI_VERSION:
            raise RuntimeError("libmapf_b200.so ABI %d != binding ABI %d; rebuild" % (lib.mapf_abi_version(), ABI_VERSION))
        _LIB = lib
    return _LIB


def last_error(lib, handle=None):
    msg = lib.mapf_last_error(handle)
    return msg.decode() if msg else ""
