"""Load the UNMODIFIED reference env classes (test infrastructure, like everything under oracle/).

Where they come from, in this order:
  1. $MAPF_REFERENCE_ROOT, if set;
  2. /root/reference (the build container);
  3. oracle/_ref/reference_py.zip -- the five files of the path, packed byte for byte by oracle/stage_ref.py in the
     build container (git-ignored like the built .so files, but shipped to the GPU box with them) and unpacked into a
     scratch directory here, which is what lets bench.py time the real Python reference on the GPU box's host cores.
Used by tests/golden/gen_golden.py (fixture generation), by bench.py's cpu_baseline / `--impl reference` legs and by
the reference-vs-oracle CPU tests.  The product package never imports this module.

The reference imports a few packages that are absent here (gym, smac, matplotlib,
od_mstar3). They are replaced by the thinnest possible stand-ins:
  gym.Env = object, gym.spaces.{Discrete,Tuple}  -- only used to declare action_space
  smac.env.multiagentenv.MultiAgentEnv           -- the reference's OWN
        MARL-curve-main/src/envs/multiagentenv.py class
  matplotlib.colors.hsv_to_rgb                   -- render-only
  od_mstar3.cpp_mstar.find_path                  -- always raises NoSolutionError, which
        makes MAPFEnv.get_blocking_reward() return 0 * BLOCKING_COST: the blocking
        reward (un-vendored third-party arithmetic) is fenced off.
"""
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(_HERE, "_ref", "reference_py.zip")


def _find_root():
    env = os.environ.get("MAPF_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/mapf_primal.py"):
        return "/root/reference"
    if os.path.isfile(STAGED):
        # the GPU box: unpack the staged archive into a scratch directory (checked against its manifest) and let
        # child processes reuse it
        import atexit
        import shutil
        import tempfile
        from . import stage_ref
        root = stage_ref.unpack(tempfile.mkdtemp(prefix="mapf_ref_"))
        os.environ["MAPF_REFERENCE_ROOT"] = root
        atexit.register(shutil.rmtree, root, True)
        return root
    return "/root/reference"


REF = _find_root()
REF_SRC = os.path.join(REF, "MARL-curve-main", "src")


def available():
    return os.path.isfile(os.path.join(REF, "mapf_primal.py"))


def _install_stubs(with_smac=True):
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.Env = object
        spaces = types.ModuleType("gym.spaces")

        class Discrete:
            def __init__(self, n):
                self.n = n

        class Tuple(tuple):
            def __new__(cls, items):
                return super().__new__(cls, items)

        spaces.Discrete = Discrete
        spaces.Tuple = Tuple
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "smac" not in sys.modules and with_smac:
        spec = importlib.util.spec_from_file_location(
            "_ref_multiagentenv", os.path.join(REF_SRC, "envs", "multiagentenv.py"))
        mae = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mae)
        smac = types.ModuleType("smac")
        smac_env = types.ModuleType("smac.env")
        smac_mae = types.ModuleType("smac.env.multiagentenv")
        smac_mae.MultiAgentEnv = mae.MultiAgentEnv
        smac.env = smac_env
        smac_env.multiagentenv = smac_mae
        sys.modules["smac"] = smac
        sys.modules["smac.env"] = smac_env
        sys.modules["smac.env.multiagentenv"] = smac_mae
    try:
        import matplotlib.colors  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        colors = types.ModuleType("matplotlib.colors")
        colors.hsv_to_rgb = lambda x: x
        mpl.colors = colors
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.colors"] = colors
    if "od_mstar3" not in sys.modules:
        od = types.ModuleType("od_mstar3")
        cpp = types.ModuleType("od_mstar3.cpp_mstar")
        csa = types.ModuleType("od_mstar3.col_set_addition")

        class NoSolutionError(Exception):
            pass

        class OutOfTimeError(Exception):
            pass

        def find_path(world, starts, goals, inflation, time_limit):
            """Stand-in for od_mstar3.cpp_mstar.find_path (un-vendored).  Default: no solution, which fences the
            blocking reward off (get_blocking_reward returns 0).  With MAPF_REF_BFS_MSTAR=1 it is a single-agent
            shortest path (4-connected BFS, what M* with inflation 1 returns for one robot): a list of joint
            configurations from start to goal inclusive, NoSolutionError when the goal cannot be reached."""
            if not os.environ.get("MAPF_REF_BFS_MSTAR"):
                raise NoSolutionError()
            import collections
            (sx, sy), (gx, gy) = tuple(starts[0]), tuple(goals[0])
            H, W = world.shape
            if world[sx, sy] != 0 or world[gx, gy] != 0:
                raise NoSolutionError()
            prev = {(sx, sy): None}
            dq = collections.deque([(sx, sy)])
            while dq:
                cur = dq.popleft()
                if cur == (gx, gy):
                    break
                for dx, dy in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                    n = (cur[0] + dx, cur[1] + dy)
                    if 0 <= n[0] < H and 0 <= n[1] < W and world[n] == 0 and n not in prev:
                        prev[n] = cur
                        dq.append(n)
            if (gx, gy) not in prev:
                raise NoSolutionError()
            path, cur = [], (gx, gy)
            while cur is not None:
                path.append((cur,))
                cur = prev[cur]
            return path[::-1]

        csa.NoSolutionError = NoSolutionError
        csa.OutOfTimeError = OutOfTimeError
        cpp.find_path = find_path
        od.cpp_mstar = cpp
        od.col_set_addition = csa
        sys.modules["od_mstar3"] = od
        sys.modules["od_mstar3.cpp_mstar"] = cpp
        sys.modules["od_mstar3.col_set_addition"] = csa
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)  # for `utils.draw`


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns (mapf_gridworld, mapf_primal, marl_partial) reference modules."""
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present at %s" % REF)
    _install_stubs()
    grid = _load("_ref_mapf_gridworld", os.path.join(REF, "mapf_gridworld.py"))
    primal = _load("_ref_mapf_primal", os.path.join(REF, "mapf_primal.py"))
    partial = _load("_ref_marl_partial", os.path.join(REF_SRC, "envs", "marl_partial.py"))
    return grid, primal, partial


def load_primal():
    """Only mapf_primal.py (numpy + the stubs: no torch / cv2 import), for the worker processes of the CPU baseline."""
    if not available():
        raise RuntimeError("reference tree not present at %s (run oracle/stage_ref.py in the build container)" % REF)
    _install_stubs(with_smac=False)
    return _load("_ref_mapf_primal", os.path.join(REF, "mapf_primal.py"))
