import sys, os, time, numpy as np, torch
sys.path.insert(0, ".")
from mapf_marl_b200.engine import MapfEngine
from mapf_marl_b200.workloads import WORKLOADS, make_world
wl = WORKLOADS["c3"]; E, N = wl["E"], wl["N"]
o, s, g = make_world(wl, E, 0)
eng = MapfEngine(E, N, 32, 32, mode="primal", fov=11)
eng.reset(o, s, g)
for want in (("reward", "terminated", "dones", "avail", "obs", "vec"), ("reward", "terminated", "obs")):
    io, bufs, h2d, d2h = eng.make_host_io(want=want)
    bufs["actions"].random_(0, 5)
    for _ in range(10):
        eng.step_observe_host(io)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); eng.step_observe_host(io); ts.append(time.perf_counter() - t0)
    print(want, "median ms %.3f min %.3f" % (np.median(ts) * 1e3, min(ts) * 1e3), flush=True)
os.environ["X"] = "1"
