#!/usr/bin/env python
"""How much of the fused launch is the per-environment reward chain (one thread adds N float64 rewards in the
reference's order)?  Times the fused step+obs launch with and without the `reward` output, c3 and c4."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mapf_marl_b200.engine import MapfEngine  # noqa: E402


def timed(fn, steps, warmup=20):
    for t in range(warmup):
        fn(t)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(steps):
        fn(t)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps * 1e3


for name in sys.argv[1:] or ("c3", "c4"):
    wl = bench.WORKLOADS[name]
    E, N = wl["E"], wl["N"]
    obst, starts, goals = bench.make_world(wl, E, 0)
    eng = MapfEngine(E, N, wl["H"], wl["W"], mode="primal", fov=wl["F"], shared_map=wl["warehouse"])
    eng.reset(obst, starts, goals)
    pool = torch.randint(0, 5, (16, E, N), device="cuda", dtype=torch.uint8)
    for want in (("reward", "terminated", "dones", "avail"), ("terminated", "dones", "avail"),
                 ("agent_reward", "terminated", "dones", "avail"), ("terminated",)):
        us = timed(lambda t: eng.step_observe(pool[t % 16], want=want), 1000)
        print(name, want, "%.2f us" % us)
    eng.close()
