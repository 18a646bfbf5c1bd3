#!/usr/bin/env python
"""GRID / PARTIAL fused step + observation on a c3-shaped batch, µs per step from CUDA graphs of 20 steps (the `modes`
leg of bench.py alone):   python profiles/modes_time.py [grid] [partial]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mapf_marl_b200.engine import MapfEngine  # noqa: E402
from mapf_marl_b200.workloads import WORKLOADS, make_world  # noqa: E402

WANT = ("reward", "terminated", "dones", "avail")


def timed(fn, reps):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def main():
    which = sys.argv[1:] or ["grid", "partial"]
    wl = WORKLOADS["c3"]
    E, N = int(os.environ.get("PROBE_ENVS", wl["E"])), wl["N"]
    obst, starts, goals = make_world(wl, E, 0)
    cfgs = {"grid": (dict(mode="grid", episode_limit=10 ** 6), (None,)),
            "partial": (dict(mode="partial", episode_limit=256, obs_window=11, obs_knn_agents=5),
                        (torch.float64, torch.float32))}
    for name in which:
        kw, dts = cfgs[name]
        if name == "partial" and os.environ.get("MODES_PARTIAL_DTYPE"):
            dts = (getattr(torch, os.environ["MODES_PARTIAL_DTYPE"]),)
        eng = MapfEngine(E, N, wl["H"], wl["W"], device="cuda:0", **kw)
        eng.reset(obst, starts, goals)
        pool = torch.stack([eng.random_actions(1234, t, dtype=torch.uint8).clone() for t in range(16)])
        for dt in dts:
            okw = {} if dt is None else {"dtype": dt}
            for t in range(3):
                eng.step_observe(pool[t], want=WANT, **okw)
            gs, go, gt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(gs):
                for t in range(20):
                    eng.step_observe(pool[t % 16], want=WANT, **okw)
            with torch.cuda.graph(go):
                for t in range(20):
                    eng.observe(**okw)
            with torch.cuda.graph(gt):
                for t in range(20):
                    eng.step(pool[t % 16], want=WANT)
            us = [np.median([timed(g.replay, 10) for _ in range(5)]) / 20 * 1e3 for g in (gs, go, gt)]
            print("%s %s: fused step+obs %.2f us, observe only %.2f us, step only %.2f us" % (
                name, "" if dt is None else str(dt).split(".")[-1], us[0], us[1], us[2]))
        assert eng.error_flags() == 0
        eng.close()


if __name__ == "__main__":
    main()
