#!/usr/bin/env python
"""Times mapf_rollout (T steps per launch) against T fused launches replayed from a CUDA graph, per workload.
    python profiles/rollout_probe.py [c2 c3 c4] [--T 16]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mapf_marl_b200.engine import MapfEngine  # noqa: E402
from mapf_marl_b200.workloads import WORKLOADS, make_world, hash_actions_torch  # noqa: E402


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
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["c2", "c3"])
    ap.add_argument("--T", type=int, default=16)
    ap.add_argument("--envs", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    want = ("reward", "terminated", "dones", "avail")
    for name in args.workloads:
        wl = dict(WORKLOADS[name])
        E = args.envs or wl["E"]
        N, F, T = wl["N"], wl["F"], args.T
        obst, starts, goals = make_world(wl, E, 0)
        eng = MapfEngine(E, N, wl["H"], wl["W"], mode="primal", fov=F, shared_map=wl["warehouse"], device=dev)
        eng.reset(obst, starts, goals)
        acts = torch.stack([hash_actions_torch(1234, 0, E, t, N, dev) for t in range(T)])
        for _ in range(3):
            eng.rollout(acts, want=want)
        ms_roll = timed(lambda: eng.rollout(acts, want=want), max(3, 2000 // T)) / T
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for t in range(T):
                eng.step_observe(acts[t], want=want)
        g.replay()
        ms_graph = timed(g.replay, max(3, 2000 // T)) / T
        bytes_step = wl["bytes_per_agent_step"] * E * N
        print(json.dumps({"workload": name, "E": E, "T": T, "epb_env": os.environ.get("MAPF_B200_EPB"),
                          "one_launch": eng.rollout_in_one_launch(),
                          "rollout_us_per_step": ms_roll * 1e3, "graph_us_per_step": ms_graph * 1e3,
                          "rollout_frac_hbm": bytes_step / (ms_roll * 1e-3) / 1e9 / 6540.2,
                          "graph_frac_hbm": bytes_step / (ms_graph * 1e-3) / 1e9 / 6540.2}))
        eng.close()


if __name__ == "__main__":
    main()
