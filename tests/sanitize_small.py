"""Small driver for compute-sanitizer (memcheck): every mode, ragged sizes, all outputs, a few steps.
Usage on the GPU box:  compute-sanitizer --tool memcheck python tests/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapf_marl_b200 import maps  # noqa: E402
from mapf_marl_b200.engine import MapfEngine  # noqa: E402

ALL_PRIMAL = ("reward", "terminated", "agent_reward", "dones", "status", "valid", "done_mid", "next_mid", "avail")
ALL_GRID = ("reward", "terminated", "agent_reward", "dones", "status", "node", "edge", "avail")


def main():
    rs = np.random.RandomState(0)
    for (E, N, H, W, F) in ((37, 7, 13, 40, 11), (5, 32, 32, 32, 11), (9, 3, 5, 5, 3), (4, 130, 30, 30, 10), (6, 9, 9, 9, 6)):
        obst, starts, goals = maps.synthetic_batch(1, E, H, W, 0.15, N, distinct=0)
        eng = MapfEngine(E, N, H, W, mode="primal", fov=F, goal_dist=True)
        eng.reset(obst, starts, goals)
        eng.refresh_goal_dist()
        eng.goal_dist(primal_costs=True)
        for t in range(4):
            a = torch.as_tensor(rs.randint(0, 5, (E, N)).astype(np.uint8), device="cuda")
            eng.step_observe(a, want=ALL_PRIMAL, dtype=torch.float32 if t % 2 else torch.uint8)
            eng.step(a, want=ALL_PRIMAL, agent_range=(0, max(1, N // 2)))
            eng.observe()
            eng.avail()
        io, bufs, _, _ = eng.make_host_io()
        bufs["actions"].zero_()
        eng.step_observe_host(io)
        eng.close()
    for mode, kw in (("grid", {}), ("partial", dict(obs_window=5, obs_knn_agents=4))):
        E, N, H, W = 11, 10, 12, 12
        obst = np.zeros((E, H, W), np.uint8)
        starts = rs.randint(0, 12, (E, N, 2)).astype(np.int16)
        goals = rs.randint(0, 12, (E, N, 2)).astype(np.int16)
        eng = MapfEngine(E, N, H, W, mode=mode, episode_limit=6, **kw)
        eng.reset(obst, starts, goals)
        for t in range(8):
            a = torch.as_tensor(rs.randint(0, 5, (E, N)).astype(np.uint8), device="cuda")
            eng.step_observe(a, want=ALL_GRID)
        if mode == "partial":
            eng.partial_state()
        eng.close()
    torch.cuda.synchronize()
    print("sanitize_small: done")


if __name__ == "__main__":
    main()
