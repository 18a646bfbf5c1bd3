#!/usr/bin/env python
"""A few PARTIAL (marl_partial.py) observation launches on a c3-shaped batch, for ncu (-k regex:mapf_partial_obs)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mapf_marl_b200 import maps  # noqa: E402
from mapf_marl_b200.engine import MapfEngine  # noqa: E402

E, N, H, W = 16384, 32, 32, 32
obst, starts, goals = maps.synthetic_batch(1, E, H, W, 0.3, N, distinct=0)
eng = MapfEngine(E, N, H, W, mode="partial", episode_limit=256, obs_window=11, obs_knn_agents=5)
eng.reset(obst, starts, goals)
acts = torch.randint(0, 5, (E, N), device="cuda", dtype=torch.uint8)
dt = torch.float32 if "--f32" in sys.argv else torch.float64
for _ in range(4):
    eng.step(acts)
    eng.observe(dtype=dt) if "--f32" in sys.argv else eng.observe()
torch.cuda.synchronize()
print("ok")
