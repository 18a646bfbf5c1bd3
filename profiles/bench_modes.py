#!/usr/bin/env python
"""Throughput of the other two reference semantics on a c3-shaped batch (informational; bench.py times the PRIMAL path
that BASELINE.json names): GRID (mapf_gridworld.py: joint step + full-map observation) and PARTIAL (marl_partial.py,
the env the reference registers: joint step + window maps + K nearest agents, float64)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mapf_marl_b200 import maps  # noqa: E402
from mapf_marl_b200.engine import MapfEngine  # noqa: E402


def timed(fn, steps, warmup=10):
    for t in range(warmup):
        fn(t)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(steps):
        fn(t)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def main():
    E, N, H, W = 16384, 32, 32, 32
    obst, starts, goals = maps.synthetic_batch(1, E, H, W, 0.3, N, distinct=0)
    pool = torch.randint(0, 5, (16, E, N), device="cuda", dtype=torch.uint8)
    out = {}
    for mode, kw in (("grid", dict(episode_limit=10 ** 6)),
                     ("partial", dict(episode_limit=256, obs_window=11, obs_knn_agents=5))):
        eng = MapfEngine(E, N, H, W, mode=mode, **kw)
        eng.reset(obst, starts, goals)
        ms_fused = timed(lambda t: eng.step_observe(pool[t % 16]), 300)
        ms_step = timed(lambda t: eng.step(pool[t % 16]), 300)
        ms_obs = timed(lambda t: eng.observe(), 300)
        sys.stderr.write("%s fused %.4f step %.4f obs %.4f ms\n" % (mode, ms_fused, ms_step, ms_obs))
        obs, _ = eng.observe()
        out[mode] = {"step_observe_ms": ms_fused, "step_ms": ms_step, "observe_ms": ms_obs,
                     "agent_steps_per_s": E * N / (ms_fused * 1e-3), "obs_bytes_per_step": obs.numel() * obs.element_size(),
                     "obs_GBps": obs.numel() * obs.element_size() / (ms_obs * 1e-3) / 1e9}
        if mode == "partial":
            ms_obs32 = timed(lambda t: eng.observe(dtype=torch.float32), 300)
            o32, _ = eng.observe(dtype=torch.float32)
            out[mode]["observe_f32_ms"] = ms_obs32
            out[mode]["obs_f32_GBps"] = o32.numel() * 4 / (ms_obs32 * 1e-3) / 1e9
            ms_reset = timed(lambda t: eng.reset(obst, starts, goals), 5, 1)
            out[mode]["reset_with_goal_maps_ms"] = ms_reset
        eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
