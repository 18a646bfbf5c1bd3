#!/usr/bin/env python
"""One worker process of the CPU reference baseline (test / benchmark infrastructure).

Steps the UNMODIFIED reference env (`mapf_primal.MAPFEnv`, loaded by oracle/refload.py) exactly the way the reference's
own rollout does: one env instance per process slot (MARL-curve-main/src/runners/parallel_runner.py:219-258), a joint
step = `_step((id, a))` swept over ids 1..N (mapf_primal.py:549-637) followed by `_observe(id)` for every agent
(:343-386) and `_listNextValidActions(id, a)` (:639-667) -- the outputs one fused GPU launch produces.

Protocol (driven by oracle/ref_pool.py): reads one JSON config line from argv[1]; builds its environments; prints
"ready"; waits for a line on stdin; runs `warmup` untimed and `steps` timed joint steps over all of its environments;
prints one JSON line {"elapsed": s, "agent_steps": n, "checksum": c}.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def build_envs(primal, wl, env_lo, n_envs, seed):
    from mapf_marl_b200 import workloads
    obst, starts, goals = workloads.make_world(wl, n_envs, env_lo, seed=seed)
    envs = []
    for e in range(n_envs):
        m = obst if wl["warehouse"] else obst[e]
        world = -(np.asarray(m).astype(int))
        gl = np.zeros_like(world)
        for k in range(wl["N"]):
            world[tuple(starts[e, k])] = k + 1
            gl[tuple(goals[e, k])] = k + 1
        envs.append(primal.MAPFEnv(num_agents=wl["N"], observation_size=wl["F"], world0=world, goals0=gl))
    return envs


def joint_step(env, actions, n_agents):
    """One environment step through the reference's public methods; returns a small digest of the outputs."""
    acc = 0
    for i in range(n_agents):
        _, r, done, nxt, on_goal, _, valid = env._step((i + 1, int(actions[i])))
        acc += int(on_goal) + int(valid) + len(nxt)
    for i in range(n_agents):
        maps4, vec = env._observe(i + 1)
        acc += int(maps4[0].sum())
    return acc


def main():
    cfg = json.loads(sys.argv[1])
    from oracle import refload
    from mapf_marl_b200 import workloads
    primal = refload.load_primal()
    wl = cfg["wl"]
    envs = build_envs(primal, wl, cfg["env_lo"], cfg["n_envs"], cfg["world_seed"])
    env_ids = range(cfg["env_lo"], cfg["env_lo"] + cfg["n_envs"])
    N = wl["N"]
    sys.stdout.write("ready\n")
    sys.stdout.flush()
    sys.stdin.readline()
    acc = 0
    t0 = None
    for t in range(cfg["warmup"] + cfg["steps"]):
        if t == cfg["warmup"]:
            t0 = time.perf_counter()
        acts = workloads.hash_actions_np(cfg["action_seed"], env_ids, t, N)
        for e, env in enumerate(envs):
            acc += joint_step(env, acts[e], N)
    elapsed = time.perf_counter() - t0 if t0 is not None else 0.0
    sys.stdout.write(json.dumps({"elapsed": elapsed, "agent_steps": cfg["steps"] * cfg["n_envs"] * N,
                                 "checksum": acc}) + "\n")
    sys.stdout.flush()


if __name__ == "__main__":
    main()
