import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from mapf_marl_b200.workloads import WORKLOADS, make_world
from mapf_marl_b200.vec_env import PrimalVecEnv
from mapf_marl_b200.batched_runner import BatchedRunner, RandomMAC
wl = WORKLOADS["c3"]; E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072; T = 8
o, s, g = make_world(wl, 4096, 0)
reps = (E + 4095) // 4096
o, s, g = np.tile(o, (reps, 1, 1))[:E], np.tile(s, (reps, 1, 1))[:E], np.tile(g, (reps, 1, 1))[:E]
env = PrimalVecEnv(o, s, g, fov=11, episode_limit=T)
r = BatchedRunner(env, RandomMAC(env.engine, seed=1), check_every=8, cuda_graph=True)
r.run(); r.run()
def tm(fn, n=5):
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
print("run() ms", tm(r.run))
print("reset() ms", tm(lambda: r.reset(True)))
print("graph replay ms", tm(lambda: r._graphs[0].replay()))
acts = torch.zeros((E, 32), dtype=torch.int64, device="cuda")
print("step_into x8 ms", tm(lambda: [env.step_into(acts, t, r.batch) for t in range(8)]))
print("random_actions x8 ms", tm(lambda: [env.engine.random_actions(1, t, avail=r.batch.tm["avail_actions"][t]) for t in range(8)]))
