import os, sys, torch
sys.path.insert(0, "/root/repo")
from mapf_marl_b200 import maps
from mapf_marl_b200.engine import MapfEngine
def timed(fn, steps=200, warmup=10):
    for t in range(warmup): fn(t)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for t in range(steps): fn(t)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / steps
E, N, H, W = 16384, 32, 32, 32
obst, starts, goals = maps.synthetic_batch(1, E, H, W, 0.3, N, distinct=0)
for Wn, K in ((10, 4), (11, 5), (11, 6), (12, 4)):
    eng = MapfEngine(E, N, H, W, mode="partial", episode_limit=256, obs_window=Wn, obs_knn_agents=K)
    eng.reset(obst, starts, goals)
    osz = 2 * Wn * Wn + 13 * K
    m64 = timed(lambda t: eng.observe()); m32 = timed(lambda t: eng.observe(dtype=torch.float32))
    print("Wn %d K %d osz %d (mod4 %d): f64 %.4f ms %.0f GB/s | f32 %.4f ms %.0f GB/s" % (Wn, K, osz, osz % 4, m64, E*N*osz*8/m64/1e6, m32, E*N*osz*4/m32/1e6))
    eng.close()
