import torch, numpy as np, sys
sys.path.insert(0,".")
from mapf_marl_b200.engine import MapfEngine
from mapf_marl_b200.workloads import WORKLOADS, make_world
for name in ("c3","c4"):
    wl=WORKLOADS[name]; E=wl["E"]
    o,s,g=make_world(wl,E,0)
    eng=MapfEngine(E,wl["N"],wl["H"],wl["W"],mode="primal",fov=11,shared_map=wl["warehouse"],goal_dist=True)
    eng.reset(o,s,g)
    for _ in range(3): eng.refresh_goal_dist()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(7):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); eng.refresh_goal_dist(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms=float(np.median(ts)); maps=E*wl["N"]
    per=2*wl["H"]*wl["W"]+wl["H"]*wl["W"]/wl["N"]*(0 if wl["warehouse"] else 1)
    print(name,"bfs ms %.4f"%ms,"maps/s %.3e"%(maps/ms*1e3),"frac %.3f"%(maps*per/ms/1e6/6540.2))
    eng.close()
