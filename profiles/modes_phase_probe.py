#!/usr/bin/env python
"""Per-phase critical path of one GRID / PARTIAL tile (cycles) on a c3-shaped batch, from the debug build of the
library (see phase_probe.py for the build line):   python profiles/modes_phase_probe.py   (on the GPU box)"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MAPF_B200_LIB", os.path.join(ROOT, "mapf_marl_b200", "libmapf_b200_dbg.so"))
from mapf_marl_b200 import _lib  # noqa: E402
from mapf_marl_b200.engine import MapfEngine  # noqa: E402
from mapf_marl_b200.workloads import WORKLOADS, make_world  # noqa: E402

WANT = ("reward", "terminated", "dones", "avail")
NAMES = ["stage", "phase A (+ moves)", "phase B", "phase C", "avail + agent bitmap", "phase D + write-back",
         "obs start", "phase 1", "phase 2 / full-map obs"]


def main():
    lib = _lib.load()
    wl = WORKLOADS["c3"]
    E = int(os.environ.get("PROBE_ENVS", wl["E"]))
    N = wl["N"]
    obst, starts, goals = make_world(wl, E, 0)
    for mname, kw in (("grid", dict(mode="grid", episode_limit=10 ** 6)),
                      ("partial", dict(mode="partial", episode_limit=256, obs_window=11, obs_knn_agents=5))):
        eng = MapfEngine(E, N, wl["H"], wl["W"], device="cuda:0", **kw)
        eng.reset(obst, starts, goals)
        acts = torch.randint(0, 5, (16, E, N), device="cuda", dtype=torch.uint8)
        for t in range(5):
            eng.step_observe(acts[t], want=WANT)
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 32)()
        lib.mapf_debug_phase_clocks(buf)
        c = [buf[i] for i in range(11)]
        c[0] = c[10]
        print(mname, "single step -- tile critical path of the middle block (SM cycles):")
        for i in range(9):
            print("   %-24s %8d" % (NAMES[i], c[i + 1] - c[i]))
        print("   %-24s %8d" % ("total", c[9] - c[0]))
        print("   inside phase D: per-env part %d, statistics atomics %d, write-back %d" % (
            buf[11] - c[5], buf[12] - buf[11], c[6] - buf[12]))
        print("   per-env part: loop incl. accumulator chain %d, barrier %d, terms + barrier %d, term chain %d" % (
            buf[13] - c[5], buf[14] - buf[13], buf[15] - buf[14], buf[11] - buf[15]))
        eng.close()


if __name__ == "__main__":
    main()
