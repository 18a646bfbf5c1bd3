#!/usr/bin/env python
"""Per-phase critical path of one tile (cycles), from a debug build of the library:

    cd mapf_marl_b200/csrc && nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo \
        -std=c++17 -Xcompiler -fPIC -shared -DMAPF_PHASE_TIMING -o ../libmapf_b200_dbg.so mapf_kernels.cu mapf_capi.cu mapf_host_unpack.cpp
    python profiles/phase_probe.py          (on the GPU box)

The middle block of the grid records clock64() at every phase boundary (PHASE_MARK in mapf_kernels.cu)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MAPF_B200_LIB"] = os.path.join(ROOT, "mapf_marl_b200", "libmapf_b200_dbg.so")
from mapf_marl_b200 import _lib  # noqa: E402
from mapf_marl_b200.engine import MapfEngine  # noqa: E402

WANT = tuple(os.environ.get("PROBE_WANT", "reward,terminated,dones,avail").split(","))
NAMES = ["stage", "phase A", "phase B", "phase C", "avail + agent bitmap", "phase D + write-back", "obs start",
         "phase 1", "phase 2"]


def main():
    lib = _lib.load()
    from mapf_marl_b200.workloads import WORKLOADS, make_world
    T = 16
    for name in ("c2", "c3", "c4"):
        wl = WORKLOADS[name]
        E, N = wl["E"], wl["N"]
        obst, starts, goals = make_world(wl, E, 0)
        eng = MapfEngine(E, N, wl["H"], wl["W"], mode="primal", fov=wl["F"], shared_map=wl["warehouse"])
        eng.reset(obst, starts, goals)
        acts = torch.randint(0, 5, (T, E, N), device="cuda", dtype=torch.uint8)
        for roll in (False, True):
            for _ in range(5):
                if roll:
                    eng.rollout(acts, want=WANT)
                else:
                    eng.step_observe(acts[0], want=WANT)
            torch.cuda.synchronize()
            buf = (ctypes.c_longlong * 32)()
            lib.mapf_debug_phase_clocks(buf)
            c = [buf[i] for i in range(11)]
            c[0] = c[10]                      # start of the (last) step's staging pass
            print(name, "rollout (last of %d steps)" % T if roll else "single step",
                  "-- tile critical path of the middle block (SM cycles):")
            for i in range(9):
                print("   %-24s %8d" % (NAMES[i], c[i + 1] - c[i]))
            print("   %-24s %8d" % ("total", c[9] - c[0]))
            print("   inside phase D: per-env part %d, statistics atomics %d, write-back %d" % (
                buf[11] - c[5], buf[12] - buf[11], c[6] - buf[12]))
            if roll and buf[16]:
                print("   pipelined kernel, step role: sweep %d, wait for a free snapshot %d, snapshot + signal %d" % (
                    buf[17] - buf[16], buf[18] - buf[17], buf[19] - buf[18]))
                print("   pipelined kernel, observation role: wait for the snapshot %d, window phase %d, goal bits %d, "
                      "expansion %d" % (buf[21] - buf[20], buf[22] - buf[21], buf[23] - buf[22], buf[24] - buf[23]))


if __name__ == "__main__":
    main()
