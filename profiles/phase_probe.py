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
import bench  # noqa: E402
from mapf_marl_b200 import _lib  # noqa: E402
from mapf_marl_b200.engine import MapfEngine  # noqa: E402

NAMES = ["stage", "phase A", "phase B", "phase C", "avail + agent bitmap", "phase D + write-back", "obs start",
         "phase 1", "phase 2"]


def main():
    lib = _lib.load()
    for name in ("c2", "c3", "c4"):
        wl = bench.WORKLOADS[name]
        E, N = wl["E"], wl["N"]
        obst, starts, goals = bench.make_world(wl, E, 0)
        eng = MapfEngine(E, N, wl["H"], wl["W"], mode="primal", fov=wl["F"], shared_map=wl["warehouse"])
        eng.reset(obst, starts, goals)
        acts = torch.randint(0, 5, (E, N), device="cuda", dtype=torch.uint8)
        for _ in range(5):
            eng.step_observe(acts)
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 32)()
        lib.mapf_debug_phase_clocks(buf)
        c = [buf[i] for i in range(10)]
        print(name, "tile critical path of the middle block (SM cycles):")
        for i in range(9):
            print("   %-24s %8d" % (NAMES[i], c[i + 1] - c[i]))
        print("   %-24s %8d" % ("total", c[9] - c[0]))


if __name__ == "__main__":
    main()
