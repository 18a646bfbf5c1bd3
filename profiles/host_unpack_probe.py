#!/usr/bin/env python
"""Host half of the e2e transport alone (no GPU work): bandwidth of the unpack pool expanding the c3 observation
(32 MB of bits -> 254 MB of uint8) into a pinned buffer, by thread count.  Gives the floor of mapf_step_observe_host."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mapf_marl_b200  # noqa: E402

lib = ctypes.CDLL(mapf_marl_b200.build())
lib.mapf_unpack_pool_create.restype = ctypes.c_void_p
lib.mapf_unpack_pool_create.argtypes = [ctypes.c_int]
lib.mapf_unpack_pool_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
lib.mapf_unpack_pool_destroy.argtypes = [ctypes.c_void_p]
cells = 16384 * 32 * 484
pin = torch.cuda.is_available()
bits = torch.randint(0, 2 ** 31, (cells // 32,), dtype=torch.int32)
out = torch.empty(cells, dtype=torch.uint8)
if pin:
    bits, out = bits.pin_memory(), out.pin_memory()
res = {}
for th in (4, 8, 12, 15, 16, 24):
    p = lib.mapf_unpack_pool_create(th)
    for _ in range(3):
        lib.mapf_unpack_pool_run(p, bits.data_ptr(), out.data_ptr(), cells, 1)
    t0 = time.perf_counter()
    for _ in range(10):
        lib.mapf_unpack_pool_run(p, bits.data_ptr(), out.data_ptr(), cells, 1)
    dt = (time.perf_counter() - t0) / 10
    res[th] = {"ms": dt * 1e3, "GBps_written": cells / dt / 1e9}
    lib.mapf_unpack_pool_destroy(p)
print(json.dumps({"pinned": pin, "cells": cells, "by_threads": res}))
