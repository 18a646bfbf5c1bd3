import ctypes, json, mmap, os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import mapf_marl_b200
lib = ctypes.CDLL(mapf_marl_b200.build())
lib.mapf_unpack_pool_create.restype = ctypes.c_void_p
lib.mapf_unpack_pool_create.argtypes = [ctypes.c_int]
lib.mapf_unpack_pool_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
cells = 16384 * 32 * 484
bits = np.random.randint(0, 2 ** 31, cells // 32, dtype=np.int32)
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
def run(dst_addr, label):
    p = lib.mapf_unpack_pool_create(16)
    for _ in range(3): lib.mapf_unpack_pool_run(p, bits.ctypes.data, dst_addr, cells, 1)
    t0 = time.perf_counter()
    for _ in range(10): lib.mapf_unpack_pool_run(p, bits.ctypes.data, dst_addr, cells, 1)
    dt = (time.perf_counter() - t0) / 10
    print(label, "%.3f ms %.1f GB/s" % (dt * 1e3, cells / dt / 1e9))
a = np.empty(cells, np.uint8); a[:] = 0
run(a.ctypes.data, "numpy (4K pages?)")
size = (cells + (2 << 20) - 1) // (2 << 20) * (2 << 20)
m = mmap.mmap(-1, size + (2 << 20))
try:
    m.madvise(mmap.MADV_HUGEPAGE)
except Exception as ex:
    print("madvise failed", ex)
buf = np.frombuffer(m, dtype=np.uint8)
addr = buf.ctypes.data
al = (addr + (2 << 20) - 1) // (2 << 20) * (2 << 20)
b2 = buf[al - addr: al - addr + cells]; b2[:] = 0
run(b2.ctypes.data, "mmap + MADV_HUGEPAGE")
for ln in open("/proc/meminfo"):
    if "AnonHugePages" in ln or "HugePages_Total" in ln: print(ln.strip())
