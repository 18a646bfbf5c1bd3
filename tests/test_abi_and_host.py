"""CPU-side tests: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/mapf_b200.h declares; struct layouts agree between the header and the ctypes binding; the host
logic (map ingestion, synthetic worlds, LUT, sharding arithmetic) behaves.  No kernel is launched."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mapf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(mapf_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_builds_loads_and_exports_every_declared_symbol():
    import mapf_marl_b200
    from mapf_marl_b200 import _lib
    path = mapf_marl_b200.build()
    assert os.path.exists(path)
    lib = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libmapf_b200.so lacks %s" % name
    assert set(declared) == set(_lib.PROTOTYPES), set(declared) ^ set(_lib.PROTOTYPES)
    assert lib.mapf_abi_version() == _lib.ABI_VERSION
    assert lib.mapf_build_arch() == b"sm_100a"


def test_library_has_no_torch_dependency_and_carries_sm100a_code():
    from mapf_marl_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "torch" not in out and "c10" not in out
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
        assert "sm_100a" in out


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    from mapf_marl_b200 import _lib
    prog = tmp_path / "layout.c"
    fields_cfg = [f[0] for f in _lib.MapfCfg._fields_]
    fields_out = [f[0] for f in _lib.MapfStepOut._fields_]
    fields_io = [f[0] for f in _lib.MapfHostIO._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mapf_b200.h"', 'int main(void){']
    lines.append('printf("%zu %zu %zu\\n", sizeof(mapf_cfg), sizeof(mapf_step_out), sizeof(mapf_host_io));')
    for f in fields_cfg:
        lines.append('printf("%%zu\\n", offsetof(mapf_cfg, %s));' % f)
    for f in fields_out:
        lines.append('printf("%%zu\\n", offsetof(mapf_step_out, %s));' % f)
    for f in fields_io:
        lines.append('printf("%%zu\\n", offsetof(mapf_host_io, %s));' % f)
    lines.append('return 0;}')
    prog.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True).stdout.split()
    sizes = [int(x) for x in out[:3]]
    assert sizes == [ctypes.sizeof(_lib.MapfCfg), ctypes.sizeof(_lib.MapfStepOut), ctypes.sizeof(_lib.MapfHostIO)]
    offs = [int(x) for x in out[3:]]
    expect = [getattr(_lib.MapfCfg, f).offset for f in fields_cfg] + \
             [getattr(_lib.MapfStepOut, f).offset for f in fields_out] + \
             [getattr(_lib.MapfHostIO, f).offset for f in fields_io]
    assert offs == expect


def test_default_cfg_carries_the_reference_defaults():
    from mapf_marl_b200 import _lib
    lib = _lib.load()
    cfg = _lib.MapfCfg()
    lib.mapf_default_cfg(ctypes.byref(cfg))
    assert (cfg.n_agents, cfg.episode_limit, cfg.step_reward, cfg.collide_reward) == (4, 10000, -0.01, -10.0)
    assert (cfg.action_cost, cfg.idle_cost, cfg.goal_reward, cfg.collision_reward) == (-0.3, -0.5, 0.0, -2.0)
    assert cfg.fov == 10 and cfg.collide_reward_is_int == 1 and cfg.step_reward_is_int == 0


def test_engine_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mapf_marl_b200.engine import MapfEngine, MapfError
    with pytest.raises(MapfError):
        MapfEngine(2, 2, 4, 4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mapf_marl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("test oracle", ""), os.path.join(dirpath, f)


def test_movingai_ingestion(tmp_path):
    from mapf_marl_b200 import maps
    mp = tmp_path / "a.map"
    mp.write_text("type octile\nheight 3\nwidth 4\nmap\n..@.\n.T..\n@@..\n")
    m = maps.read_movingai_map(str(mp))
    assert m.tolist() == [[0, 0, 1, 0], [0, 1, 0, 0], [1, 1, 0, 0]]
    sc = tmp_path / "a-1.scen"
    sc.write_text("version 1\n0\ta.map\t4\t3\t0\t1\t3\t2\t4.0\n5\ta.map\t4\t3\t3\t0\t0\t0\t3.0\n")
    lines = maps.read_scen_lines(str(sc))
    assert [maps.scen_fields(ln) for ln in lines] == [(0, 1, 3, 2), (3, 0, 0, 0)]


def test_synthetic_batch_is_a_function_of_the_global_env_index():
    from mapf_marl_b200 import maps
    full = maps.synthetic_batch(9, 12, 10, 10, 0.2, 4, distinct=0)
    a = maps.synthetic_batch(9, 6, 10, 10, 0.2, 4, env_offset=0, distinct=0)
    b = maps.synthetic_batch(9, 6, 10, 10, 0.2, 4, env_offset=6, distinct=0)
    for k in range(3):
        assert np.array_equal(full[k], np.concatenate([a[k], b[k]]))
    obst, starts, goals = full
    for e in range(12):
        assert len({tuple(p) for p in starts[e].tolist()}) == 4 and len({tuple(p) for p in goals[e].tolist()}) == 4
        lab = maps.label_components(obst[e])
        for k in range(4):
            assert obst[e][tuple(starts[e, k])] == 0 and obst[e][tuple(goals[e, k])] == 0
            assert lab[tuple(starts[e, k])] == lab[tuple(goals[e, k])]


def test_magnitude_lut_uses_the_reference_expression():
    from mapf_marl_b200.engine import magnitude_lut
    lut = magnitude_lut(32, 32)
    assert len(lut) == 2 * 31 * 31 + 1
    for dx, dy in ((0, 0), (3, 4), (-7, 19), (31, -31), (1, 2)):
        assert lut[dx * dx + dy * dy] == (dx ** 2 + dy ** 2) ** .5


def test_highway_layout_alphabet(tmp_path):
    from mapf_marl_b200 import maps
    p = tmp_path / "highways.txt"
    p.write_text("height (n_rows): 3\nwidth (n_cols): 5\nHighway Map:\n@eeI@\n@n.s@\nXwwL`\n")
    m = maps.read_highway_map(str(p))
    assert m["obst"].tolist() == [[1, 0, 0, 0, 1], [1, 0, 0, 0, 1], [0, 0, 0, 0, 0]]
    assert m["lane"].tolist() == [[0, 3, 3, 0, 0], [0, 1, 0, 2, 0], [0, 4, 4, 0, 0]]
    assert m["lock"][2, 3] == 1 and m["junction"][0, 3] == 1 and m["junction"][2, 0] == 1
    with pytest.raises(ValueError):
        maps.parse_highway_rows(["@?"])
    w = maps.warehouse_layout(64, 64)
    assert w.shape == (64, 64) and w[0].sum() == 0 and 0.3 < w.mean() < 0.8


def test_multiagentenv_contract_is_enforced():
    from mapf_marl_b200.multiagentenv import CONTRACT, MultiAgentEnv
    from mapf_marl_b200.mapf_gridworld import MAPF_GRID
    from mapf_marl_b200.marl_partial import MARL_PARTIAL_ENV
    for cls in (MAPF_GRID, MARL_PARTIAL_ENV):
        assert issubclass(cls, MultiAgentEnv)
        for name in CONTRACT:
            assert getattr(cls, name) is not getattr(MultiAgentEnv, name)
    with pytest.raises(TypeError):
        class Broken(MultiAgentEnv):          # forgets most of the contract
            def reset(self):
                return None
    with pytest.raises(NotImplementedError):
        MultiAgentEnv().step(None)


def test_host_unpack_pool_expands_bits_exactly():
    """The host half of the bit-packed PCIe transport (csrc/mapf_host_unpack.cpp) needs no GPU: bit i of the stream
    becomes element i of the output, for uint8 and float32 cells, any alignment, partial last word."""
    import ctypes
    import numpy as np
    from mapf_marl_b200 import _lib
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.mapf_unpack_pool_create.restype = ctypes.c_void_p
    lib.mapf_unpack_pool_create.argtypes = [ctypes.c_int]
    lib.mapf_unpack_pool_destroy.argtypes = [ctypes.c_void_p]
    lib.mapf_unpack_pool_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rs = np.random.RandomState(0)
    for threads in (1, 3, 8):
        pool = lib.mapf_unpack_pool_create(threads)
        assert pool
        for cells in (1, 31, 32, 33, 1000, 4 * 121 * 8 * 37, 100003):
            words = rs.randint(0, 2 ** 32, (cells + 31) // 32, dtype=np.uint64).astype(np.uint32)
            want = np.unpackbits(words.view(np.uint8), bitorder="little")[:cells]
            for off in (0, 1, 32):
                raw = np.full(cells + off + 64, 9, np.uint8)
                lib.mapf_unpack_pool_run(pool, words.ctypes.data, raw[off:].ctypes.data, cells, 1)
                assert np.array_equal(raw[off:off + cells], want), (threads, cells, off)
                assert (raw[:off] == 9).all() and (raw[off + cells:] == 9).all()     # nothing outside the range
            for off in (0, 1):
                rawf = np.full(cells + off + 16, 9.0, np.float32)
                lib.mapf_unpack_pool_run(pool, words.ctypes.data, rawf[off:].ctypes.data, cells, 4)
                assert np.array_equal(rawf[off:off + cells], want.astype(np.float32)), (threads, cells, off)
                assert (rawf[:off] == 9).all() and (rawf[off + cells:] == 9).all()
        lib.mapf_unpack_pool_destroy(pool)


def test_host_unpack_pool_progressive_publish():
    """One job per host call: the words are published chunk by chunk (as they arrive over PCIe) while the workers are
    already claiming blocks; the result is the same as a one-shot expansion."""
    import ctypes
    import threading
    import time
    import numpy as np
    from mapf_marl_b200 import _lib
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.mapf_unpack_pool_create.restype = ctypes.c_void_p
    lib.mapf_unpack_pool_create.argtypes = [ctypes.c_int]
    lib.mapf_unpack_pool_destroy.argtypes = [ctypes.c_void_p]
    lib.mapf_unpack_pool_begin.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    lib.mapf_unpack_pool_publish.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.mapf_unpack_pool_finish.argtypes = [ctypes.c_void_p]
    rs = np.random.RandomState(1)
    cells = 1000003
    nwords = (cells + 31) // 32
    final = rs.randint(0, 2 ** 32, nwords, dtype=np.uint64).astype(np.uint32)
    want = np.unpackbits(final.view(np.uint8), bitorder="little")[:cells]
    pool = lib.mapf_unpack_pool_create(4)
    for rep in range(3):
        words = np.zeros(nwords, np.uint32)           # the staging buffer: garbage until a chunk "arrives"
        out = np.full(cells + 64, 7, np.uint8)
        lib.mapf_unpack_pool_begin(pool, words.ctypes.data, out.ctypes.data, cells, 1)
        step = nwords // 7 + 1
        for lo in range(0, nwords, step):
            time.sleep(0.002)
            hi = min(nwords, lo + step)
            words[lo:hi] = final[lo:hi]
            lib.mapf_unpack_pool_publish(pool, min(cells, hi * 32))
        lib.mapf_unpack_pool_finish(pool)
        assert np.array_equal(out[:cells], want)
        assert (out[cells:] == 7).all()
    lib.mapf_unpack_pool_destroy(pool)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU: exactly one JSON line
    on stdout with the contract's keys, the UNMODIFIED Python reference as `cpu_baseline` (kind "reference", one worker
    process per host core), the C oracle port beside it as `cpu_port`, zero bytes in `e2e`."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--workload", "c2", "--envs", "256"], stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent-steps/sec (step+obs)" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    cores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == cores
    assert d["cpu_baseline"]["value"] == d["value"] and "UNMODIFIED" in d["cpu_baseline"]["sample"]
    assert d["cpu_port"]["kind"] == "port" and d["cpu_port"]["value"] > d["value"]      # C beats Python
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_worker_steps_exactly_what_the_oracle_steps():
    """The CPU baseline's worker (oracle/ref_worker.py, the UNMODIFIED mapf_primal.MAPFEnv) and the C oracle -- hence
    the GPU arm, which is tested against the oracle -- advance the same worlds with the same counter-hash actions: the
    digest of the worker's outputs is recomputed from the oracle's outputs."""
    import numpy as np
    import pytest
    from mapf_marl_b200 import workloads
    from oracle import Oracle, refload
    from oracle.oracle import MODE_PRIMAL
    from oracle import ref_worker
    if not refload.available():
        pytest.skip("reference files not available (neither /root/reference nor oracle/_ref)")
    primal = refload.load_primal()
    wl = dict(workloads.WORKLOADS["c2"])
    n_envs, lo, T, N, F = 6, 37, 5, wl["N"], wl["F"]
    envs = ref_worker.build_envs(primal, wl, lo, n_envs, 1000)
    obst, starts, goals = workloads.make_world(wl, n_envs, lo)
    orc = Oracle(n_envs, N, wl["H"], wl["W"], MODE_PRIMAL, fov=F, threads=1)
    orc.reset(obst, starts, goals)
    for t in range(T):
        acts = workloads.hash_actions_np(1234, range(lo, lo + n_envs), t, N)
        got = sum(ref_worker.joint_step(env, acts[e], N) for e, env in enumerate(envs))
        out = orc.primal_sweep(acts)
        obs, _ = orc.primal_observe()
        want = int(out["dones"].sum()) + int(out["valid"].sum()) + int(out["next_mid"].sum()) + int(obs[:, :, 0].sum())
        # joint_step counts on_goal as _step returns it (mid-sweep == the agent's own final flag), valid, len(nextActions)
        assert got == want, (t, got, want)
        for e, env in enumerate(envs):
            assert [tuple(p) for p in env.getPositions()] == [tuple(int(v) for v in p) for p in orc.positions()[e]]


def test_host_unpack_expands_any_cell_range_of_a_bit_stream():
    """mapf_host_unpack (the lazy view over a MAPF_BITS host observation): any [first_cell, first_cell + n) range, word
    aligned or not, as uint8 or float32, touches exactly n output elements.  Pure host code: runs without a GPU."""
    import numpy as np
    from mapf_marl_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(0)
    bits = rs.randint(0, 2 ** 32, 2000, dtype=np.uint64).astype(np.uint32)
    cells = np.unpackbits(bits.view(np.uint8), bitorder="little")
    for first, n in [(0, 64000), (0, 0), (5, 100), (37, 3000), (64, 31), (31, 1), (96, 640), (1, 63999), (15488 * 3, 15488)]:
        for dt, npdt in ((_lib.U8, np.uint8), (_lib.F32, np.float32)):
            out = np.full(n + 8, 7, npdt)
            rc = lib.mapf_host_unpack(bits.ctypes.data, first, n, out.ctypes.data, dt)
            assert rc == 0
            assert np.array_equal(out[:n], cells[first:first + n].astype(npdt)) and (out[n:] == 7).all(), (first, n, dt)
    assert lib.mapf_host_unpack(bits.ctypes.data, 0, 32, bits.ctypes.data, _lib.I64) != 0      # unsupported dtype


def test_counter_hash_actions_and_checksum_agree_between_numpy_and_torch():
    """The inputs of every arm: workloads.hash_actions_np (CPU reference workers, oracle) == hash_actions_torch (the
    GPU test of mapf_random_actions closes the triangle); state_checksum_np == state_checksum_torch."""
    import numpy as np
    import torch
    from mapf_marl_b200 import workloads as w
    for seed, lo, t, n in ((1234, 0, 0, 32), (7, 917504, 19, 8), (2 ** 31 + 3, 5, 70001, 5)):
        a = w.hash_actions_np(seed, range(lo, lo + 97), t, n)
        assert a.dtype == np.uint8 and a.max() <= 4
        assert np.array_equal(a, w.hash_actions_torch(seed, lo, 97, t, n, "cpu").numpy())
        av = (np.random.RandomState(t % 1000).rand(97, n, 5) < 0.5).astype(np.uint8)
        av[..., 0] |= av.sum(-1) == 0
        m = w.hash_actions_np(seed, range(lo, lo + 97), t, n, avail=av)
        assert np.array_equal(m, w.hash_actions_torch(seed, lo, 97, t, n, "cpu", avail=torch.as_tensor(av)).numpy())
        assert (np.take_along_axis(av, m[..., None].astype(np.int64), -1) == 1).all()
    # a shard draws what the whole batch draws for its environments
    whole = w.hash_actions_np(5, range(0, 64), 3, 6)
    assert np.array_equal(whole[40:], w.hash_actions_np(5, range(40, 64), 3, 6))
    # roughly uniform
    big = w.hash_actions_np(1, range(20000), 0, 16)
    assert np.abs(np.bincount(big.ravel(), minlength=5) / big.size - 0.2).max() < 0.01
    x = np.random.RandomState(0).randint(-300, 300, (50, 7, 2)).astype(np.int16)
    y = np.random.RandomState(1).randint(0, 2 ** 31, 12345).astype(np.int32)
    assert w.state_checksum_np(x, y) == w.state_checksum_torch(torch.as_tensor(x), torch.as_tensor(y))
    assert w.state_checksum_np(x, y) != w.state_checksum_np(y, x)


def test_reference_staging_archive_roundtrip(tmp_path):
    """oracle/stage_ref.py: the archive that carries the reference's own files to the GPU box unpacks to files whose
    sha256 equal the manifest (and the originals, where /root/reference is present)."""
    import hashlib
    import json
    import os
    import pytest
    from oracle import stage_ref
    if stage_ref.stage() is None and not os.path.exists(stage_ref.ARCHIVE):
        pytest.skip("no reference tree and no staged archive")
    root = stage_ref.unpack(str(tmp_path))
    man = json.load(open(os.path.join(stage_ref.DST, "MANIFEST.json")))["files"]
    assert sorted(man) == sorted(stage_ref.FILES)
    for rel, sha in man.items():
        data = open(os.path.join(root, rel), "rb").read()
        assert hashlib.sha256(data).hexdigest() == sha
        src = os.path.join(stage_ref.SRC, rel)
        if os.path.exists(src):
            assert open(src, "rb").read() == data
    # no reference source is tracked by git: the staging directory is ignored
    ign = open(os.path.join(os.path.dirname(stage_ref.HERE), ".gitignore")).read()
    assert "oracle/_ref/" in ign


def test_split_sum_of_rewards_equals_the_interpreters_sum():
    """The tile kernel evaluates CPython >= 3.12's `sum(rewards)` (GRID:141, PARTIAL:310) in three pieces
    (mapf_kernels.cu: py_sum_head / py_sum_term / py_sum_tail): the accumulator chain, every item's compensation
    term computed independently from (accumulator before the item, item), and the chain that folds the terms -- a
    skipped term is stored as +0.0.  The same decomposition in Python doubles must reproduce the interpreter's builtin
    bit for bit, for int / float mixes, signed zeros, huge and tiny magnitudes, infinities."""
    if sys.version_info < (3, 12):
        pytest.skip("the compensated builtin sum exists from CPython 3.12 on")
    import struct

    def bits(x):
        return struct.pack("<d", float(x))

    def split_sum(items):
        acc, pre, first_float = 0.0, [], len(items)
        for i, x in enumerate(items):                      # head: one add per item
            pre.append(acc)
            acc = acc + float(x)
            if isinstance(x, float) and first_float == len(items):
                first_float = i
        terms = []
        for i, x in enumerate(items):                      # terms: independent of each other
            if isinstance(x, float) and i > first_float:
                a, xi = pre[i], float(x)
                t = a + xi
                big = abs(a) >= abs(xi)
                terms.append(((a if big else xi) - t) + (xi if big else a))
            else:
                terms.append(0.0)
        c = 0.0
        for t in terms:                                    # tail: folded in order
            c = c + t
        if first_float < len(items) and c != 0.0 and np.isfinite(c):
            acc = acc + c
        return acc

    rs = np.random.RandomState(11)
    pool_f = [-0.01, -10.0, -0.3, -2.0, 0.5, 1e16, -1e16, 1e-300, 3.3333333333333335, -0.0, 0.0, 1e308, -1e308]
    pool_i = [0, -1, -10, 3, 20, 10 ** 9]
    for trial in range(4000):
        n = int(rs.randint(1, 40))
        p_int = rs.rand()
        items = [pool_i[rs.randint(len(pool_i))] if rs.rand() < p_int else
                 (pool_f[rs.randint(len(pool_f))] if rs.rand() < 0.7 else float(rs.randn() * 10.0 ** rs.randint(-8, 9)))
                 for _ in range(n)]
        want = sum(items)
        got = split_sum(items)
        assert bits(want) == bits(got) or (np.isnan(float(want)) and np.isnan(got)), (items, want, got)
    assert bits(split_sum([-0.0, -0.0])) == bits(sum([-0.0, -0.0]))
    assert bits(split_sum([float("inf"), 1.0, -5.0])) == bits(sum([float("inf"), 1.0, -5.0]))


def test_unit_step_closer_reward_needs_no_division():
    """PARTIAL's shaping term (dist[old] - dist[new]) / episode_limit (marl_partial.py:229-234): the kernel replaces the
    division by +-(1.0 / limit) or +0.0 when the hop distance changed by -1, 0 or +1 (partial_phase_a).  IEEE division is
    sign-symmetric and 0 / L is +0.0, so the substitution is exact for every limit."""
    lim = np.arange(1, 200001, dtype=np.float64)
    inv = 1.0 / lim
    assert np.array_equal((np.float64(-1.0) / lim).view(np.uint64), (-inv).view(np.uint64))
    assert np.array_equal((np.float64(1.0) / lim).view(np.uint64), inv.view(np.uint64))
    assert np.array_equal((np.float64(0.0) / lim).view(np.uint64), np.zeros_like(lim).view(np.uint64))
    for L in (1, 3, 7, 256, 10 ** 6, 2 ** 31 - 1):          # Python's float division is the same IEEE operation
        assert (-1) / L == -(1.0 / L) and (1 / L) == 1.0 / float(L) and str(0 / L) == "0.0"
