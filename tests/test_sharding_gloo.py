"""The N>1 path on CPU: world_size-2 gloo process group, environments sharded by index, the statistics
all-reduce, and G-invariance of the synthetic inputs (each rank's slice == the slice of the global batch)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, E, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mapf_marl_b200 import maps
    from mapf_marl_b200._lib import STAT_NAMES
    from mapf_marl_b200.sharding import reduce_stats, shard_range
    from oracle import Oracle
    from oracle.oracle import MODE_PRIMAL
    lo, hi = shard_range(E, rank, world)
    obst, starts, goals = maps.synthetic_batch(77, hi - lo, 12, 12, 0.2, 5, env_offset=lo, distinct=0)
    # step the shard with the CPU oracle (stands in for the device here) and reduce per-rank statistics
    orc = Oracle(hi - lo, 5, 12, 12, MODE_PRIMAL, fov=5, threads=1)
    orc.reset(obst, starts, goals)
    moved = 0
    for t in range(4):
        acts = np.random.RandomState(1000 * t).randint(0, 5, (E, 5)).astype(np.uint8)[lo:hi]   # global action tensor
        out = orc.primal_sweep(acts)
        moved += int((out["status"] >= 0).sum())
    stats = {k: 0 for k in STAT_NAMES}
    stats["env_steps"] = 4 * (hi - lo)
    stats["agent_steps"] = 4 * (hi - lo) * 5
    stats["goal_arrivals"] = moved
    total = reduce_stats(stats, torch.device("cpu"))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, obst=obst, starts=starts, goals=goals,
             pos=orc.positions(), total=np.array([total[k] for k in STAT_NAMES]), moved=moved)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_is_invariant_and_stats_reduce(tmp_path):
    from mapf_marl_b200 import maps
    from mapf_marl_b200._lib import STAT_NAMES
    from mapf_marl_b200.sharding import shard_range
    from oracle import Oracle
    from oracle.oracle import MODE_PRIMAL
    E, world = 13, 2
    mp.spawn(_worker, args=(world, _free_port(), E, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    assert [int(p["lo"]) for p in parts] == [0, 7] and [int(p["hi"]) for p in parts] == [7, 13]
    obst, starts, goals = maps.synthetic_batch(77, E, 12, 12, 0.2, 5, distinct=0)
    assert np.array_equal(np.concatenate([p["obst"] for p in parts]), obst)
    assert np.array_equal(np.concatenate([p["starts"] for p in parts]), starts)
    assert np.array_equal(np.concatenate([p["goals"] for p in parts]), goals)
    # a single-rank run over the global batch gives the same positions as the two shards
    orc = Oracle(E, 5, 12, 12, MODE_PRIMAL, fov=5, threads=1)
    orc.reset(obst, starts, goals)
    moved = 0
    for t in range(4):
        out = orc.primal_sweep(np.random.RandomState(1000 * t).randint(0, 5, (E, 5)).astype(np.uint8))
        moved += int((out["status"] >= 0).sum())
    assert np.array_equal(np.concatenate([p["pos"] for p in parts]), orc.positions())
    tot = dict(zip(STAT_NAMES, parts[0]["total"].tolist()))
    assert tot == dict(zip(STAT_NAMES, parts[1]["total"].tolist()))
    assert tot["env_steps"] == 4 * E and tot["agent_steps"] == 4 * E * 5 and tot["goal_arrivals"] == moved
    assert sum(int(p["moved"]) for p in parts) == moved


def test_shard_range_covers_everything():
    from mapf_marl_b200.sharding import shard_range
    for E in (1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            spans = [shard_range(E, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == E
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
