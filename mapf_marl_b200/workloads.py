"""Synthetic inputs of the BASELINE.json configurations (host side; shared by bench.py, the tests and the CPU
reference workers so that every arm steps the SAME worlds with the SAME actions).

Everything is a function of the GLOBAL environment index: a shard [lo, hi) of a batch sees exactly the worlds and
actions the unsharded batch would give those environments (results must not depend on the number
of GPUs).
"""
import numpy as np

from . import maps

WORKLOADS = {
    # name: map shape, obstacle density, agents, FOV side, envs per GPU, shared warehouse map, algorithmic bytes per
    # agent-step
    "c2": dict(H=20, W=20, density=0.2, N=8, F=11, E=4096, warehouse=False, bytes_per_agent_step=578),
    "c3": dict(H=32, W=32, density=0.3, N=32, F=11, E=16384, warehouse=False, bytes_per_agent_step=559),
    "c4": dict(H=64, W=64, density=0.0, N=128, F=11, E=8192, warehouse=True, bytes_per_agent_step=526),
    # c5: the 1M-env sweep of the c3 shape; E is the TOTAL, split evenly over the GPUs (strong scaling)
    "c5": dict(H=32, W=32, density=0.3, N=32, F=11, E=1048576, warehouse=False, bytes_per_agent_step=559, total=True),
}

DISTINCT_WORLDS = 4096      # worlds cycle with this period over the global env index (host generation cost)


def make_world(wl, n_envs, env_offset, seed=1000, distinct=DISTINCT_WORLDS):
    """(obst, starts, goals) of environments [env_offset, env_offset + n_envs) of workload `wl`.
    obst: uint8 [n_envs,H,W], or [H,W] for the shared warehouse map; starts / goals: int16 [n_envs,N,2]."""
    if wl["warehouse"]:
        obst = maps.warehouse_layout(wl["H"], wl["W"])
        free = np.argwhere(obst == 0)
        starts = np.zeros((n_envs, wl["N"], 2), np.int16)
        goals = np.zeros((n_envs, wl["N"], 2), np.int16)
        base = {}
        for e in range(n_envs):
            k = (env_offset + e) % distinct if distinct else env_offset + e
            if k not in base:
                rs = np.random.RandomState(seed + k)
                base[k] = (free[rs.permutation(len(free))[:wl["N"]]], free[rs.permutation(len(free))[:wl["N"]]])
            starts[e], goals[e] = base[k]
        return obst, starts, goals
    return maps.synthetic_batch(seed, n_envs, wl["H"], wl["W"], wl["density"], wl["N"], env_offset=env_offset,
                                distinct=distinct)


def make_goal_queue(wl, obst, goals, n_envs, env_offset, depth=8, seed=77000, distinct=DISTINCT_WORLDS):
    """Lifelong goal queues int16 [n_envs,N,depth,2] for the warehouse workload: every queued goal is a free cell
    that no agent of that environment ever has as a goal (PRIMAL's goals grid holds one id per cell)."""
    free = np.argwhere(obst == 0)
    N = wl["N"]
    queue = np.zeros((n_envs, N, depth, 2), np.int16)
    base = {}
    for e in range(n_envs):
        k = (env_offset + e) % distinct if distinct else env_offset + e
        if k not in base:
            rs = np.random.RandomState(seed + k)
            taken = set(map(tuple, goals[e].tolist()))
            cand = np.array([c for c in free[rs.permutation(len(free))].tolist() if tuple(c) not in taken], np.int16)
            base[k] = cand[:N * depth].reshape(N, depth, 2)
        queue[e] = base[k]
    return queue


# ------------------------------------------------------------------------------------------------------------------
# Counter-based actions: a pure function of (seed, GLOBAL env index, step, agent).  The numpy and the torch version
# compute the same 32-bit mix, so the CPU arms and every GPU shard draw identical actions.
# ------------------------------------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF


def _mix_np(x):
    x = x & _M32
    x ^= x >> 15
    x = (x * 0x2C1B3C6D) & _M32
    x ^= x >> 12
    x = (x * 0x297A2D39) & _M32
    x ^= x >> 15
    return x


def hash_u32_np(seed, env_ids, t, n_agents):
    """uint32-valued int64 [len(env_ids), n_agents]."""
    e = np.asarray(env_ids, np.int64).reshape(-1, 1)
    a = np.arange(n_agents, dtype=np.int64).reshape(1, -1)
    x = (e * 0x9E3779B1 + int(t) * 0x85EBCA77 + a * 0xC2B2AE3D + int(seed) * 0x27D4EB2F) & _M32
    return _mix_np(_mix_np(x) + 0x165667B1)


def hash_actions_np(seed, env_ids, t, n_agents, n_actions=5, avail=None):
    """Uniform actions in [0, n_actions) -- or, with avail uint8 [E,N,n_actions], uniform over the set bits of every
    agent's mask (the r-th available action, r = hash mod popcount)."""
    h = hash_u32_np(seed, env_ids, t, n_agents)
    if avail is None:
        return (((h >> 8) * n_actions) >> 24).astype(np.uint8)
    av = np.asarray(avail) != 0
    cnt = av.sum(-1)
    r = (h >> 8) % np.maximum(cnt, 1)
    pick = (np.cumsum(av, -1) == (r + 1)[..., None]) & av
    return np.where(cnt > 0, pick.argmax(-1), 0).astype(np.uint8)


def hash_actions_torch(seed, env_lo, n_envs, t, n_agents, device, n_actions=5, avail=None):
    """The same values as hash_actions_np(seed, range(env_lo, env_lo + n_envs), ...) as a uint8 device tensor."""
    import torch
    e = torch.arange(env_lo, env_lo + n_envs, dtype=torch.int64, device=device).reshape(-1, 1)
    a = torch.arange(n_agents, dtype=torch.int64, device=device).reshape(1, -1)

    def mix(x):
        x = x & _M32
        x = x ^ (x >> 15)
        x = (x * 0x2C1B3C6D) & _M32
        x = x ^ (x >> 12)
        x = (x * 0x297A2D39) & _M32
        return x ^ (x >> 15)
    x = (e * 0x9E3779B1 + int(t) * 0x85EBCA77 + a * 0xC2B2AE3D + int(seed) * 0x27D4EB2F) & _M32
    h = mix(mix(x) + 0x165667B1)
    if avail is None:
        return (((h >> 8) * n_actions) >> 24).to(torch.uint8)
    av = avail != 0
    cnt = av.sum(-1)
    r = (h >> 8) % cnt.clamp(min=1)
    pick = (av.cumsum(-1) == (r + 1).unsqueeze(-1)) & av
    return torch.where(cnt > 0, pick.to(torch.uint8).argmax(-1), torch.zeros_like(cnt)).to(torch.uint8)


_CK_A, _CK_B, _CK_MUL = 0x9E3779B97F4A7C15 - (1 << 64), 0x7F4A7C15, 1000003     # int64 two's-complement constants


def state_checksum_np(*arrays):
    """Index-weighted 64-bit checksum of integer arrays (positions, packed observation words, ...), wrap-around int64
    arithmetic: the value bench.py prints as rank0_state_checksum.  It must not depend on how many GPUs shared the
    batch.  state_checksum_torch computes the same number on the device."""
    acc = np.int64(0)
    with np.errstate(over="ignore"):
        for arr in arrays:
            v = np.ascontiguousarray(arr).view(np.uint8).astype(np.int64).ravel()
            idx = np.arange(v.size, dtype=np.int64)
            part = ((v + np.int64(1)) * (idx * np.int64(_CK_A) + np.int64(_CK_B))).sum(dtype=np.int64)
            acc = acc * np.int64(_CK_MUL) + part
    return int(acc) & 0xFFFFFFFFFFFFFFFF


def state_checksum_torch(*tensors):
    import torch
    acc = None
    for t in tensors:
        v = t.contiguous().view(torch.uint8).reshape(-1).to(torch.int64)
        idx = torch.arange(v.numel(), dtype=torch.int64, device=v.device)
        part = ((v + 1) * (idx * _CK_A + _CK_B)).sum()
        acc = part if acc is None else acc * _CK_MUL + part
    return int(acc.item()) & 0xFFFFFFFFFFFFFFFF
