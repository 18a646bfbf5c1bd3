#!/usr/bin/env python
"""bench.py -- agent-steps/s (step + observation) of the batched MAPF engine on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

One "step" = one pass of the hot path over one batch: every agent of every environment is advanced one
environment step and receives its observation, action mask, reward and done flag (one fused kernel
launch).  Workloads (BASELINE.json configs):
    c2  20x20 map, density 0.2,  8 agents, FOV 11,  4096 envs per GPU
    c3  32x32 map, density 0.3, 32 agents, FOV 11, 16384 envs per GPU   <- default: the config the
        north-star target (>= 1e9 agent-steps/s on 8 GPUs, obs kernel >= 50 % of HBM roofline) is quoted on
    c4  64x64 warehouse layout (one shared map), 128 agents, FOV 11, 8192 envs per GPU
    c5  the c3 shape with 1 048 576 envs in TOTAL, split evenly over the GPUs (strong scaling)
Multi-GPU: environments shard by index, `--gpus N` ranks under torchrun each own the same number of
environments (weak scaling); there is no data-path collective, NCCL only reduces the statistics vector.

The JSON line carries: value (device-resident inputs), e2e (host buffers in / host buffers out through
mapf_step_observe_host), roofline of the fused kernel against the measured HBM peak, and cpu_baseline
(the CPU oracle port on this box's host cores, bounded sample).  `--impl reference` times that CPU port
alone (the reference itself is pure Python and does not travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (H, W, density, N, F, envs per GPU, shared warehouse map, algorithmic bytes per agent-step)
    "c2": dict(H=20, W=20, density=0.2, N=8, F=11, E=4096, warehouse=False, bytes_per_agent_step=578),
    "c3": dict(H=32, W=32, density=0.3, N=32, F=11, E=16384, warehouse=False, bytes_per_agent_step=559),
    "c4": dict(H=64, W=64, density=0.0, N=128, F=11, E=8192, warehouse=True, bytes_per_agent_step=526),
    # c5: the 1M-env sweep of the c3 shape; E is the TOTAL, split evenly over the GPUs (strong scaling)
    "c5": dict(H=32, W=32, density=0.3, N=32, F=11, E=1048576, warehouse=False, bytes_per_agent_step=559, total=True),
}
METRIC = "agent-steps/sec (step+obs)"
UNIT = "agent-steps/s"


def make_world(wl, n_envs, env_offset, seed=1000):
    from mapf_marl_b200 import maps
    if wl["warehouse"]:
        obst = maps.warehouse_layout(wl["H"], wl["W"])
        free = np.argwhere(obst == 0)
        starts = np.zeros((n_envs, wl["N"], 2), np.int16)
        goals = np.zeros((n_envs, wl["N"], 2), np.int16)
        base = {}
        for e in range(n_envs):
            k = (env_offset + e) % 64
            if k not in base:
                rs = np.random.RandomState(seed + k)
                base[k] = (free[rs.permutation(len(free))[:wl["N"]]], free[rs.permutation(len(free))[:wl["N"]]])
            starts[e], goals[e] = base[k]
        return obst, starts, goals
    return maps.synthetic_batch(seed, n_envs, wl["H"], wl["W"], wl["density"], wl["N"], env_offset=env_offset,
                                distinct=64)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and clock-event (throttle) reasons through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.bits, self.max_mhz, self._stop_evt = index, [], 0, None, threading.Event()
        self.err = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop_evt.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.bits |= int(get_reasons(h))
                self._stop_evt.wait(0.002)
        except Exception as ex:  # keep the benchmark alive; the record says why clocks are missing
            self.err = repr(ex)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": reasons, "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


def cpu_port_rate(wl, n_envs, steps, warmup, threads=0, want_obs=True):
    """The CPU oracle port stepping + observing `n_envs` environments of the workload; agent-steps/s."""
    from oracle import Oracle
    from oracle.oracle import MODE_PRIMAL
    if not threads:   # all host threads this process may use (torchrun pins OMP_NUM_THREADS=1, so ask the OS)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    obst, starts, goals = make_world(wl, n_envs, 0)
    orc = Oracle(n_envs, wl["N"], wl["H"], wl["W"], MODE_PRIMAL, fov=wl["F"], shared_map=wl["warehouse"],
                 threads=threads)
    orc.reset(obst, starts, goals)
    rs = np.random.RandomState(0)
    acts = rs.randint(0, 5, (4, n_envs, wl["N"])).astype(np.uint8)
    want = ("agent_reward", "dones", "avail", "terminated", "reward")
    obs = np.empty((n_envs, wl["N"], 4, wl["F"], wl["F"]), np.uint8)
    vec = np.empty((n_envs, wl["N"], 3), np.float64)
    import ctypes

    def one(t):
        orc.primal_sweep(acts[t % 4], want=want)
        if want_obs:
            orc._lib.oracle_primal_observe(orc._h, obs.ctypes.data_as(ctypes.c_void_p),
                                           vec.ctypes.data_as(ctypes.c_void_p))
    for t in range(warmup):
        one(t)
    t0 = time.perf_counter()
    for t in range(steps):
        one(t)
    dt = time.perf_counter() - t0
    return n_envs * wl["N"] * steps / dt, dt, threads


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    # each step is a bounded sample of the workload, sized so that K steps take about two minutes at most
    r0, _, _ = cpu_port_rate(wl, min(wl["E"], 1024), 3, 1)
    n_envs = int(min(wl["E"], 2048, max(16, r0 * 120.0 / (max(args.steps, 1) * wl["N"]))))
    rate, dt, cores = cpu_port_rate(wl, n_envs, args.steps, min(args.warmup, 10))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.workload, wl, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d envs x %d agents x %d steps of %s (C oracle port, OpenMP over envs, %d threads)"
                                   % (n_envs, wl["N"], args.steps, args.workload, cores)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python and cannot travel to the GPU box; this is the C oracle port, "
                "pinned bit-exact to the live reference by tests/test_oracle_golden.py",
    }
    emit(line)


def workload_config(name, wl, gpus):
    return {"workload": "%s: %dx%d map, density %.2f, %d agents, FOV %d, %d envs per GPU%s" % (
        name, wl["H"], wl["W"], wl["density"], wl["N"], wl["F"], wl["E"],
        ", shared warehouse map" if wl["warehouse"] else ""),
        "mode": "primal (sequential claim) + 4-channel FOV observation + goal vector, fused step+obs kernel",
        "n_envs_total": wl["E"] * gpus, "n_agents": wl["N"],
        "l2": "per-step working set (obs output %.0f MB) exceeds the 126 MB L2" % (
            wl["E"] * wl["N"] * 4 * wl["F"] ** 2 / 1e6),
        "parallelism": "envs sharded by index over %d GPU(s), no data-path collective" % gpus}


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--f32", action="store_true", help="emit float32 observations instead of uint8")
    ap.add_argument("--lean", action="store_true",
                    help="no repeat passes and no extra clock-sampling load: a fixed launch order for the ncu recipes")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: library banners (NCCL prints its version to stdout) are sent to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    wl = dict(WORKLOADS[args.workload])
    if args.envs:
        wl["E"] = args.envs
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    strong = bool(wl.get("total"))
    if strong:
        wl["E"] = wl["E"] // max(world, 1)        # the same total work on any number of GPUs
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        from oracle import build_oracle
        if rank == 0:
            build_oracle()
        run_reference_arm(args, wl, rank)
        return

    import torch
    import torch.distributed as dist
    import mapf_marl_b200
    from mapf_marl_b200.engine import MapfEngine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (use --impl reference for the CPU arm)")
    mapf_marl_b200.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, N, F = wl["E"], wl["N"], wl["F"]
    obst, starts, goals = make_world(wl, E, rank * E)
    eng = MapfEngine(E, N, wl["H"], wl["W"], mode="primal", fov=F, shared_map=wl["warehouse"], goal_dist=True,
                     device=dev)
    eng.reset(obst, starts, goals)
    # goal-distance maps: computed at reset (and on goal reassignment), reported separately
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.refresh_goal_dist()
    torch.cuda.synchronize()
    ev0.record()
    eng.refresh_goal_dist()
    ev1.record()
    torch.cuda.synchronize()
    bfs_ms = ev0.elapsed_time(ev1)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool = torch.randint(0, 5, (16, E, N), generator=gen, device=dev, dtype=torch.uint8)
    want = ("reward", "terminated", "dones", "avail")
    odt = torch.float32 if args.f32 else torch.uint8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for t in range(warmup):
            fn(t)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for t in range(steps):
            fn(t)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput (the fused step+observation kernel), clocks sampled meanwhile.
    #      The K timed steps are replayed from CUDA graphs of `chunk` consecutive steps each (what a captured rollout
    #      loop pays: no Python / ctypes cost per launch); eager launches are timed next to it.
    args.warmup = max(args.warmup, 3)          # timing rule: at least three untimed warm-up steps
    for t in range(args.warmup):
        eng.step_observe(pool[t % 16], want=want, dtype=odt)
    chunk = max(c for c in range(1, 21) if args.steps % c == 0)
    graph = None
    if chunk > 1 or args.steps == 1:
        try:
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                for t in range(chunk):
                    eng.step_observe(pool[t % 16], want=want, dtype=odt)
        except Exception as exc:   # capture is a convenience of the host library, never a reason to lose the bench line
            sys.stderr.write("CUDA graph capture failed, timing eager launches: %r\n" % (exc,))
            graph = None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    if graph is not None:
        ms_total = timed(lambda t: graph.replay(), args.steps // chunk, 2)
    else:
        ms_total = timed(lambda t: eng.step_observe(pool[t % 16], want=want, dtype=odt), args.steps, 0)
    launches = args.steps            # one tile-kernel launch per step (inside the graphs when replayed)
    ms_step = ms_total / args.steps
    # four more passes over the same K steps (informational: spread of the measurement; `value` is the first pass)
    repeats = [ms_step]
    for _ in range(0 if args.lean else 4):
        if graph is not None:
            repeats.append(timed(lambda t: graph.replay(), args.steps // chunk, 0) / args.steps)
        else:
            repeats.append(timed(lambda t: eng.step_observe(pool[t % 16], want=want, dtype=odt), args.steps, 0) / args.steps)
    value = world * E * N * args.steps / (ms_total * 1e-3)
    n_eager = max(min(args.steps, 2000), 1)
    ms_eager = timed(lambda t: eng.step_observe(pool[t % 16], want=want, dtype=odt), n_eager, 3) / n_eager
    # the clock sampler (2 ms period) keeps running through the eager leg of the same kernel; a very short run
    # (small --steps) is followed by more of the same load, untimed, until the sampler has seen about 0.3 s of it
    if sampler and not args.lean:
        t_load = time.perf_counter()
        while len(sampler.sm) < 100 and time.perf_counter() - t_load < 0.5:
            for t in range(50):
                eng.step_observe(pool[t % 16], want=want, dtype=odt)
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    # ---- the fused launch with the observation left bit-packed (MAPF_BITS, for consumers that take bits; informational)
    ms_bits = None
    if eng.bits_supported():
        ms_bits = timed(lambda t: eng.step_observe(pool[t % 16], want=want, dtype="bits"), n_eager, 3) / n_eager

    # ---- breakdown: the step-only and observe-only launches of the same tile kernel
    ms_obs = timed(lambda t: eng.observe(dtype=odt), max(args.steps // 4, 5), 3) / max(args.steps // 4, 5)
    ms_stp = timed(lambda t: eng.step(pool[t % 16], want=want), max(args.steps // 4, 5), 3) / max(args.steps // 4, 5)

    # ---- c4 only: the lifelong variant of the step (BASELINE config 4): after every fused launch, agents standing on
    #      their goal pop the next one from their queue (mapf_pop_goals) and the distance maps of exactly those goals are
    #      recomputed (mapf_bfs with the dirty mask): three launches per step, nothing leaves the device.
    lifelong = None
    if wl["warehouse"]:
        from mapf_marl_b200.lifelong import LifelongGoals
        Q = 8
        free = np.argwhere(obst == 0)
        qbase = {}
        queue = np.zeros((E, N, Q, 2), np.int16)
        for e in range(E):
            k = (rank * E + e) % 64
            if k not in qbase:
                # every queued goal is a cell no other agent ever has as a goal (PRIMAL's goals grid holds one id per cell)
                rs = np.random.RandomState(77000 + k)
                taken = set(map(tuple, goals[e].tolist()))
                cand = np.array([c for c in free[rs.permutation(len(free))].tolist() if tuple(c) not in taken], np.int16)
                qbase[k] = cand[:N * Q].reshape(N, Q, 2)
            queue[e] = qbase[k]
        eng.reset(obst, starts, goals)
        eng.refresh_goal_dist()
        life = LifelongGoals(eng, queue, overlap=True)
        n_life = max(min(args.steps, 1000), 1)

        def life_step(t):
            eng.step_observe(pool[t % 16], want=want, dtype=odt)
            life.reassign()
        ms_life = timed(life_step, n_life, 3) / n_life
        life.sync()
        life_serial = LifelongGoals(eng, queue)
        life_serial.head.copy_(life.head)

        def life_step_serial(t):
            eng.step_observe(pool[t % 16], want=want, dtype=odt)
            life_serial.reassign()
        ms_life_serial = timed(life_step_serial, n_life, 3) / n_life
        popped = int(life.head.sum().item())
        lifelong = {"ms_per_step": ms_life, "agent_steps_per_s": world * E * N / (ms_life * 1e-3),
                    "goal_queue_depth": Q, "reassignments_per_step": popped / (n_life + 3),
                    "launches_per_step": 4, "ms_per_step_bfs_on_the_same_stream": ms_life_serial,
                    "note": "fused step+obs, then mapf_pop_goals and mapf_bfs(dirty) = list compaction + BFS of the "
                            "agents that arrived, the BFS on a side stream under the next step's launch; random "
                            "actions, so arrivals are rare"}

    # ---- end to end: pinned host actions in, every output back on the host, through the C-ABI host entry point
    io, bufs, h2d, d2h = eng.make_host_io(obs_dtype=odt)
    host_pool = pool[:4].cpu()
    host_pool_np = host_pool.numpy()
    act_np = bufs["actions"].numpy()          # view of the pinned action buffer

    def e2e_step(t):
        # a plain memcpy into the pinned buffer (torch's CPU copy_ would fan out to its OpenMP pool, whose threads then
        # spin on every core for milliseconds and compete with the library's unpack threads)
        np.copyto(act_np, host_pool_np[t % 4])
        eng.step_observe_host(io)
    ms_e2e = timed(e2e_step, args.e2e_steps, 8)      # the first calls start the unpack pool and ramp the host clocks
    e2e_value = world * E * N * args.e2e_steps / (ms_e2e * 1e-3)
    e2e_transport = "bit-packed observation over PCIe, expanded to the requested dtype by the library's host threads" \
        if eng.host_transport() == 1 else "dense copies"
    # the same call with dense copies of the uint8 observation (informational)
    ms_e2e_dense = None
    if eng.host_transport() == 1:
        eng.host_transport(False)
        _, _, _, d2h_dense = eng.make_host_io(obs_dtype=odt)
        ms_e2e_dense = timed(e2e_step, args.e2e_steps, 2) / args.e2e_steps
        eng.host_transport(True)

    # the same call handing the bit stream itself to the host (MAPF_BITS host output, no expansion; informational)
    ms_e2e_bits = None
    if eng.bits_supported():
        io_b, bufs_b, _, d2h_bits = eng.make_host_io(obs_dtype="bits")
        act_b = bufs_b["actions"].numpy()

        def e2e_bits_step(t):
            np.copyto(act_b, host_pool_np[t % 4])
            eng.step_observe_host(io_b)
        ms_e2e_bits = timed(e2e_bits_step, args.e2e_steps, 2) / args.e2e_steps

    # ---- the same host entry point when the policy lives on the GPU (pymarl's controller does): actions come from
    #      the host, reward / terminated go back, the observation stays in device memory for the agent network
    acts_dev = torch.empty((E, N), dtype=torch.uint8, device=dev)
    pin_act = torch.empty((E, N), dtype=torch.uint8).pin_memory()
    pin_rew = torch.empty((E,), dtype=torch.float64).pin_memory()
    pin_term = torch.empty((E,), dtype=torch.uint8).pin_memory()
    h2d2, d2h2 = E * N, E * 9

    def e2e_dev_obs(t):
        pin_act.copy_(host_pool[t % 4])
        acts_dev.copy_(pin_act, non_blocking=True)
        out = eng.step_observe(acts_dev, want=want, dtype=odt)       # one fused launch, observation stays in HBM
        pin_rew.copy_(out["reward"], non_blocking=True)
        pin_term.copy_(out["terminated"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e2e2 = timed(e2e_dev_obs, max(args.e2e_steps * 10, 50), 3) / max(args.e2e_steps * 10, 50)

    # ---- statistics: the only collective of the path (one all-reduce of 8 int64 over NCCL)
    stats = eng.stats()
    svec = torch.tensor([stats[k] for k in sorted(stats)], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(svec, op=dist.ReduceOp.SUM)
    flags = eng.error_flags()

    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_per = wl["bytes_per_agent_step"] + (3 * 4 * F * F if args.f32 else 0)
        alg_bytes = bytes_per * E * N
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r1_fused_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(args.workload)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32 obs / u8 state" if args.f32 else "u8", "data": "synthetic",
            "config": dict(workload_config(args.workload, wl, world),
                           launch=("CUDA graph replay, %d steps per graph" % chunk) if graph is not None
                           else "eager launches (one C-ABI call per step)"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.e2e_steps, "steps": args.e2e_steps, "transport": e2e_transport},
            "e2e_dense_transport": None if ms_e2e_dense is None else {
                "value": world * E * N / (ms_e2e_dense * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_dense,
                "d2h_bytes_per_step": d2h_dense,
                "note": "informational: the same host call with the observation tensor copied densely over PCIe"},
            "e2e_bits_to_host": None if ms_e2e_bits is None else {
                "value": world * E * N / (ms_e2e_bits * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_bits,
                "d2h_bytes_per_step": d2h_bits,
                "note": "informational: the same host call with obs_dtype MAPF_BITS (the observation arrives in host "
                        "memory as a bit stream, one bit per cell; no host expansion)"},
            "e2e_obs_on_device": {"value": world * E * N / (ms_e2e2 * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e2,
                                  "h2d_bytes_per_step": h2d2, "d2h_bytes_per_step": d2h2,
                                  "note": "informational: host actions in, reward/terminated out, observation left "
                                          "in HBM for a GPU-resident policy (one fused launch per step)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "mapf_tile_kernel<%d> (fused step+obs)" % F,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes_per_agent_step": bytes_per},
            "breakdown_ms": {"fused_step_obs": ms_step, "fused_step_obs_5_passes": repeats,
                             "fused_step_obs_median_of_5": float(np.median(repeats)),
                             "fused_step_obs_eager_launches": ms_eager,
                             "fused_step_obs_bit_packed_output": ms_bits,
                             "observe_only": ms_obs, "step_only": ms_stp,
                             "observe_only_GBps": (4 * F * F * (4 if args.f32 else 1) + 24 + 8 +
                                                   wl["H"] * wl["W"] / N) * E * N / (ms_obs * 1e-3) / 1e9,
                             "goal_bfs_all_maps": bfs_ms, "goal_maps_per_s": E * N / (bfs_ms * 1e-3)},
            "lifelong": lifelong,
            "clocks": clocks,
            "stats": dict(zip(sorted(stats), [int(v) for v in svec.tolist()])),
            "device_error_flags": flags,
        }
        if not args.no_cpu:
            n_cpu = min(E, 2048)
            r0, dt0, cores = cpu_port_rate(wl, n_cpu, 4, 1)
            n_steps = int(max(8, min(20000, 10.0 / max(dt0 / 4, 1e-6))))      # about 10 s of CPU work
            rate, dt, cores = cpu_port_rate(wl, n_cpu, n_steps, 2)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d envs x %d agents x %d steps of %s (step+obs), C oracle port, OpenMP over envs on "
                          "%d threads, %.1f s" % (n_cpu, N, n_steps, args.workload, cores, dt)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
