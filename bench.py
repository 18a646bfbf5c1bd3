#!/usr/bin/env python
"""bench.py -- agent-steps/s (step + observation) of the batched MAPF engine on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

One "step" = one pass of the hot path over one batch: every agent of every environment is advanced one
environment step and receives its observation, action mask, reward and done flag (one fused kernel
launch).  Workloads (BASELINE.json configs, mapf_marl_b200/workloads.py):
    c2  20x20 map, density 0.2,  8 agents, FOV 11,  4096 envs per GPU
    c3  32x32 map, density 0.3, 32 agents, FOV 11, 16384 envs per GPU   <- default: the config the
        north-star target (>= 1e9 agent-steps/s on 8 GPUs, obs kernel >= 50 % of HBM roofline) is quoted on
    c4  64x64 warehouse layout (one shared map), 128 agents, FOV 11, 8192 envs per GPU
    c5  the c3 shape with 1 048 576 envs in TOTAL, split evenly over the GPUs (strong scaling)
Multi-GPU: environments shard by index, `--gpus N` ranks under torchrun each own the same number of
environments (weak scaling); there is no data-path collective, NCCL only reduces the statistics vector.
Worlds AND actions are functions of the GLOBAL environment index (counter hash of seed, env, step, agent), so rank 0
computes the same thing on any number of GPUs: `rank0_state_checksum` must be identical in every line.

The JSON line carries: value (device-resident inputs; median of >= 5 timed passes of exactly K steps, >= 0.2 s timed in
total), e2e (host buffers in / host buffers out through mapf_step_observe_host) with its variants and a transfer-limit
model, roofline of the fused kernel against the measured HBM peak, `sweep` (every BASELINE config and an E-sweep of
c3), `rollout` (the reference-API rollout loop: BatchedRunner + controller), and cpu_baseline: the UNMODIFIED Python
reference env stepped by one worker process per host core (kind "reference"; the C oracle port is reported beside it
as cpu_port).  `--impl reference` times that Python reference alone.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mapf_marl_b200.workloads import WORKLOADS, make_goal_queue, make_world  # noqa: E402  (numpy only)

METRIC = "agent-steps/sec (step+obs)"
UNIT = "agent-steps/s"
ACTION_SEED = 1234
WANT = ("reward", "terminated", "dones", "avail")


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


# ------------------------------------------------------------------------------------------------------------------
# CPU arms: the unmodified Python reference (oracle/ref_pool.py) and the C oracle port.  Both step the SAME worlds with
# the SAME counter-hash actions as the GPU arm.
# ------------------------------------------------------------------------------------------------------------------
def cpu_port_rate(wl, n_envs, min_seconds, threads=0):
    """The C oracle port (OpenMP over envs) stepping + observing `n_envs` environments of the workload."""
    import ctypes
    from mapf_marl_b200 import workloads
    from oracle import Oracle, build_oracle
    from oracle.oracle import MODE_PRIMAL
    build_oracle()
    if not threads:   # all host threads this process may use (torchrun pins OMP_NUM_THREADS=1, so ask the OS)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    obst, starts, goals = make_world(wl, n_envs, 0)
    orc = Oracle(n_envs, wl["N"], wl["H"], wl["W"], MODE_PRIMAL, fov=wl["F"], shared_map=wl["warehouse"],
                 threads=threads)
    orc.reset(obst, starts, goals)
    acts = [workloads.hash_actions_np(ACTION_SEED, range(n_envs), t, wl["N"]) for t in range(16)]
    want = ("agent_reward", "dones", "avail", "terminated", "reward")
    obs = np.empty((n_envs, wl["N"], 4, wl["F"], wl["F"]), np.uint8)
    vec = np.empty((n_envs, wl["N"], 3), np.float64)

    def one(t):
        orc.primal_sweep(acts[t % 16], want=want)
        orc._lib.oracle_primal_observe(orc._h, obs.ctypes.data_as(ctypes.c_void_p), vec.ctypes.data_as(ctypes.c_void_p))
    for t in range(2):
        one(t)
    steps, t0 = 0, time.perf_counter()
    while True:
        one(steps)
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds and steps >= 2:
            break
    return dict(value=n_envs * wl["N"] * steps / dt, unit=UNIT, cores=threads, kind="port",
                sample="%d envs x %d agents x %d steps (step+obs), C oracle port, OpenMP over envs on %d threads, "
                       "%.1f s" % (n_envs, wl["N"], steps, threads, dt))


def reference_pool_rate(wl, steps, warmup, target_step_seconds=1.0, max_envs_per_proc=64):
    """The UNMODIFIED Python reference env (mapf_primal.MAPFEnv) in one worker process per host core, each stepping
    `envs_per_proc` environments per step; envs_per_proc is calibrated so that one step takes about
    target_step_seconds."""
    from oracle.ref_pool import host_cores, run_reference_pool
    procs = host_cores()
    cal = run_reference_pool(wl, procs, 1, 2, 1, action_seed=ACTION_SEED)        # 2 timed env steps per worker
    t_env = max(cal["elapsed_s"] / 2, 1e-4)
    per = int(max(1, min(max_envs_per_proc, round(target_step_seconds / t_env))))
    res = run_reference_pool(wl, procs, per, steps, warmup, action_seed=ACTION_SEED)
    res["envs_per_proc"] = per
    return res


def reference_baseline(name, wl, res):
    return dict(value=res["value"], unit=UNIT, cores=res["procs"], kind="reference",
                sample="%d envs (%d per process) x %d agents x %d joint steps of %s through the UNMODIFIED "
                       "mapf_primal.MAPFEnv: _step swept over ids 1..N, then _observe for every agent "
                       "(mapf_primal.py:549-637, 343-386), %d worker processes (one per host core, like "
                       "parallel_runner.py:219-258), %.1f s"
                       % (res["envs"], res["envs_per_proc"], wl["N"], res["agent_steps"] // (res["envs"] * wl["N"]),
                          name, res["procs"], res["elapsed_s"]))


def workload_config(name, wl, gpus):
    return {"workload": "%s: %dx%d map, density %.2f, %d agents, FOV %d, %d envs per GPU%s" % (
        name, wl["H"], wl["W"], wl["density"], wl["N"], wl["F"], wl["E"],
        ", shared warehouse map" if wl["warehouse"] else ""),
        "mode": "primal (sequential claim) + 4-channel FOV observation + goal vector, fused step+obs kernel",
        "n_envs_total": wl["E"] * gpus, "n_agents": wl["N"],
        "actions": "counter hash of (seed, GLOBAL env index, step, agent), uniform over the 5 actions",
        "l2": ("per-step working set (obs output %.0f MB) exceeds the 126 MB L2" if
               wl["E"] * wl["N"] * 4 * wl["F"] ** 2 > 126e6 else
               "per-step obs output is %.0f MB (< 126 MB L2): the timed passes cycle 16 action tensors and every step "
               "overwrites the whole output, dirty lines are written back between steps") % (
            wl["E"] * wl["N"] * 4 * wl["F"] ** 2 / 1e6),
        "parallelism": "envs sharded by index over %d GPU(s), no data-path collective" % gpus}


def run_reference_arm(args, wl, rank):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    res = reference_pool_rate(wl, args.steps, min(args.warmup, 5))
    base = reference_baseline(args.workload, wl, res)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["elapsed_s"] / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.workload, wl, args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_checksum": res["checksum"],
        "note": "each step = every worker process advancing its %d environments by one joint step (a bounded sample "
                "of the workload: %d of the configuration's %d environments)" % (res["envs_per_proc"], res["envs"],
                                                                                wl["E"]),
    }
    if not args.no_cpu:
        try:
            line["cpu_port"] = cpu_port_rate(wl, wl["E"], 2.0)       # the full batch, >= 2 s timed
        except Exception as exc:
            line["cpu_port"] = {"error": repr(exc)}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-sweep", action="store_true", help="skip the per-config sweep")
    ap.add_argument("--no-rollout", action="store_true", help="skip the BatchedRunner rollout legs")
    ap.add_argument("--f32", action="store_true", help="emit float32 observations instead of uint8")
    ap.add_argument("--lean", action="store_true",
                    help="fused-step timing only, a fixed launch order: for the ncu recipes")
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
    if args.lean:
        args.no_cpu = args.no_sweep = args.no_rollout = True
    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return

    # ---- CPU baselines first (rank 0 at N = 1 only): nothing of this process touches CUDA or a process group yet, so
    #      the worker processes have the host cores to themselves
    cpu_baseline = cpu_port = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            res = reference_pool_rate(wl, steps=10, warmup=1)        # about 10 s of the Python reference
            cpu_baseline = reference_baseline(args.workload, wl, res)
        except Exception as exc:
            sys.stderr.write("Python reference baseline failed: %r\n" % (exc,))
        try:
            cpu_port = cpu_port_rate(wl, min(wl["E"], 16384), 4.0)
            if cpu_baseline is None:
                cpu_baseline = cpu_port
        except Exception as exc:
            sys.stderr.write("C port baseline failed: %r\n" % (exc,))

    import torch
    import torch.distributed as dist
    import mapf_marl_b200
    from mapf_marl_b200 import workloads
    from mapf_marl_b200.engine import MapfEngine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (use --impl reference for the CPU arm)")
    mapf_marl_b200.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, N, F = wl["E"], wl["N"], wl["F"]
    env_lo = rank * E                              # this rank's slice of the global batch
    odt = torch.float32 if args.f32 else torch.uint8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def timed_passes(fn_pass, n_pass, sync_ranks=True):
        """ms of each of n_pass consecutive passes (CUDA events on the launching stream, max over ranks)."""
        if sync_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_pass + 1)]
        evs[0].record()
        for i in range(n_pass):
            fn_pass()
            evs[i + 1].record()
        if sync_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_pass)]
        return max_over_ranks(ms) if sync_ranks else ms

    def timed(fn, steps, warmup, sync_ranks=True):
        for t in range(warmup):
            fn(t)
        return timed_passes(lambda: [fn(t) for t in range(steps)], 1, sync_ranks)[0]

    def measure_fused(eng, pool, steps, min_total_ms=200.0, min_passes=5, max_passes=400, dtype=odt, sync_ranks=True):
        """Median ms per pass of exactly `steps` fused steps replayed from CUDA graphs of `chunk` steps; passes are
        repeated until min_total_ms have been timed."""
        chunk = max(c for c in range(1, 21) if steps % c == 0)
        n_pool = pool.shape[0]
        graph = None
        try:
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                for t in range(chunk):
                    eng.step_observe(pool[t % n_pool], want=WANT, dtype=dtype)
        except Exception as exc:   # capture is a convenience of the host library, never a reason to lose the bench line
            sys.stderr.write("CUDA graph capture failed, timing eager launches: %r\n" % (exc,))
            graph = None
        if graph is not None:
            one_pass = lambda: [graph.replay() for _ in range(steps // chunk)]   # noqa: E731
        else:
            one_pass = lambda: [eng.step_observe(pool[t % n_pool], want=WANT, dtype=dtype) for t in range(steps)]   # noqa: E731
        first = timed_passes(one_pass, 2, sync_ranks)
        n_pass = int(min(max_passes, max(min_passes, np.ceil(min_total_ms / max(min(first), 1e-3)))))
        n_pass = int(max_over_ranks([n_pass])[0]) if sync_ranks else n_pass
        ms = timed_passes(one_pass, n_pass, sync_ranks)
        return ms, chunk, graph is not None

    def gen_actions(engine, lo, t, avail=None, seed=ACTION_SEED):
        """uint8 [E, N] actions of step t: the engine's counter-hash kernel (mapf_random_actions; tested equal to
        workloads.hash_actions_np, which the CPU arms use) keyed by the GLOBAL env index."""
        return engine.random_actions(seed, t, avail=avail, env_offset=lo, dtype=torch.uint8)

    def action_pool(engine, lo, n=16):
        return torch.stack([gen_actions(engine, lo, t).clone() for t in range(n)])

    # ---- the engine of the headline configuration
    obst, starts, goals = make_world(wl, E, env_lo)
    eng = MapfEngine(E, N, wl["H"], wl["W"], mode="primal", fov=F, shared_map=wl["warehouse"], goal_dist=True,
                     device=dev)
    eng.reset(obst, starts, goals)
    # goal-distance maps: computed at reset (and on goal reassignment), reported separately
    eng.refresh_goal_dist()
    bfs_ms = float(np.median(timed_passes(eng.refresh_goal_dist, 5)))
    args.warmup = max(args.warmup, 3)          # timing rule: at least three untimed warm-up steps

    # ---- canonical segment: reset, then exactly W + K steps with the counter-hash actions -> the state checksum
    #      (a function of the GLOBAL env indices this rank owns: identical for rank 0 at any GPU count)
    eng.reset(obst, starts, goals)
    for t in range(args.warmup + args.steps):
        last = eng.step_observe(gen_actions(eng, env_lo, t), want=WANT,
                                dtype="bits" if eng.bits_supported() else torch.uint8)
    checksum = workloads.state_checksum_torch(eng.positions(), last["obs"], last["avail"], last["terminated"])
    stats_canonical = eng.stats()

    # ---- device-resident throughput (the fused step+observation kernel), clocks sampled meanwhile
    pool = action_pool(eng, env_lo)
    for t in range(args.warmup):
        eng.step_observe(pool[t % 16], want=WANT, dtype=odt)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    passes, chunk, graphed = measure_fused(eng, pool, args.steps, min_total_ms=0.0 if args.lean else 200.0,
                                           min_passes=1 if args.lean else 5)
    ms_pass = float(np.median(passes))
    ms_step = ms_pass / args.steps
    value = world * E * N * args.steps / (ms_pass * 1e-3)
    launches = args.steps            # one tile-kernel launch per step (inside the graphs when replayed)
    n_eager = max(min(args.steps * 5, 2000), 1)
    ms_eager = timed(lambda t: eng.step_observe(pool[t % 16], want=WANT, dtype=odt), n_eager, 3) / n_eager
    # the clock sampler (2 ms period) keeps running through the eager leg of the same kernel; more of the same load,
    # untimed, until it has seen about 0.3 s of it
    if sampler and not args.lean:
        t_load = time.perf_counter()
        while len(sampler.sm) < 100 and time.perf_counter() - t_load < 0.5:
            for t in range(50):
                eng.step_observe(pool[t % 16], want=WANT, dtype=odt)
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    line_extra = {}
    if not args.lean:
        # ---- the same K steps as ONE mapf_rollout launch per chunk (tile resident in shared memory between steps)
        T_roll = 16
        racts = pool[:T_roll].contiguous()
        roll = None
        try:
            for _ in range(2):
                eng.rollout(racts, want=WANT, dtype=odt)
            ms_roll = float(np.median(timed_passes(lambda: eng.rollout(racts, want=WANT, dtype=odt),
                                                   max(5, int(200.0 / max(ms_step * T_roll, 1e-3)) // 4)))) / T_roll
            roll = {"ms_per_step": ms_roll, "steps_per_launch": T_roll, "one_launch": eng.rollout_in_one_launch(odt),
                    "kernel": eng.rollout_plan(T_roll, odt),
                    "agent_steps_per_s": world * E * N / (ms_roll * 1e-3),
                    "frac_of_hbm_peak": (wl["bytes_per_agent_step"] + (3 * 4 * F * F if args.f32 else 0)) * E * N /
                    (ms_roll * 1e-3) / 1e9 / measured_peak()[0]}
        except Exception as exc:
            roll = {"error": repr(exc)}
        line_extra["rollout_kernel"] = roll

        # ---- the fused launch with the observation left bit-packed (MAPF_BITS, for consumers that take bits)
        ms_bits = None
        if eng.bits_supported():
            ms_bits = timed(lambda t: eng.step_observe(pool[t % 16], want=WANT, dtype="bits"), n_eager, 3) / n_eager
        # ---- breakdown: the step-only and observe-only launches of the same tile kernel
        nb = max(args.steps, 20)
        ms_obs = timed(lambda t: eng.observe(dtype=odt), nb, 3) / nb
        ms_stp = timed(lambda t: eng.step(pool[t % 16], want=WANT), nb, 3) / nb

        # ---- actions sampled from the action mask (agents keep moving: the claim / arrival paths of the sweep get
        #      exercised).  The trajectory is recorded once (the mask of step t decides the actions of step t + 1), then
        #      replayed from reset: the state evolution is deterministic, so the recorded actions stay mask-consistent.
        T_rec = 64
        eng.reset(obst, starts, goals)
        avail = eng.avail()
        rec = torch.empty((T_rec, E, N), dtype=torch.uint8, device=dev)
        s0 = eng.stats()
        for t in range(T_rec):
            rec[t] = gen_actions(eng, env_lo, t, avail=avail, seed=ACTION_SEED + 1)
            avail = eng.step_observe(rec[t], want=WANT, dtype=odt)["avail"]
        s1 = eng.stats()
        g_m = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g_m):
            for t in range(T_rec):
                eng.step_observe(rec[t], want=WANT, dtype=odt)
        ms_masked = []
        for _ in range(5):
            eng.reset(obst, starts, goals)
            ms_masked.append(timed_passes(g_m.replay, 1)[0] / T_rec)
        del g_m
        masked = {"ms_per_step": float(np.median(ms_masked)),
                  "agent_steps_per_s": world * E * N / (float(np.median(ms_masked)) * 1e-3),
                  "steps_per_pass": T_rec, "passes": 5,
                  "rank0_stats_of_one_pass": {k: s1[k] - s0[k] for k in s1},
                  "note": "actions drawn uniformly from each agent's available actions "
                          "(_listNextValidActions): no wall / robot bumps, agents travel and reach goals"}
        line_extra["actions_avail_masked"] = masked
        eng.reset(obst, starts, goals)

    # ---- c4 only: the lifelong variant of the step (BASELINE config 4): after every fused launch, agents standing on
    #      their goal pop the next one from their queue (mapf_pop_goals) and the distance maps of exactly those goals are
    #      recomputed (mapf_bfs with the dirty mask): nothing leaves the device.
    def lifelong_leg(eng_l, wl_l, obst_l, starts_l, goals_l, lo_l, pool_l, n_steps):
        from mapf_marl_b200.lifelong import LifelongGoals
        E_l, N_l = eng_l.E, eng_l.N
        queue = make_goal_queue(wl_l, obst_l, goals_l, E_l, lo_l, depth=8)
        eng_l.reset(obst_l, starts_l, goals_l)
        eng_l.refresh_goal_dist()
        life = LifelongGoals(eng_l, queue, overlap=True, fused=True)

        def life_step(t):
            eng_l.step_observe(pool_l[t % pool_l.shape[0]], want=WANT, dtype=odt)
            life.reassign()
        # replayed from a CUDA graph of 10 steps (the side stream of the overlapped BFS is forked and joined inside
        # the capture): the eager loop is four launches plus event traffic per step from Python, i.e. it times the host
        launch_l, n_done = "eager", 0
        for t in range(3):
            life_step(t)
        life.sync()
        n_done += 3
        try:
            g_l = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g_l):
                for t in range(10):
                    life_step(t)
                life.sync()
            reps = max(n_steps // 10, 2)
            ms_life = float(np.median(timed_passes(lambda: [g_l.replay() for _ in range(reps)], 3,
                                                   sync_ranks=False))) / (10 * reps)
            n_done += 30 * reps
            launch_l = "CUDA graph replay, 10 steps per graph"
        except Exception as exc:
            sys.stderr.write("lifelong leg: graph capture failed, eager launches: %r\n" % (exc,))
            torch.cuda.synchronize()
            ms_life = timed(life_step, n_steps, 0, sync_ranks=False) / n_steps
            life.sync()
            n_done += n_steps
        popped = int(life.head.sum().item())
        return {"ms_per_step": ms_life, "agent_steps_per_s": E_l * N_l / (ms_life * 1e-3), "launch": launch_l,
                "goal_queue_depth": 8, "reassignments_per_step": popped / n_done, "launches_per_step": 3,
                "note": "queues bound to the handle (mapf_lifelong_bind): the fused step+obs launch pops the goal queue "
                        "of every agent that arrived and lists it, mapf_bfs_popped = BFS of the listed agents (+ its "
                        "overflow pass) on a side stream under the next step's launch; uniform random actions, so "
                        "arrivals are rare"}

    lifelong = None
    if wl["warehouse"] and not args.lean:
        lifelong = lifelong_leg(eng, wl, obst, starts, goals, env_lo, pool, max(min(args.steps * 5, 1000), 20))
        eng.reset(obst, starts, goals)

    # ---- end to end: pinned host actions in, every output back on the host, through the C-ABI host entry point
    e2e = e2e_variants = None
    if not args.lean:
        io, bufs, h2d, d2h = eng.make_host_io(obs_dtype=odt)
        host_pool_np = pool[:4].cpu().numpy()
        act_np = bufs["actions"].numpy()          # view of the pinned action buffer

        def e2e_step(t):
            # a plain memcpy into the pinned buffer (torch's CPU copy_ would fan out to its OpenMP pool, whose threads
            # then spin on every core for milliseconds and compete with the library's unpack threads)
            np.copyto(act_np, host_pool_np[t % 4])
            eng.step_observe_host(io)
        for t in range(8):                        # the first calls start the unpack pool and ramp the host clocks
            e2e_step(t)
        e2e_ms = timed_passes(lambda: [e2e_step(t) for t in range(args.e2e_steps)], 3)
        ms_e2e = float(np.median(e2e_ms)) / args.e2e_steps
        packed = eng.host_transport() == 1
        e2e = {"value": world * E * N / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": args.e2e_steps, "passes": 3,
               "transport": "bit-packed observation over PCIe, expanded to the requested dtype by the library's host "
                            "threads" if packed else "dense copies",
               "host_bytes_written_per_step": int(bufs["obs"].numel() * bufs["obs"].element_size())}
        e2e_variants = {}
        # the same call with dense copies of the observation tensor
        if packed:
            eng.host_transport(False)
            _, _, _, d2h_dense = eng.make_host_io(obs_dtype=odt)
            ms = float(np.median(timed_passes(lambda: [e2e_step(t) for t in range(args.e2e_steps)], 2))) / args.e2e_steps
            eng.host_transport(True)
            e2e_variants["dense_transport"] = {"value": world * E * N / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                               "d2h_bytes_per_step": d2h_dense}
        # the bit stream itself as the host output (MAPF_BITS; a CPU consumer expands what it reads: mapf_host_unpack)
        if eng.bits_supported():
            io_b, bufs_b, _, d2h_bits = eng.make_host_io(obs_dtype="bits")
            act_b = bufs_b["actions"].numpy()

            def e2e_bits_step(t):
                np.copyto(act_b, host_pool_np[t % 4])
                eng.step_observe_host(io_b)
            for t in range(3):
                e2e_bits_step(t)
            ms = float(np.median(timed_passes(lambda: [e2e_bits_step(t) for t in range(args.e2e_steps)], 3))) / args.e2e_steps
            t0 = time.perf_counter()
            sl = eng.unpack_host_obs(bufs_b["obs"], 0, min(E, 256))
            t_unpack = time.perf_counter() - t0
            e2e_variants["bits_to_host"] = {
                "value": world * E * N / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h_bits,
                "lazy_view": "mapf_host_unpack expands env slices on demand: %d envs (%.1f MB of uint8) in %.2f ms on "
                             "the calling thread" % (sl.shape[0], sl.numel() / 1e6, t_unpack * 1e3)}
        # the policy lives on the GPU (pymarl's controller does): actions come from the host, reward / terminated go
        # back, the observation stays in device memory for the agent network
        acts_dev = torch.empty((E, N), dtype=torch.uint8, device=dev)
        pin_act = torch.empty((E, N), dtype=torch.uint8).pin_memory()
        pin_rew = torch.empty((E,), dtype=torch.float64).pin_memory()
        pin_term = torch.empty((E,), dtype=torch.uint8).pin_memory()
        pin_np = pin_act.numpy()

        def e2e_dev_obs(t):
            np.copyto(pin_np, host_pool_np[t % 4])
            acts_dev.copy_(pin_act, non_blocking=True)
            out = eng.step_observe(acts_dev, want=WANT, dtype=odt)       # one fused launch, observation stays in HBM
            pin_rew.copy_(out["reward"], non_blocking=True)
            pin_term.copy_(out["terminated"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        n2 = max(args.e2e_steps * 5, 50)
        for t in range(3):
            e2e_dev_obs(t)
        ms = float(np.median(timed_passes(lambda: [e2e_dev_obs(t) for t in range(n2)], 3))) / n2
        e2e_variants["obs_on_device"] = {"value": world * E * N / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                         "h2d_bytes_per_step": E * N, "d2h_bytes_per_step": E * 9}
        # ---- transfer-limit model of the headline e2e call, from two measurements on THIS box: the PCIe device->host
        #      rate of the packed stream and the rate at which the unpack pool alone writes the expanded bytes
        try:
            import ctypes
            nbits_bytes = eng.packed_obs_bytes()
            src = torch.empty(nbits_bytes, dtype=torch.uint8, device=dev)
            dst = torch.empty(nbits_bytes, dtype=torch.uint8).pin_memory()
            for _ in range(2):
                dst.copy_(src, non_blocking=True)
            ms_pcie = float(np.median(timed_passes(lambda: dst.copy_(src, non_blocking=True), 5, sync_ranks=False)))
            lib = eng.lib
            lib.mapf_unpack_pool_create.restype = ctypes.c_void_p
            lib.mapf_unpack_pool_create.argtypes = [ctypes.c_int]
            lib.mapf_unpack_pool_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                 ctypes.c_int]
            lib.mapf_unpack_pool_destroy.argtypes = [ctypes.c_void_p]
            cells = E * N * 4 * F * F
            elem = 4 if args.f32 else 1
            hp = lib.mapf_unpack_pool_create(0)
            obs_ptr = bufs["obs"].data_ptr()
            for _ in range(2):
                lib.mapf_unpack_pool_run(hp, dst.data_ptr(), obs_ptr, cells, elem)
            tt = []
            for _ in range(5):
                t0 = time.perf_counter()
                lib.mapf_unpack_pool_run(hp, dst.data_ptr(), obs_ptr, cells, elem)
                tt.append(time.perf_counter() - t0)
            lib.mapf_unpack_pool_destroy(hp)
            ms_host = float(np.median(tt)) * 1e3
            e2e["limit"] = {
                "pcie_d2h_GBps": nbits_bytes / (ms_pcie * 1e-3) / 1e9, "host_expand_write_GBps": cells * elem / (ms_host * 1e-3) / 1e9,
                "ms_pcie_packed_stream": ms_pcie, "ms_host_expansion_alone": ms_host, "ms_kernel": ms_step,
                "predicted_ms_if_overlapped": max(ms_pcie, ms_host) + ms_step,
                "predicted_ms_if_serial": ms_pcie + ms_host + ms_step, "measured_ms": ms_e2e,
                "note": "the call leaves %d MB of expanded bytes in host memory: it cannot beat the slower of the PCIe "
                        "transfer of the packed stream and the host-memory write of the expansion; boxes where DMA "
                        "writes and 16 streaming cores share one memory path land near the serial figure"
                        % (cells * elem // 1000000)}
            del src, dst
        except Exception as exc:
            e2e["limit"] = {"error": repr(exc)}

    # ---- statistics: the only collective of the path (one all-reduce of 8 int64 over NCCL)
    stats = eng.stats()
    svec = torch.tensor([stats[k] for k in sorted(stats)], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(svec, op=dist.ReduceOp.SUM)
    flags = eng.error_flags()
    peak, peak_src = measured_peak()

    # ---- sweep: every BASELINE configuration and an E-sweep of the c3 shape ("agent-steps/s vs N envs"), per GPU.
    #      Bounded: each entry is timed for about 0.1 s.  Every rank runs it on its own GPU (no cross-rank step);
    #      rank 0 reports its numbers.
    sweep = None
    if not args.no_sweep:
        eng.close()
        del eng
        torch.cuda.empty_cache()
        sweep = []
        c3w = WORKLOADS["c3"]
        base_world = {}

        def world_for(name, wlx, n_envs, lo):
            # worlds cycle with period DISTINCT_WORLDS over the global env index: generate one period, tile it
            per = workloads.DISTINCT_WORLDS
            if wlx["warehouse"] or n_envs <= per or lo % per:
                return make_world(wlx, n_envs, lo)
            if name not in base_world:
                base_world[name] = make_world(wlx, per, 0)
            o, s, g = base_world[name]
            reps = (n_envs + per - 1) // per
            return (np.tile(o, (reps, 1, 1))[:n_envs], np.tile(s, (reps, 1, 1))[:n_envs], np.tile(g, (reps, 1, 1))[:n_envs])

        entries = [("c2", WORKLOADS["c2"], WORKLOADS["c2"]["E"], "BASELINE config 2"),
                   ("c4", WORKLOADS["c4"], WORKLOADS["c4"]["E"], "BASELINE config 4 (+ lifelong goal reassignment)"),
                   ("c5_share", c3w, 131072, "BASELINE config 5: one GPU's share of the 1M-env sweep on 8 GPUs")]
        entries += [("c3_E%d" % e, c3w, e, "c3 shape, E-sweep") for e in (1024, 4096, 16384, 65536, 262144)]
        for name, wlx, Ex, what in entries:
            try:
                lo = rank * Ex
                o, s, g = world_for(name, wlx, Ex, lo)
                ex = MapfEngine(Ex, wlx["N"], wlx["H"], wlx["W"], mode="primal", fov=wlx["F"],
                                shared_map=wlx["warehouse"], goal_dist=wlx["warehouse"], device=dev)
                ex.reset(o, s, g)
                px = action_pool(ex, lo, 16)
                for t in range(3):
                    ex.step_observe(px[t], want=WANT, dtype=odt)
                ms, ck, gr = measure_fused(ex, px, 20, min_total_ms=100.0, min_passes=5, max_passes=200,
                                           sync_ranks=False)
                us = float(np.median(ms)) / 20 * 1e3
                bytes_step = wlx["bytes_per_agent_step"] * Ex * wlx["N"]
                ent = {"name": name, "what": what, "envs_per_gpu": Ex, "n_agents": wlx["N"],
                       "map": "%dx%d" % (wlx["H"], wlx["W"]), "us_per_step": us,
                       "agent_steps_per_s_per_gpu": Ex * wlx["N"] / (us * 1e-6),
                       "frac_of_hbm_peak": bytes_step / (us * 1e-6) / 1e9 / peak, "passes": len(ms)}
                try:
                    ra = px.contiguous()
                    for _ in range(2):
                        ex.rollout(ra, want=WANT, dtype=odt)
                    n_r = max(5, int(100.0 / max(us * 16e-3, 1e-3)))
                    usr = float(np.median(timed_passes(lambda: ex.rollout(ra, want=WANT, dtype=odt), min(n_r, 200),
                                                       sync_ranks=False))) / 16 * 1e3
                    ent["rollout_kernel_us_per_step"] = usr
                    ent["rollout_kernel"] = ex.rollout_plan(16, odt)
                    ent["rollout_kernel_frac_of_hbm_peak"] = bytes_step / (usr * 1e-6) / 1e9 / peak
                except Exception as exc:
                    ent["rollout_kernel_error"] = repr(exc)
                if wlx["warehouse"]:
                    ent["lifelong"] = lifelong_leg(ex, wlx, o, s, g, lo, px, 200)
                    bfs = float(np.median(timed_passes(ex.refresh_goal_dist, 3, sync_ranks=False)))
                    ent["goal_bfs_all_maps_ms"] = bfs
                    ent["goal_maps_per_s"] = Ex * wlx["N"] / (bfs * 1e-3)
                sweep.append(ent)
                ex.close()
                del ex, px, o, s, g
                torch.cuda.empty_cache()
            except Exception as exc:
                sweep.append({"name": name, "error": repr(exc)})
        eng = None

    # ---- the other two reference semantics on a c3-shaped batch (same worlds): GRID (mapf_gridworld.py: detect-and-
    #      penalise step + full-map observation) and PARTIAL (marl_partial.py, the env the reference registers: window
    #      maps + K nearest agents, float64 like the reference and float32 like pymarl's episode batch stores it)
    modes = None
    if not args.no_sweep:
        if eng is not None:
            eng.close()
            del eng
            eng = None
            torch.cuda.empty_cache()
        modes = {}
        c3w = WORKLOADS["c3"]
        Em = c3w["E"]
        om, sm, gm = make_world(c3w, Em, rank * Em)
        pm = None
        for mname, kw, dts in (("grid", dict(mode="grid", episode_limit=10 ** 6), (None,)),
                               ("partial", dict(mode="partial", episode_limit=256, obs_window=11, obs_knn_agents=5),
                                (torch.float64, torch.float32))):
            try:
                em = MapfEngine(Em, c3w["N"], c3w["H"], c3w["W"], device=dev, **kw)
                em.reset(om, sm, gm)
                if pm is None:
                    pm = torch.stack([em.random_actions(ACTION_SEED, t, env_offset=rank * Em, dtype=torch.uint8).clone()
                                      for t in range(16)])
                for dt in dts:
                    okw = {} if dt is None else {"dtype": dt}
                    for t in range(3):
                        em.step_observe(pm[t], want=WANT, **okw)
                    n_m = 200 if mname == "grid" else 60
                    # replayed from CUDA graphs of 20 steps like the headline (a GRID step is shorter than the host
                    # side of an eager launch through ctypes: eager timing measured the host, not the kernel)
                    launch_m = "eager"
                    step_pass = lambda: [em.step_observe(pm[t % 16], want=WANT, **okw) for t in range(n_m)]   # noqa: E731
                    obs_pass = lambda: [em.observe(**okw) for _ in range(n_m)]   # noqa: E731
                    try:
                        g_s, g_o = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                        torch.cuda.synchronize()
                        with torch.cuda.graph(g_s):
                            for t in range(20):
                                em.step_observe(pm[t % 16], want=WANT, **okw)
                        with torch.cuda.graph(g_o):
                            for t in range(20):
                                em.observe(**okw)
                        step_pass = lambda: [g_s.replay() for _ in range(n_m // 20)]   # noqa: E731
                        obs_pass = lambda: [g_o.replay() for _ in range(n_m // 20)]   # noqa: E731
                        launch_m = "CUDA graph replay, 20 steps per graph"
                    except Exception as exc:
                        sys.stderr.write("modes leg: graph capture failed, eager launches: %r\n" % (exc,))
                    ms_f = float(np.median(timed_passes(step_pass, 5, sync_ranks=False))) / n_m
                    ms_o = float(np.median(timed_passes(obs_pass, 5, sync_ranks=False))) / n_m
                    obs_t = em.observe(**okw)[0]
                    obs_bytes = obs_t.numel() * obs_t.element_size()
                    # algorithmic bytes per step: the observation + per agent actions 1, pos r/w 4, goal 2, done r/w 2,
                    # avail 5 (+ PARTIAL: two distance gathers 4, bookkeeping r/w 18) + per env reward 8, flag 1, step 8
                    per_agent = 14 + (22 if mname == "partial" else 0)
                    alg = obs_bytes + Em * c3w["N"] * per_agent + Em * 17
                    key = mname if dt is None else "%s_%s" % (mname, str(dt).split(".")[-1])
                    modes[key] = {"envs_per_gpu": Em, "n_agents": c3w["N"], "fused_step_obs_us": ms_f * 1e3,
                                  "observe_only_us": ms_o * 1e3, "launches_per_step": 1 if mname == "grid" else 2,
                                  "launch": launch_m,
                                  "agent_steps_per_s_per_gpu": Em * c3w["N"] / (ms_f * 1e-3),
                                  "obs_bytes_per_step": obs_bytes, "algorithmic_bytes_per_step": alg,
                                  "frac_of_hbm_peak": alg / (ms_f * 1e-3) / 1e9 / peak,
                                  "observe_only_frac_of_hbm_peak": obs_bytes / (ms_o * 1e-3) / 1e9 / peak}
                em.close()
                del em
                torch.cuda.empty_cache()
            except Exception as exc:
                modes[mname] = {"error": repr(exc)}
        del om, sm, gm

    # ---- rollout: the reference-API loop (pymarl ParallelRunner semantics) through BatchedRunner over PrimalVecEnv,
    #      the kernel writing straight into the time-major episode batch; controllers: random over the action mask,
    #      and pymarl's recurrent agent (Linear-GRUCell-Linear, torch) -- the network is a caller, not the path
    rollout = None
    if not args.no_rollout:
        from mapf_marl_b200.batched_runner import BatchedRunner, RandomMAC, RNNAgentMAC
        from mapf_marl_b200.vec_env import PrimalVecEnv
        if eng is not None:
            eng.close()
            del eng
            eng = None
            torch.cuda.empty_cache()
        rollout = {}
        c3w = WORKLOADS["c3"]
        for name, Ex, T in (("c3", c3w["E"], 16), ("c5_share", 131072, 8)):
            try:
                lo = rank * Ex
                o, s, g = make_world(c3w, min(Ex, workloads.DISTINCT_WORLDS), 0) if Ex > workloads.DISTINCT_WORLDS and lo % workloads.DISTINCT_WORLDS == 0 \
                    else make_world(c3w, Ex, lo)
                if o.shape[0] < Ex:
                    reps = (Ex + o.shape[0] - 1) // o.shape[0]
                    o, s, g = (np.tile(o, (reps, 1, 1))[:Ex], np.tile(s, (reps, 1, 1))[:Ex], np.tile(g, (reps, 1, 1))[:Ex])
                env = PrimalVecEnv(o, s, g, fov=c3w["F"], episode_limit=T, device=dev)
                ent = {"envs_per_gpu": Ex, "steps_per_episode": T}
                for mac_name, mac in (("random_mac", RandomMAC(env.engine, seed=7, env_offset=lo)),
                                      ("rnn_mac", RNNAgentMAC(4 * c3w["F"] ** 2, 5, dev, extra_dim=3))):
                    runner = BatchedRunner(env, mac, check_every=8, cuda_graph=True)
                    l0 = env.engine.launch_count()
                    n_rst = 3 if env._maps_loaded else 4               # obstacle rows are built by the first reset only
                    runner.run()                                       # eager warm-up episode (allocates the batch)
                    # launches of this library in one eager episode: 3-4 at reset (obstacle rows, reset, observe, masks),
                    # then per environment step the fused step+obs kernel (+ the random policy's own kernel)
                    n_mac = T if mac_name == "random_mac" else 0
                    n_book = 2 * T if runner.fused_bookkeeping else 0   # mapf_runner_mask_actions + mapf_runner_account
                    env_launches_per_step = (env.engine.launch_count() - l0 - n_rst - n_mac - n_book) / float(T)
                    runner.run()                                       # captures the CUDA graphs, replays them
                    ms = []
                    for _ in range(3):
                        barrier()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        runner.run(reuse_batch=True)
                        e1.record()
                        barrier()
                        ms.append(e0.elapsed_time(e1))
                    ms = max_over_ranks([float(np.median(ms))])[0]

                    ent[mac_name] = {"value": world * Ex * c3w["N"] * T / (ms * 1e-3), "unit": UNIT,
                                     "ms_per_env_step": ms / T, "engine_launches_per_env_step": env_launches_per_step,
                                     "policy_launches_per_env_step": n_mac / float(T),
                                     "bookkeeping_launches_per_env_step": n_book / float(T),
                                     "loop": "one episode = reset + CUDA graphs of 8 environment steps each (controller + "
                                             "masking + fused engine launch + bookkeeping), one host check per graph; "
                                             "launch counts are those of the eager warm-up episode",
                                     "batch_bytes": runner.batch.nbytes()}
                    del runner
                rollout[name] = ent
                env.close()
                del env, o, s, g
                torch.cuda.empty_cache()
            except Exception as exc:
                rollout[name] = {"error": repr(exc)}
        if e2e_variants and "obs_on_device" in e2e_variants and "random_mac" in rollout.get("c3", {}):
            rollout["c3"]["random_mac"]["vs_e2e_obs_on_device"] = \
                rollout["c3"]["random_mac"]["value"] / e2e_variants["obs_on_device"]["value"]

    if rank == 0:
        bytes_per = wl["bytes_per_agent_step"] + (3 * 4 * F * F if args.f32 else 0)
        alg_bytes = bytes_per * E * N
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        traffic, traffic_source = None, None
        for cand in ("r2_fused_traffic.json", "r1_fused_traffic.json"):
            tp = os.path.join(ROOT, "profiles", cand)
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get(args.workload)
                    traffic_source = ("profiles/%s: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` "
                                      "capture of this kernel at this workload (a static record, not this run)" % cand)
                    break
                except Exception:
                    traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32 obs / u8 state" if args.f32 else "u8", "data": "synthetic",
            "config": workload_config(args.workload, wl, world),
            "launch": ("CUDA graph replay, %d steps per graph" % chunk) if graphed
            else "eager launches (one C-ABI call per step)",
            "timing": {"passes": len(passes), "steps_per_pass": args.steps, "value_from": "median pass",
                       "ms_per_pass_median": ms_pass, "ms_per_pass_min": float(min(passes)),
                       "ms_per_pass_max": float(max(passes)), "timed_total_ms": float(sum(passes))},
            "rank0_state_checksum": "%016x" % checksum,
            "rank0_state_checksum_of": "positions + bit-packed observation + action masks + terminated flags of rank "
                                       "0's envs (global indices 0..%d) after reset + %d + %d steps" % (
                                           E - 1, args.warmup, args.steps),
            "e2e": e2e, "e2e_variants": e2e_variants,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                         "peak_source": peak_src,
                         "kernel": "mapf_tile_kernel<%d> (fused step+obs)" % F,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes_per_agent_step": bytes_per},
            "clocks": clocks,
            "stats": dict(zip(sorted(stats), [int(v) for v in svec.tolist()])),
            "stats_rank0_canonical_segment": stats_canonical,
            "device_error_flags": flags,
            "sweep": sweep, "modes": modes, "rollout": rollout, "lifelong": lifelong,
        }
        if not args.lean:
            line["breakdown_ms"] = {
                "fused_step_obs": ms_step, "fused_step_obs_eager_launches": ms_eager,
                "fused_step_obs_bit_packed_output": ms_bits, "observe_only": ms_obs, "step_only": ms_stp,
                "observe_only_GBps": (4 * F * F * (4 if args.f32 else 1) + 24 + 8 + wl["H"] * wl["W"] / N) * E * N /
                (ms_obs * 1e-3) / 1e9,
                "goal_bfs_all_maps": bfs_ms, "goal_maps_per_s": E * N / (bfs_ms * 1e-3)}
        line.update(line_extra)
        if e2e is None:
            line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                           "note": "--lean run: not measured"}
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
            if cpu_port is not None and cpu_port is not cpu_baseline:
                line["cpu_port"] = cpu_port
        elif world > 1:
            line["cpu_baseline"] = None     # measured on rank 0 at N = 1 only (idle ranks would disturb it)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
