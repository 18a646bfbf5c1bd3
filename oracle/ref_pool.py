"""The CPU reference baseline: P worker processes, each stepping its own instances of the UNMODIFIED reference env
(oracle/ref_worker.py), mirroring the reference's ParallelRunner (one env per worker process,
MARL-curve-main/src/runners/parallel_runner.py:23-31, 219-258).  Test / benchmark infrastructure only."""
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference_pool(wl, procs, envs_per_proc, steps, warmup, world_seed=1000, action_seed=1234, timeout=1800):
    """Returns dict(value agent-steps/s, elapsed_s (slowest worker), wall_s, procs, envs, agent_steps)."""
    from . import refload
    if not refload.available():
        raise RuntimeError("the reference files are not available (neither /root/reference nor oracle/_ref)")
    env = dict(os.environ)
    env["MAPF_REFERENCE_ROOT"] = refload.REF
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        env[k] = "1"                               # one core per worker, like one env process per core
    workers = []
    for w in range(procs):
        cfg = dict(wl={k: v for k, v in wl.items()}, env_lo=w * envs_per_proc, n_envs=envs_per_proc, steps=steps,
                   warmup=warmup, world_seed=world_seed, action_seed=action_seed)
        workers.append(subprocess.Popen([sys.executable, os.path.join(HERE, "ref_worker.py"), json.dumps(cfg)],
                                        stdin=subprocess.PIPE, stdout=subprocess.PIPE, env=env, text=True))
    try:
        for p in workers:
            line = p.stdout.readline()
            if line.strip() != "ready":
                raise RuntimeError("reference worker failed to start: %r" % line)
        t0 = time.perf_counter()
        for p in workers:
            p.stdin.write("go\n")
            p.stdin.flush()
        results = []
        for p in workers:
            line = p.stdout.readline()
            if not line:
                raise RuntimeError("reference worker died")
            results.append(json.loads(line))
        wall = time.perf_counter() - t0
    finally:
        for p in workers:
            try:
                p.stdin.close()
            except Exception:
                pass
        for p in workers:
            try:
                p.wait(timeout=10)
            except Exception:
                p.kill()
    total = sum(r["agent_steps"] for r in results)
    slowest = max(r["elapsed"] for r in results)
    return dict(value=total / slowest if slowest > 0 else 0.0, elapsed_s=slowest, wall_s=wall, procs=procs,
                envs=procs * envs_per_proc, agent_steps=total, checksum=sum(r["checksum"] for r in results))
