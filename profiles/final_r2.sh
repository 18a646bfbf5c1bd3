#!/bin/bash
# Round-2 evidence run (one B200): the GPU test suite, smoke, and the plain bench lines of every workload.
OUT=gpurun_out
( time python -m pytest tests/ -x -q -m gpu ) > $OUT/r2_pytest_gpu.log 2>&1
tail -3 $OUT/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2_bench_c3.json 2> $OUT/r2_bench_c3.err ) 2>&1 | grep real
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r2_bench_reference_arm.json 2> $OUT/r2_bench_reference_arm.err ) 2>&1 | grep real
python bench.py --workload c2 --no-cpu --steps 20 --warmup 5 --no-sweep --no-rollout > $OUT/r2_bench_c2.json 2> $OUT/r2_bench_c2.err
python bench.py --workload c4 --no-cpu --steps 20 --warmup 5 --no-sweep --no-rollout > $OUT/r2_bench_c4.json 2> $OUT/r2_bench_c4.err
python bench.py --f32 --no-cpu --steps 20 --warmup 5 --no-sweep --no-rollout > $OUT/r2_bench_c3_f32.json 2> $OUT/r2_bench_c3_f32.err
python profiles/bfs_time.py > $OUT/r2_bfs_time.txt 2>&1
python profiles/runner_probe.py 16384 > $OUT/r2_runner_probe.txt 2>&1
python profiles/runner_probe.py 131072 >> $OUT/r2_runner_probe.txt 2>&1
python - <<PY
import json
for n in ("c3", "c2", "c4", "c3_f32", "reference_arm"):
    try:
        d = json.load(open("gpurun_out/r2_bench_%s.json" % n))
        print(n, "value %.4g ms %.5f frac %s e2e %s" % (d["value"], d.get("ms_per_step", 0), (d.get("roofline") or {}).get("frac"), (d.get("e2e") or {}).get("value")))
    except Exception as exc:
        print(n, "FAILED", exc)
d = json.load(open("gpurun_out/r2_bench_c3.json"))
json.dump(d["modes"], open("gpurun_out/r2_bench_modes.json", "w"), indent=1)
print({k: round(v["fused_step_obs_us"], 1) for k, v in d["modes"].items()})
print({k: {m: "%.3g" % v["value"] for m, v in e.items() if isinstance(v, dict)} for k, e in d["rollout"].items()})
print("cpu", d["cpu_baseline"]["value"], d["cpu_port"]["value"], "checksum", d["rank0_state_checksum"])
c4 = json.load(open("gpurun_out/r2_bench_c4.json"))
print("c4 lifelong", c4["lifelong"]["ms_per_step"])
PY
