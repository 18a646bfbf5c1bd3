#!/bin/bash
# torchrun runs of the bench, one rank per GPU: profiles/scale_r2.sh <N> [extra bench args]
N=${1:-2}; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/r2_bench_c3_${N}gpu.json 2> gpurun_out/r2_bench_c3_${N}gpu.err
tail -c 400 gpurun_out/r2_bench_c3_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm_${N}gpu.json 2> gpurun_out/r2_bench_reference_arm_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_c3_${N}gpu.json"))
print("N=${N} value %.4g ms %.5f checksum %s e2e %.4g" % (d["value"], d["ms_per_step"], d["rank0_state_checksum"], d["e2e"]["value"]))
print({k: "%.3g" % v["value"] for k, v in d["e2e_variants"].items()})
print({k: {m: "%.3g" % v["value"] for m, v in e.items() if isinstance(v, dict)} for k, e in d["rollout"].items()})
PY
