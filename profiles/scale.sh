#!/bin/bash
# Multi-GPU runs (gpurun --gpus N): the c3 bench line (weak scaling: 16384 envs per GPU) and the c5 sweep (1 048 576 envs
# in total, strong scaling), one rank per GPU under torchrun, as the driver launches them.
N=${1:-8}
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --no-cpu > $OUT/r1_bench_c3_${N}gpu.json 2> $OUT/r1_bench_c3_${N}gpu.err
if [ -n "$C3_ONLY" ]; then cat $OUT/r1_bench_c3_${N}gpu.json | cut -c1-600; exit 0; fi
$TR bench.py --gpus $N --workload c5 --no-cpu --steps 200 --warmup 5 --e2e-steps 4 > $OUT/r1_bench_c5_${N}gpu.json 2> $OUT/r1_bench_c5_${N}gpu.err
$TR bench.py --gpus $N --impl reference --steps 3 --warmup 1 > $OUT/r1_bench_reference_arm_${N}gpu.json 2>> $OUT/r1_bench_c3_${N}gpu.err
tail -n 2 $OUT/r1_bench_c3_${N}gpu.err $OUT/r1_bench_c5_${N}gpu.err
cat $OUT/r1_bench_c3_${N}gpu.json $OUT/r1_bench_c5_${N}gpu.json $OUT/r1_bench_reference_arm_${N}gpu.json | cut -c1-900
