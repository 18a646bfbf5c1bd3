#!/bin/bash
# one `--set full` capture of a tile-kernel launch of the c3 bench command: capture_one.sh <tag> <fused|obs|step>
set -e
TAG=$1; WHICH=${2:-fused}
# tile-kernel launch order of the bench command below: 0-25 fused step+obs (eager warm-up, graph replays, eager leg),
# 26-33 fused with bit-packed output, 34-41 observe only, 42-49 step only
case $WHICH in fused) SKIP=5;; bits) SKIP=28;; obs) SKIP=36;; step) SKIP=44;; esac
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 2 --lean --workload ${WL:-c3}"
$BENCH > gpurun_out/${TAG}_plain.log 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:mapf_tile_kernel -s $SKIP -c 1 -o gpurun_out/${TAG}_${WHICH}_${WL:-c3} $BENCH > gpurun_out/${TAG}_ncu_${WHICH}.log 2>&1
