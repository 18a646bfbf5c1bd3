#!/bin/bash
# Profiling recipe for the c3 bench command (run on the GPU box through gpurun, ONE GPU):
#   1. the plain run must exit 0 first; 2. launch list; 3. one `--set full` capture per kernel of interest.
# Tile-kernel launch order inside `bench.py --steps 5 --warmup 3 --lean`: 0-25 fused step+obs (eager warm-up, graph replays,
# eager leg), 26-33 fused with bit-packed observation output, 34-41 observe only, 42-49 step only.
# gpurun brings back at most 64 MiB of gpurun_out/ (local contents included) and a report with sources is ~11 MB:
# the reports are written to /tmp on the box, summarised there, and only the fused one travels (KEEP_REPORTS=1: all).
set -e
TAG=${1:-r1}
OUT=gpurun_out
REP=/tmp/ncu_${TAG}
mkdir -p $REP
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 2 --lean"
$BENCH > $OUT/${TAG}_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c3.csv $BENCH > $OUT/${TAG}_ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:mapf_tile_kernel -s 5 -c 1 -o $REP/${TAG}_fused_c3 $BENCH > $OUT/${TAG}_ncu_fused.log 2>&1
$NCU -k regex:mapf_tile_kernel -s 36 -c 1 -o $REP/${TAG}_obs_c3 $BENCH > $OUT/${TAG}_ncu_obs.log 2>&1
$NCU -k regex:mapf_tile_kernel -s 28 -c 1 -o $REP/${TAG}_bits_c3 $BENCH > $OUT/${TAG}_ncu_bits.log 2>&1
$NCU -k regex:mapf_tile_kernel -s 44 -c 1 -o $REP/${TAG}_step_c3 $BENCH > $OUT/${TAG}_ncu_step.log 2>&1
$NCU -k regex:mapf_bfs_warp -s 1 -c 1 -o $REP/${TAG}_bfs_c3 $BENCH > $OUT/${TAG}_ncu_bfs.log 2>&1
python profiles/summarize_ncu.py $REP/${TAG}_fused_c3.ncu-rep $REP/${TAG}_obs_c3.ncu-rep $REP/${TAG}_bits_c3.ncu-rep \
    $REP/${TAG}_step_c3.ncu-rep $REP/${TAG}_bfs_c3.ncu-rep > $OUT/${TAG}_ncu_summary_c3.txt
if [ -n "$KEEP_REPORTS" ]; then cp $REP/*.ncu-rep $OUT/; else cp $REP/${TAG}_fused_c3.ncu-rep $OUT/; fi
ls -la $OUT | grep ${TAG}_
