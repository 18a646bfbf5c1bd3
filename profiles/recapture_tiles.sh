set -e
OUT=gpurun_out; REP=/tmp/ncu_r1; mkdir -p $REP
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 2 --lean"
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:mapf_tile_kernel -s 5 -c 1 -o $REP/r1_fused_c3 $BENCH > $OUT/r1_ncu_fused.log 2>&1
$NCU -k regex:mapf_tile_kernel -s 36 -c 1 -o $REP/r1_obs_c3 $BENCH > $OUT/r1_ncu_obs.log 2>&1
$NCU -k regex:mapf_tile_kernel -s 28 -c 1 -o $REP/r1_bits_c3 $BENCH > $OUT/r1_ncu_bits.log 2>&1
$NCU -k regex:mapf_tile_kernel -s 44 -c 1 -o $REP/r1_step_c3 $BENCH > $OUT/r1_ncu_step.log 2>&1
python profiles/summarize_ncu.py $REP/r1_fused_c3.ncu-rep $REP/r1_obs_c3.ncu-rep $REP/r1_bits_c3.ncu-rep $REP/r1_step_c3.ncu-rep > $OUT/r1_ncu_summary_tiles_c3.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r1_launches_c3.csv $BENCH > $OUT/r1_ncu_launches.log 2>&1
