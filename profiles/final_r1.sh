#!/bin/bash
# Round-1 evidence run (one B200): plain bench lines for every workload, the launch list of the c3 command and the
# ncu capture of the BFS kernel.  (The `--set full` captures of the tile kernel are taken by capture.sh /
# capture_one.sh; gpurun brings back at most 64 MiB, so the reports are summarised on the box and only the
# summaries travel: SUMMARIZE=1 below.)
set -x
OUT=gpurun_out
python bench.py > $OUT/r1_bench_c3.json 2> $OUT/r1_bench_c3.err
python bench.py --workload c2 --no-cpu > $OUT/r1_bench_c2.json 2>> $OUT/r1_bench_c3.err
python bench.py --workload c4 --no-cpu --steps 2000 > $OUT/r1_bench_c4.json 2>> $OUT/r1_bench_c3.err
python bench.py --envs 1048576 --no-cpu --steps 200 --warmup 5 --e2e-steps 2 > $OUT/r1_bench_c5_1gpu.json 2>> $OUT/r1_bench_c3.err
python bench.py --f32 --no-cpu --steps 1000 > $OUT/r1_bench_c3_f32.json 2>> $OUT/r1_bench_c3.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r1_bench_reference_arm.json 2>> $OUT/r1_bench_c3.err
python profiles/bench_modes.py > $OUT/r1_bench_modes.json 2>> $OUT/r1_bench_c3.err
python profiles/host_unpack_probe.py > $OUT/r1_host_unpack_probe.json 2>> $OUT/r1_bench_c3.err
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 2 --lean"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r1_launches_c3.csv $BENCH > $OUT/r1_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:mapf_bfs_warp -s 1 -c 1 -o /tmp/r1_bfs_c3 $BENCH > $OUT/r1_ncu_bfs.log 2>&1
python profiles/summarize_ncu.py /tmp/r1_bfs_c3.ncu-rep > $OUT/r1_ncu_summary_bfs_c3.txt 2>&1
