#!/bin/bash
# Round-1 evidence run (one B200): plain bench lines for every workload, then the ncu captures of the c3 command.
set -x
OUT=gpurun_out
python bench.py > $OUT/r1_bench_c3.json 2> $OUT/r1_bench_c3.err
python bench.py --workload c2 --no-cpu > $OUT/r1_bench_c2.json 2>> $OUT/r1_bench_c3.err
python bench.py --workload c4 --no-cpu --steps 2000 > $OUT/r1_bench_c4.json 2>> $OUT/r1_bench_c3.err
python bench.py --envs 1048576 --no-cpu --steps 200 --warmup 5 --e2e-steps 2 > $OUT/r1_bench_c5_1gpu.json 2>> $OUT/r1_bench_c3.err
python bench.py --f32 --no-cpu --steps 1000 > $OUT/r1_bench_c3_f32.json 2>> $OUT/r1_bench_c3.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r1_bench_reference_arm.json 2>> $OUT/r1_bench_c3.err
bash profiles/capture.sh r1
