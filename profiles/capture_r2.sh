#!/bin/bash
# Round-2 profiling recipe (run on the GPU box through gpurun, ONE GPU): profiles/capture_r2.sh
#   1. the plain run must exit 0 first; 2. launch list of the same command; 3. one `--set full` capture per kernel.
# Tile-kernel launch order of `bench.py --steps 5 --warmup 3 --lean`: 0-7 the canonical checksum segment (bit-packed
# observation output, each preceded by the mapf_random_actions launch that draws its actions), 8-10 warm-up, 11-25 the
# timed CUDA-graph replays (3 passes x 5 steps: nothing but this kernel), 26-53 eager launches.
# Reports are written to /tmp on the box and summarised there (a report with sources is ~11 MB, gpurun brings back at
# most 64 MiB); only the text summaries and the c3 fused report travel.
set -e
TAG=${1:-r2}
OUT=gpurun_out
REP=/tmp/ncu_${TAG}
mkdir -p $REP
BENCH="python bench.py --steps 5 --warmup 3 --lean"
$BENCH > $OUT/${TAG}_plain_c3.json 2> $OUT/${TAG}_plain_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c3.csv $BENCH > $OUT/${TAG}_ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:mapf_tile_kernel -s 30 -c 1 -o $REP/${TAG}_fused_c3 $BENCH > $OUT/${TAG}_ncu_fused_c3.log 2>&1
$NCU -k regex:mapf_bfs_warp -s 1 -c 1 -o $REP/${TAG}_bfs_c3 $BENCH > $OUT/${TAG}_ncu_bfs_c3.log 2>&1
for WL in c4 c2; do
  B2="python bench.py --steps 5 --warmup 3 --lean --workload $WL"
  $B2 > $OUT/${TAG}_plain_$WL.json 2> $OUT/${TAG}_plain_$WL.err
  $NCU -k regex:mapf_tile_kernel -s 30 -c 1 -o $REP/${TAG}_fused_$WL $B2 > $OUT/${TAG}_ncu_fused_$WL.log 2>&1
done
# mapf_rollout at c2: the pipelined kernel (launches 0-2 warm-up, 3.. timed), and the tile kernel's in-kernel loop
# (MAPF_B200_PIPE=0) for comparison
python profiles/rollout_probe.py c2 > $OUT/${TAG}_rollout_probe_c2.log 2>&1
$NCU -k regex:mapf_pipe_kernel -s 4 -c 1 -o $REP/${TAG}_rollout_c2 python profiles/rollout_probe.py c2 > $OUT/${TAG}_ncu_rollout_c2.log 2>&1
MAPF_B200_PIPE=0 $NCU -k regex:mapf_tile_kernel -s 4 -c 1 -o $REP/${TAG}_rollout_tile_c2 python profiles/rollout_probe.py c2 > $OUT/${TAG}_ncu_rollout_tile_c2.log 2>&1
python profiles/summarize_ncu.py $REP/${TAG}_fused_c3.ncu-rep $REP/${TAG}_bfs_c3.ncu-rep > $OUT/${TAG}_ncu_summary_c3.txt
python profiles/summarize_ncu.py $REP/${TAG}_fused_c4.ncu-rep > $OUT/${TAG}_ncu_summary_c4.txt
python profiles/summarize_ncu.py $REP/${TAG}_fused_c2.ncu-rep $REP/${TAG}_rollout_c2.ncu-rep $REP/${TAG}_rollout_tile_c2.ncu-rep > $OUT/${TAG}_ncu_summary_c2.txt
python - <<PY > $OUT/${TAG}_fused_traffic.json
import csv, io, json, subprocess
out = {}
for wl in ("c2", "c3", "c4"):
    txt = subprocess.run(["ncu", "-i", "$REP/${TAG}_fused_%s.ncu-rep" % wl, "--page", "raw", "--csv"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, r = rows[0], rows[1], rows[2]
    def val(m):
        i = hdr.index(m)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    out[wl] = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
print(json.dumps(out))
PY
python profiles/by_line.py $REP/${TAG}_fused_c3.ncu-rep 40 > $OUT/${TAG}_by_line_fused_c3.txt 2>/dev/null || true
python profiles/by_line.py $REP/${TAG}_fused_c4.ncu-rep 40 > $OUT/${TAG}_by_line_fused_c4.txt 2>/dev/null || true
cp $REP/${TAG}_fused_c3.ncu-rep $OUT/ || true
ls -la $OUT | grep ${TAG}_
