#!/bin/bash
# variant_sweep.sh <workloads...> -- <variant names...>: fused launch (graph) and rollout µs per step of every
# experimental build in tmp_libs/ (profiles/build_variant.sh), one process per variant
wls=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do wls+=("$1"); shift; done; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/variant_sweep.log
  MAPF_B200_LIB=$PWD/tmp_libs/lib_$v.so timeout 300 python profiles/rollout_probe.py "${wls[@]}" 2>&1 | tail -n ${#wls[@]} \
    | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln.strip()); continue
    print('%s graph %.2f us (%.3f)  rollout %.2f us (%.3f)' % (d['workload'], d['graph_us_per_step'], d['graph_frac_hbm'], d['rollout_us_per_step'], d['rollout_frac_hbm']))
" | tee -a gpurun_out/variant_sweep.log
done
