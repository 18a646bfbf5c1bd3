#!/bin/bash
# variant_modes.sh <mode> -- <variants...>: profiles/modes_time.py for every experimental build in tmp_libs/
mode=$1; shift; shift
for v in "$@"; do
  echo "== $v"
  MAPF_B200_LIB=$PWD/tmp_libs/lib_$v.so timeout 300 python profiles/modes_time.py $mode 2>&1 | tail -2
done
