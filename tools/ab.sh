#!/bin/bash
# tools/ab.sh "<icm_bench args>" variant...   -- times tools/icm_bench.py under tools/prev_lib/lib_<variant>.so
args=$1; shift
for v in "$@"; do
  echo "== $v"; RAYUELA_B200_LIB=/root/repo/tools/prev_lib/lib_$v.so python tools/icm_bench.py $args | tail -2
done
