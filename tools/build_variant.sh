#!/bin/bash
# A/B builds: tools/build_variant.sh <name> "<-D flags>" [file ...]  ->  tools/prev_lib/lib_<name>.so
# (recompiles the named csrc files (default icm.cu) with the flags, links them with the other objects of the current build;
#  select at run time with RAYUELA_B200_LIB)
set -e
cd "$(dirname "$0")/../rayuela.jl_b200"
name=$1; flags=$2; shift 2 || true
files=${@:-icm.cu}
mkdir -p build_$name ../tools/prev_lib
objs=""
for f in csrc/*.cu; do
  b=$(basename $f .cu)
  if [[ " $files " == *" $b.cu "* ]]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 -ccbin /usr/bin/g++ --threads 4 $flags -c -o build_$name/$b.o $f
    objs="$objs build_$name/$b.o"
  else
    objs="$objs build/$b.o"
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../tools/prev_lib/lib_$name.so $objs
rm -rf build_$name
echo ../tools/prev_lib/lib_$name.so
