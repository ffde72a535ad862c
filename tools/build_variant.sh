#!/bin/bash
# A/B builds of one kernel file: tools/build_variant.sh <name> <file.cu> "<-D flags>"  -> variants/lib_<name>.so
# (the other objects come from geoa3_b200/build/, so run `python -m geoa3_b200.build` first; select a variant at run
# time with GEOA3_SO_PATH=variants/lib_<name>.so — measurement only)
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; defs=$3
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  --expt-relaxed-constexpr --extended-lambda $defs -c geoa3_b200/csrc/$src -o variants/${name}_${src%.cu}.o
objs=""
for o in geoa3_b200/build/*.o; do
  if [ "$(basename $o)" == "${src%.cu}.o" ]; then objs="$objs variants/${name}_${src%.cu}.o"; else objs="$objs $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_${name}.so $objs
echo variants/lib_${name}.so
