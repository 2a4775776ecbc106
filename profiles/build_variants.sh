#!/bin/bash
# Builds tuning variants of the library (one translation unit recompiled with -D overrides, linked with the standard
# objects) into variants/<name>.so for A/B runs on one GPU box:  RETINA_B200_LIB=variants/<name>.so python profiles/dev_bench.py ...
#   usage: profiles/build_variants.sh <source.cu[,source2.cu...]> name1 "-DX=1 -DY=2" name2 "..." ...
set -e
cd "$(dirname "$0")/.."
python -c "from neuralnetworklibrary_b200 import _lib; _lib.build_library()"
SRC=$1; shift
B=neuralnetworklibrary_b200/csrc/_build
mkdir -p variants
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC"
SRCS=${SRC//,/ }
while [ $# -gt 1 ]; do
  name=$1; defs=$2; shift 2
  ( objs=$(ls $B/rn_*.o); vobjs=""
    for src in $SRCS; do
      nvcc $FLAGS $defs -c -o $B/var_${name}_${src%.cu}.o neuralnetworklibrary_b200/csrc/$src &
      objs=$(echo "$objs" | grep -v "/${src%.cu}.o"); vobjs="$vobjs $B/var_${name}_${src%.cu}.o"
    done
    wait
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/$name.so $objs $vobjs ) &
done
wait
ls -la variants/
