#!/bin/bash
# build_variant.sh NAME FILE.cu -DFLAG...: kernel-experiment variant of the library -> build/variants/NAME/libmode_b200.so
# (FILE.cu recompiled with the extra flags, every other object taken from the normal build).  Use with MODE_B200_LIB=<that path>.
set -e
name=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
python -m mode_2022_b200.build > /dev/null
out=$root/build/variants/$name; mkdir -p $out
base=$(basename $src .cu)
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -O3 --expt-relaxed-constexpr "$@" -c $root/mode_2022_b200/csrc/$base.cu -o $out/$base.o
objs=$(ls $root/build/obj/*.o | grep -v "/$base.o")
/usr/local/cuda/bin/nvcc -shared -o $out/libmode_b200.so $objs $out/$base.o -gencode arch=compute_100a,code=sm_100a -lcudart -lcuda
echo $out/libmode_b200.so
