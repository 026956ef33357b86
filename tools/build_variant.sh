#!/bin/bash
# tools/build_variant.sh <name> [extra nvcc flags...]: builds variants/<name>/libcuhe_b200.so for A/B runs
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p variants/$name
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin g++"
nvcc $F "$@" -c cuhe_b200/csrc/capi.cu -o variants/$name/capi.o &
nvcc $F "$@" -c cuhe_b200/csrc/ntt_launch.cu -o variants/$name/ntt_launch.o &
wait
nvcc -shared -o variants/$name/libcuhe_b200.so variants/$name/capi.o variants/$name/ntt_launch.o -gencode arch=compute_100a,code=sm_100a -ccbin g++
echo built variants/$name
