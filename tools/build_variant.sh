#!/bin/bash
# build an A/B variant of libdrt.so: tools/build_variant.sh <out.so> <extra nvcc flags for the kernel TUs...>
set -e
OUT=$1; shift
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++"
mkdir -p build/variant
nvcc $F "$@" -c distraytracer_b200/csrc/drt_api.cu -o build/variant/api.o &
nvcc $F -c distraytracer_b200/csrc/drt_mesh.cu -o build/variant/mesh.o &
nvcc $F -fmad=false "$@" -c distraytracer_b200/csrc/drt_kernels_f64.cu -o build/variant/f64.o &
nvcc $F "$@" -c distraytracer_b200/csrc/drt_kernels_f32.cu -o build/variant/f32.o &
wait
nvcc -shared -ccbin /usr/bin/g++ -o "$OUT" build/variant/api.o build/variant/mesh.o build/variant/f64.o build/variant/f32.o
