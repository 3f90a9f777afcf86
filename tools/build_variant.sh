#!/bin/bash
# build an A/B variant of libdrt.so: tools/build_variant.sh <out.so> <extra nvcc flags for the kernel TUs...>
set -e
OUT=$1; shift
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++"
D=build/variant_$(basename "$OUT" .so)
mkdir -p "$D"
nvcc $F "$@" -c distraytracer_b200/csrc/drt_api.cu -o $D/api.o &
nvcc $F -c distraytracer_b200/csrc/drt_mesh.cu -o $D/mesh.o &
nvcc $F -fmad=false -c distraytracer_b200/csrc/drt_skeleton.cu -o $D/skel.o &
for g in 0 1 2; do
  nvcc $F -fmad=false "$@" -c distraytracer_b200/csrc/drt_kernels_f64_g$g.cu -o $D/f64_g$g.o &
  nvcc $F "$@" -c distraytracer_b200/csrc/drt_kernels_f32_g$g.cu -o $D/f32_g$g.o &
done
wait
nvcc -shared -ccbin /usr/bin/g++ -o "$OUT" $D/*.o
