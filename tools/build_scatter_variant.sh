#!/bin/bash
# Development: A/B build of the scatter kernel.  tools/build_scatter_variant.sh NAME [-DMACRO=...]  ->  build/variants/libdrr_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants build/obj_$name
F="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
nvcc $F "$@" -c -o build/obj_$name/drr_scatter.o deepdrr_b200/csrc/drr_scatter.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libdrr_$name.so build/obj/drr_capi.o build/obj/drr_march.o build/obj/drr_march_warp.o build/obj/drr_march_warp_r1.o build/obj/drr_mesh.o build/obj_$name/drr_scatter.o build/obj/drr_spectral.o
echo built build/variants/libdrr_$name.so
