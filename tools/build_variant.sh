#!/bin/bash
# build a kernel variant: tools/build_variant.sh NAME -DGG_MIN_CTAS=5 ...   -> /root/repo/gpurun_variants/NAME.so
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants/$name
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-pthread"
for f in gg_api.cu gg_tree_kernel.cu gg_ewald.cu gg_moments.cu gg_tree_gpu.cu gg_state.cu gg_orb.cu gg_comm.cu gg_tree_build.cpp; do
  nvcc $F "$@" -c gasoline_b200/csrc/$f -o gpurun_variants/$name/${f%.*}.o &
done; wait
nvcc -shared -o gpurun_variants/$name.so gpurun_variants/$name/*.o -lcudart -lpthread -ldl 2>/dev/null
echo built gpurun_variants/$name.so
