#!/bin/bash
# build/variants/<name>.so with extra -D flags (development aid): tools/build_variant.sh name -DDRTB_X=1 ...
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -cudart static \
  -DDRTB_MIN_BLOCKS=5 -DDRTB_MIN_BLOCKS_F32=7 -DDRTB_MESH_MIN_BLOCKS=6 "$@" -ccbin /usr/bin/g++ -I include \
  -o build/variants/$name.so differentiable-renderer_b200/csrc/drtb.cu
