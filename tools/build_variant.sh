#!/bin/bash
# build/variants/<name>.so with extra -D flags (development aid): tools/build_variant.sh name -DDRTB_X=1 ...
# All four translation units are compiled in parallel with the flags of differentiable-renderer_b200/build.py.
name=$1; shift
mkdir -p build/variants/obj_$name
C=differentiable-renderer_b200/csrc
for tu in drtb render_f64 render_f32 mesh multi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
    -DDRTB_MIN_BLOCKS=7 -DDRTB_MIN_BLOCKS_F32=7 -DDRTB_MESH_MIN_BLOCKS=6 -DDRTB_WF_MIN_BLOCKS=7 "$@" -ccbin /usr/bin/g++ -I include \
    -c -o build/variants/obj_$name/$tu.o $C/$tu.cu 2> build/variants/obj_$name/$tu.log &
done
wait
nvcc -shared -cudart static -ccbin /usr/bin/g++ -o build/variants/$name.so build/variants/obj_$name/*.o && echo build/variants/$name.so
