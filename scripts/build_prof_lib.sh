#!/bin/bash
# the -DT1_PROFILE build of the library (phase ticks of kernel 1t), next to the product build: libtlc_b200_prof.so
set -e
cd "$(dirname "$0")/../tlc-gnn_b200/csrc"
make -s -j8
mkdir -p build_prof
for f in build/*.o; do cp $f build_prof/; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O2 -fmad=false -DT1_PROFILE -c k1t_sssp_table.cu -o build_prof/k1t_sssp_table.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../tlc_b200/libtlc_b200_prof.so build_prof/*.o -lcudart
