#!/bin/bash
# developer builds of the library with the kNN timing switches (SCF_KNN_DEBUG) compiled in: csrc/build/libscarf_b200_dbgN.so
set -e
cd "$(dirname "$0")/../scarf_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC"
OTHERS=$(ls build/*.o | grep -v knn_tc | grep -v dbg)
for d in "$@"; do
  nvcc $FLAGS -DSCF_KNN_DEBUG=$d -c knn_tc.cu -o build/knn_tc_dbg$d.o
  nvcc -shared -o build/libscarf_b200_dbg$d.so $OTHERS build/knn_tc_dbg$d.o -gencode arch=compute_100a,code=sm_100a
done
