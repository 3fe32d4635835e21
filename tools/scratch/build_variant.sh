#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. ..." : libsmm_b200 with abm_dense.cu compiled under extra flags -> socialmediamodel_b200/variants/libsmm_NAME.so
set -e
cd "$(dirname "$0")/../../socialmediamodel_b200/csrc"
mkdir -p ../variants build
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $2 -c abm_dense.cu -o build/abm_dense_$1.o
OBJS=$(ls build/*.o | grep -v "abm_dense")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libsmm_$1.so $OBJS build/abm_dense_$1.o -lpthread -ldl
