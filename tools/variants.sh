#!/bin/bash
# Kernel-tuning helper: builds alternative libmmo_b200 variants (extra -D flags for direct_fp32.cu) under build/variants/
# so that ONE gpurun call can bench several of them back to back:  MMO_B200_LIB=build/variants/<name>.so python bench.py
#   tools/variants.sh name1 "-DFOO=1" name2 "-DBAR=2 -DBAZ" ...
set -e
cd "$(dirname "$0")/../mmo_b200/csrc"
make -s -j8
mkdir -p ../../build/variants
OBJS="runtime.o host_math.o molecules.o molfile.o strict_fp64.o mask.o desolv.o scan.o api.o mc.o nccl_merge.o"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ \
    -Xcompiler -fPIC,-ffp-contract=off $flags -Xptxas -v -c direct_fp32.cu -o ../../build/variants/$name.o 2> ../../build/variants/$name.ptxas.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../../build/variants/$name.so \
    $OBJS ../../build/variants/$name.o -lcudart_static -ldl -lrt -lpthread
  echo "built build/variants/$name.so ($flags)"
done
