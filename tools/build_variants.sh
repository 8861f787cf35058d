#!/bin/bash
# tools/build_variants.sh NAME "-DFLAG ..." : the product library with extra defines -> tools/libclb_NAME.so (select with CLB_LIB_PATH)
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $2 -o tools/libclb_$1.so careless_b200/csrc/clb_api.cu
