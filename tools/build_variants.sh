#!/bin/bash
# Alternative builds of libdrs_b200.so for A/B measurements (loaded with DRS_LIB_PATH=...).
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -I include"
$NV -DDRS_CHAIN_MIN_CTAS=4 -o tools/bin/libdrs_chain80.so deeprecsys_b200/csrc/engine.cu
$NV -DDRS_CHAIN_MIN_CTAS=3 -o tools/bin/libdrs_chain112.so deeprecsys_b200/csrc/engine.cu
echo built: $(ls tools/bin/*.so)
