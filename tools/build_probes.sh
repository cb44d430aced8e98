#!/bin/bash
# Build the standalone measurement tools into tools/bin/ (git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I deeprecsys_b200/csrc"
$NV -o tools/bin/hbm_probe tools/hbm_probe.cu
$NV -o tools/bin/k1_probe tools/k1_probe.cu
$NV -DDRS_SLS_UNROLL=16 -o tools/bin/k1_probe_u16 tools/k1_probe.cu
$NV -DDRS_SLS_UNROLL=10 -o tools/bin/k1_probe_u10 tools/k1_probe.cu
$NV -DDRS_SLS_UNROLL=6 -o tools/bin/k1_probe_u6 tools/k1_probe.cu
$NV -DDRS_ROW_EVICT_FIRST=1 -o tools/bin/k1_probe_ef tools/k1_probe.cu
echo built: $(ls tools/bin)
