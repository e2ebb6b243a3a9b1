#!/bin/bash
# Build the library from the current working tree into gpurun_out/<name>.so (A/B timing on one box).
#   usage: bash scripts/build_variant.sh <name> [extra nvcc flags]
NAME=$1; shift
mkdir -p /root/repo/ab_libs
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
  -o /root/repo/ab_libs/$NAME.so /root/repo/human-3d-reconstruction_b200/csrc/smpl_b200.cu && echo built ab_libs/$NAME.so
