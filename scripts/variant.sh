#!/bin/bash
# scripts/variant.sh NAME [nvcc flags...]  -> centroidal_mpc_b200/csrc/variants/libcmpc_NAME.so
# Builds a variant of the point-contact solver (cmpc_api.cu with the given flags, linked with the cmpc_wrench.o
# of the last regular build) for A/B timing on the GPU box (scripts/ab.py).
set -e
cd "$(dirname "$0")/../centroidal_mpc_b200/csrc"
mkdir -p variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC \
  -diag-suppress 550 "$@" -c -o variants/cmpc_api_$name.o cmpc_api.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libcmpc_$name.so variants/cmpc_api_$name.o cmpc_wrench.o
rm -f variants/cmpc_api_$name.o
echo built $name
