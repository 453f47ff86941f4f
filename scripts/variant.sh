#!/bin/bash
# scripts/variant.sh NAME [nvcc flags...]  -> centroidal_mpc_b200/csrc/variants/libcmpc_NAME.so
# Builds a variant of the CUDA library for A/B timing on the GPU box (scripts/ab.sh).
set -e
cd "$(dirname "$0")/../centroidal_mpc_b200/csrc"
mkdir -p variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -shared -Xcompiler -fPIC \
  -diag-suppress 550 "$@" -o variants/libcmpc_$name.so cmpc_api.cu
echo built $name
