#!/bin/bash
# Builds a variant of libogb.so with extra nvcc flags for same-box A/B runs: bash profiles/build_variant.sh name -DOGB_X=1 ...
# Output: profiles/variants/libogb_<name>.so (git-ignored, travels to the GPU box); use with OGB_LIB=$PWD/profiles/variants/libogb_<name>.so
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p profiles/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function "$@" -shared \
  -o profiles/variants/libogb_$name.so metagenomics_b200/csrc/ogb_device.cu metagenomics_b200/csrc/ogb_host.cpp -ldl -lpthread
echo built profiles/variants/libogb_$name.so
