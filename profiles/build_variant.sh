#!/bin/bash
# build_variant.sh <name> [-D...]: an experimental build of the library into tmp_libs/lib_<name>.so (git-ignored,
# shipped to the GPU box); select it at run time with MAPF_B200_LIB=tmp_libs/lib_<name>.so
cd "$(dirname "$0")/../mapf_marl_b200/csrc" || exit 1
name=$1; shift
mkdir -p ../../tmp_libs
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
  -Xcompiler -pthread -shared "$@" -o ../../tmp_libs/lib_$name.so mapf_kernels.cu mapf_capi.cu mapf_host_unpack.cpp 2>&1 \
  | grep -E "error|warning: v|spill" | head
