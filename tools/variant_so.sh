#!/bin/bash
# Build a variant of libcdm_b200.so with extra -D flags on ONE source file (kernel experiments):
#   tools/variant_so.sh conv_tc3.cu NAME -DFOO -DBAR   ->  composable_diffusion_models_b200/build/libvar_NAME.so
# Run with  CDM_LIB_PATH=composable_diffusion_models_b200/build/libvar_NAME.so python tools/layer_times.py
set -e
cd "$(dirname "$0")/.."
src=$1; name=$2; shift 2
B=composable_diffusion_models_b200/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c composable_diffusion_models_b200/csrc/$src -o $B/var_$name.o
objs=$(ls $B/*.o | grep -v "/var_" | grep -v "/${src%.cu}.o")
nvcc -shared -cudart static -o $B/libvar_$name.so $objs $B/var_$name.o
echo $B/libvar_$name.so
