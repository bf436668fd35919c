#!/bin/bash
# Builds a kernel variant for tuning experiments: scripts/build_variant.sh <name> [-DSK_POOL=512 ...]
# -> gpurun_out/variants/lib_<name>.so ; select at run time with SK_ENGINE_LIB=<path>.
set -e
NAME=$1; shift
mkdir -p variants
cd skirt9_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -shared -Xcompiler -fPIC "$@" -Xptxas -v -o ../../variants/lib_$NAME.so engine.cu 2>&1 | grep -A1 "sk_life_cycle_kernelILi2" | grep -E "registers|spill" | sed "s/^/$NAME: /"
