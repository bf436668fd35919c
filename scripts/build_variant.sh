#!/bin/bash
# Builds a kernel variant for tuning experiments: scripts/build_variant.sh <name> [-DSK_TRACE_MINBLOCKS=5 ...]
# -> variants/lib_<name>.so (git-ignored, travels to the GPU box); scripts/tune.py times every variant on the same input.
set -e
NAME=$1; shift
mkdir -p variants
cd skirt9_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -shared -Xcompiler -fPIC "$@" -Xptxas -v -o ../../variants/lib_$NAME.so engine.cu 2>&1 | grep -A2 "Compiling entry function '_Z11sk_wf_traceILi2ELi[02]ELb0" | grep -E "Used|spill" | tr '\n' ' ' | sed "s/^/$NAME: /"
echo
