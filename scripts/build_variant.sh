#!/bin/bash
# usage: build_variant.sh <name> [-DTL_...=v ...]  -> variants/lib_<name>.so   (tuning only)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -prec-sqrt=true -prec-div=true -ftz=false -Xcompiler -fPIC -shared -cudart static "$@" -o variants/lib_$name.so teeline_b200/csrc/*.cu -ldl
