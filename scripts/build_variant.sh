#!/bin/bash
# usage: build_variant.sh R TI MINB  -> variants/lib_R_TI_MINB.so   (tuning only)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -prec-sqrt=true -prec-div=true -ftz=false -Xcompiler -fPIC -shared -cudart static -DTL_SCAN_R=$1 -DTL_SCAN_TI=$2 -DTL_SCAN_MINB=$3 -o variants/lib_$1_$2_$3.so teeline_b200/csrc/*.cu -ldl
