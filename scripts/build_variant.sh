#!/bin/bash
# experiment builds of libzelll_b200.so with extra -D flags: scripts/build_variant.sh <name> [-DFOO=1 ...]
# -> build/libzb_<name>.so; select it with ZB_LIB=build/libzb_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC -shared "$@" -o build/libzb_$name.so zelll_b200/csrc/zelll_b200.cu
echo build/libzb_$name.so
