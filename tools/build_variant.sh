#!/bin/sh
# Build a compile-time variant of the library for A/B measurements:
#   tools/build_variant.sh <name> [-DMACRO ...]   ->  melvin.py_b200/melvin/_lib/variant_<name>.so
# run with MLV_LIB=melvin.py_b200/melvin/_lib/variant_<name>.so
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME=$1; shift
cd "$ROOT/melvin.py_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -diag-suppress 186 -shared -Xcompiler -fPIC "$@" \
    -o "$ROOT/melvin.py_b200/melvin/_lib/variant_$NAME.so" mlv_api.cu
echo "built variant_$NAME.so"
