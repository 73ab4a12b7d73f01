#!/bin/sh
# Build the CPU emulation library of the kernels (development harness only).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/../../melvin.py_b200/csrc"
mkdir -p "$HERE/_build"
g++ -O1 -g -std=c++17 -DMLV_EMU -fPIC -shared -Wall -Wno-unknown-pragmas -Wno-psabi -Wno-unused-function \
    -x c++ "$SRC/mlv_api.cu" -x c++ "$HERE/emu_rt.cpp" -o "$HERE/_build/libmelvin_emu.so"
echo "built $HERE/_build/libmelvin_emu.so"
