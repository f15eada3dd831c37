#!/bin/bash
# Compiles tests/cpp_client/ref_wrapper_client.cpp together with the reference's own, unmodified C++ wrapper
# (/root/reference/include/mlf.hpp + src/cpp_intf/mlfcpp.cpp, read where they lie -- never copied) against
# libmlfortran_b200.so.  Output: tests/cpp_client/_bin/ (git-ignored, travels to the GPU box like the built .so).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${MLF_REFERENCE:-/root/reference}"
[ -f "$REF/include/mlf.hpp" ] || { echo "reference not present at $REF: nothing built"; exit 3; }
CXX=/usr/bin/g++; [ -x $CXX ] || CXX=g++
mkdir -p "$HERE/_bin"
$CXX -std=c++17 -O1 -Wall -Wno-unused-variable -Wno-implicit-fallthrough -I "$REF/include" "$HERE/ref_wrapper_client.cpp" "$REF/src/cpp_intf/mlfcpp.cpp" \
  -o "$HERE/_bin/ref_wrapper_client" -L "$ROOT/mlfortran_b200/lib" -lmlfortran_b200 '-Wl,-rpath,$ORIGIN/../../../mlfortran_b200/lib' -ldl
echo "built $HERE/_bin/ref_wrapper_client"
