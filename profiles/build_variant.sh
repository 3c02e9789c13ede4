#!/bin/bash
# usage: profiles/build_variant.sh <name> [nvcc flags]  -- libclr_b200_<name>.so whose deepz_ks25 unit (the C4 network's kernel) is
# compiled with the given flags; every other object comes from the main build (run `make -C .../csrc` first).
set -e
name=$1; shift
cd "$(dirname "$0")/../closedloopreachability.jl_b200/csrc"
mkdir -p build_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c -o build_$name/deepz_ks25.o deepz_ks25.cu 2> build_$name/deepz_ks25.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libclr_b200_$name.so build/clr_api.o build/deepz_ks8.o build/deepz_ks16.o build_$name/deepz_ks25.o build/deepz_ks32.o -ldl
grep -A2 "Lb0ELb0ELb0" build_$name/deepz_ks25.log | grep -E "registers|spill" | tr '\n' ' '; echo " -> libclr_b200_$name.so"
