#!/bin/bash
# recompile ONE source of frankenz_b200/csrc into lib/obj and relink libfzb200.so (development shortcut; build.py does all)
set -e
src=$1; shift
cd "$(dirname "$0")/../frankenz_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c csrc/$src.cu -o lib/obj/$src.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a lib/obj/*.o -o lib/libfzb200.so
