#!/bin/bash
# A/B library builds: scripts/build_variant.sh <name> <file.cu> "<extra nvcc flags>"  ->  ab/libpde_<name>.so
# (one translation unit recompiled with the flags, the rest taken from pde_solver_b200/build/); select at run time
# with PDE_B200_LIB=ab/libpde_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; flags=$3
mkdir -p ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
  -x cu -c pde_solver_b200/csrc/$src -o ab/$name.$src.o
objs=""
for o in pde_solver_b200/build/*.o; do
  case "$o" in *"/$src.o") objs="$objs ab/$name.$src.o";; *) objs="$objs $o";; esac
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ab/libpde_$name.so $objs -ldl
rm -f ab/$name.$src.o
echo ab/libpde_$name.so
