#!/bin/bash
# usage: tools/build_variant.sh <name> [-DADP_...=...]   -> tools/ab/lib_<name>.so: the library with cmfd_kernels.cu and
# nodal_kernels.cu compiled with the given macros (A/B runs with tools/ab_step.py, tools/nodal_ab.py); the other objects come
# from adpres_b200/build/
set -e
cd "$(dirname "$0")/.."
name=$1; shift
python -c "from adpres_b200 import build; build.build()" > /dev/null
mkdir -p tools/ab /tmp/abobj
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off"
nvcc $F "$@" -c adpres_b200/csrc/cmfd_kernels.cu -o /tmp/abobj/cmfd_$name.o &
nvcc $F "$@" -Xptxas -v -c adpres_b200/csrc/nodal_kernels.cu -o /tmp/abobj/nodal_$name.o 2> /tmp/abobj/nodal_$name.log &
wait
objs=$(ls adpres_b200/build/*.o | grep -v "cmfd_kernels.o\|nodal_kernels.o")
nvcc -shared -o tools/ab/lib_$name.so /tmp/abobj/cmfd_$name.o /tmp/abobj/nodal_$name.o $objs -ldl -gencode arch=compute_100a,code=sm_100a
echo tools/ab/lib_$name.so
