#!/bin/bash
# scratch build of a library variant: scratch/build_variant.sh NAME file.cu "-DFLAG ..."  -> scratch/libmural_NAME.so
# (only file.cu is recompiled with the flags; the other objects come from the product build under mural_b200/csrc/_obj)
set -e
NAME=$1; SRC=$2; FLAGS=$3
cd "$(dirname "$0")/../mural_b200/csrc"
mkdir -p /tmp/mv_obj_$NAME
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $FLAGS -c $SRC -o /tmp/mv_obj_$NAME/${SRC%.cu}.o
OBJS=$(ls _obj/*.o | grep -v "/${SRC%.cu}.o")
nvcc -shared -o ../../scratch/libmural_$NAME.so $OBJS /tmp/mv_obj_$NAME/${SRC%.cu}.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lz
