#!/bin/bash
# scratch build of the library with in-kernel phase timing of the stage kernels (not the product build)
set -e
cd "$(dirname "$0")/../mural_b200/csrc"
mkdir -p /tmp/mt_obj
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DMURAL_TC_TIMING -c $f -o /tmp/mt_obj/${f%.cu}.o &
done
wait
nvcc -shared -o ../../scratch/libmural_timing.so /tmp/mt_obj/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
ls -la ../../scratch/libmural_timing.so
