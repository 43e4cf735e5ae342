#!/bin/bash
# Builds alternative libhtm_b200 libraries under variants/ (git-ignored, they travel with gpurun) that differ only in
# compile-time tuning macros of one translation unit; tools/variant_sweep.py times them through HTM_B200_LIB.
#   tools/build_variants.sh <name> <file.cu> "<-D flags>" ...
set -e
cd "$(dirname "$0")/../hypotremormcmc_b200/csrc"
mkdir -p ../../variants
make -s -j8
while [ $# -ge 3 ]; do
  name=$1; src=$2; flags=$3; shift 3
  obj=/tmp/variant_${name}.o
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall $flags -c $src -o $obj
  others=$(ls *.o | grep -v "^${src%.cu}.o$")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libhtm_${name}.so $obj $others -ldl -lpthread
  echo "built variants/libhtm_${name}.so ($flags)"
done
