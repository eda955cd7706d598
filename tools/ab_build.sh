#!/bin/bash
# A/B builds of ONE translation unit: tools/ab_build.sh <tag> <file.cu> [-DDEFS...] -> tools/_ab/lib_<tag>.so (other objects from the
# last regular build); run with DXB200_LIB=tools/_ab/lib_<tag>.so
set -e
tag=$1; src=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/tools/_ab
obj=$root/tools/_ab/${tag}_$(basename $src .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I $root/include -I $root/deepxube_b200/csrc "$@" -c $root/deepxube_b200/csrc/$src -o $obj
others=$(ls $root/deepxube_b200/build/*.o | grep -v "/$(basename $src .cu).o")
nvcc -shared -o $root/tools/_ab/lib_${tag}.so $obj $others -gencode arch=compute_100a,code=sm_100a -lcudart
echo $root/tools/_ab/lib_${tag}.so
