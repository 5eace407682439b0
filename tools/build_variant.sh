#!/bin/sh
# Tuning build of the C-ABI library with extra -D switches: tools/build_variant.sh <tag> [-DNAME=VAL ...]
# -> gpurun_in/lib_<tag>.so (select it at run time with BLS381_B200_LIB=gpurun_in/lib_<tag>.so)
set -e
cd "$(dirname "$0")/.."
tag="$1"; shift
mkdir -p gpurun_in
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --extended-lambda \
    -Xcompiler -fPIC -shared -Xptxas -v "$@" -o "gpurun_in/lib_$tag.so" noble_bls12_381_b200/csrc/api.cu -ldl 2> "gpurun_in/lib_$tag.log"
grep -A2 "vm_kernelILi8ELi2" "gpurun_in/lib_$tag.log" | grep -E "spill|Used"
