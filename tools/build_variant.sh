#!/bin/bash
# build an A/B variant of the library: tools/build_variant.sh NAME [-DFLAG=..]...
# -> sequnwinder_b200/_ab/lib_NAME.so (git-ignored; travels with gpurun)
set -e
name=$1; shift
cd "$(dirname "$0")/../sequnwinder_b200/csrc"
mkdir -p ../_ab _obj
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v "$@" \
    -c admm.cu -o _obj/admm_$name.o 2> _obj/admm_$name.ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../_ab/lib_$name.so _obj/admm_$name.o _obj/common.o \
    _obj/kmer_count.o _obj/prepare_predict.o _obj/hills_scan.o -ldl
grep -A3 "admm_xupdate_kernelILi1ELi8" _obj/admm_$name.ptxas.log | grep -E "spill"
