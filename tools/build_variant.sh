#!/bin/bash
# usage: tools/build_variant.sh <name> "<extra nvcc -D flags>"  -> rcbevdet_b200/lib/variants/lib_<name>.so
# Experiment helper: builds a copy of the library with extra defines; select it at run time with
# RCB_LIB_PATH=<path>.  The variants directory is git-ignored.
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
OUT=rcbevdet_b200/lib/variants; mkdir -p $OUT /tmp/rcbv_$NAME
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-extended-lambda -Xcompiler -fPIC --fmad=true --prec-div=true --ftz=false -Xptxas -v $*"
for f in api prepare prepare_lsd pool_plan pool_fwd pool_fwd_cells pool_bwd strips layout radar bev_shift depth_context trt_plugin; do
  nvcc $FLAGS -c rcbevdet_b200/csrc/$f.cu -o /tmp/rcbv_$NAME/$f.o > /tmp/rcbv_$NAME/$f.log 2>&1 &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/lib_$NAME.so /tmp/rcbv_$NAME/*.o
echo $OUT/lib_$NAME.so
