#!/bin/bash
# builds tuning variants of the closed-form lean pair into variants/cf2_*.so
# usage: tools/cf2_variants.sh "NT OCCF OCCA SF SA LXMAX [extra -D flags]" ...
set -e
cd "$(dirname "$0")/../3d-physics-based-ai-surrogate-reservoir-model_b200/csrc"
mkdir -p ../../variants
for cfg in "$@"; do
  set -- $cfg
  name="cf2_$1_$2_$3_$4_$5_$6"
  extra="${@:7}"
  [ -n "$extra" ] && name="${name}_$(echo $extra | tr -d ' =-' )"
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=true \
    -DCF2_NT=$1 -DCF2_OCCF=$2 -DCF2_OCCA=$3 -DCF2_SF=$4 -DCF2_SA=$5 -DCF2_LXMAX=$6 $extra -c kernels_cf2.cu -o /tmp/$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/$name.so capi.o kernels_ref.o kernels_ref2.o kernels_dg4.o kernels_gc.o kernels_cf.o /tmp/$name.o kernels_misc.o kernels_glue.o -lcudart
  echo built variants/$name.so
done
