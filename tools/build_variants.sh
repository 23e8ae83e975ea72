#!/bin/bash
# builds tuning variants of libsrm_physics.so (kernels_dg4.cu tile parameters) into variants/ (git-ignored *.so)
# usage: tools/build_variants.sh "CX CPT TY OCCF OCCA" ...
set -e
cd "$(dirname "$0")/../3d-physics-based-ai-surrogate-reservoir-model_b200/csrc"
mkdir -p ../../variants
for cfg in "$@"; do
  set -- $cfg
  name="v_$1_$2_$3_$4_$5"
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=true \
    -DSRM_D4_CX=$1 -DSRM_D4_CPT=$2 -DSRM_D4_TY=$3 -DSRM_D4_OCCF=$4 -DSRM_D4_OCCA=$5 $EXTRA_DEFS -c kernels_dg4.cu -o /tmp/$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/$name.so capi.o kernels_ref.o kernels_ref2.o /tmp/$name.o kernels_gc.o kernels_cf.o kernels_cf2.o kernels_misc.o kernels_glue.o -lcudart
  echo built variants/$name.so
done
