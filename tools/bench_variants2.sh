#!/bin/bash
# usage: tools/bench_variants2.sh <tag> <glob> [bench args]  -- one kernels-only bench line per library matching variants/<glob>
tag=$1; pat=$2; shift; shift
for so in variants/$pat; do
  n=$(basename $so .so)
  SRM_PHYSICS_LIB=$PWD/$so timeout 300 python bench.py --steps 10 --warmup 3 --kernels-only "$@" > gpurun_out/bv_${tag}_$n.json 2> gpurun_out/bv_${tag}_$n.err
  echo "$n $(cat gpurun_out/bv_${tag}_$n.json | cut -c1-300)"
done
