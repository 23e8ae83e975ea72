#!/bin/bash
# ncu full capture of the cf2 pair for the default build and the CPT=2 variant
mkdir -p gpurun_out
for tag in cpt4 cpt2; do
  if [ $tag = cpt2 ]; then export SRM_PHYSICS_LIB=$PWD/variants/cf2_256_4_3_3_3_32_DCF2_CPT2.so; else unset SRM_PHYSICS_LIB; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_(fwd|adj)_cf2" -s 2 -c 2 -o gpurun_out/prof_r2_$tag -f \
      python tools/prof_step.py cfg5 closed_form 2 3 > gpurun_out/ncu_f_r2_$tag.log 2>&1
  tail -2 gpurun_out/ncu_f_r2_$tag.log
done
